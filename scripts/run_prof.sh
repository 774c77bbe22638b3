set -x
python bench.py --steps 300 --warmup 20 > gpurun_out/bench_small.json 2> gpurun_out/bench_small.err; tail -3 gpurun_out/bench_small.err; cat gpurun_out/bench_small.json
python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/ncu1.log 2>&1
python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:z_alloc_fused -s 20 -c 2 -o gpurun_out/prof_z python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/ncu2.log 2>&1
ls -la gpurun_out
