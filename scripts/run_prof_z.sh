# ncu full capture of z_alloc_fused (after a plain run of the same command exits 0)
set -x
CMD="python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:z_alloc_fused -s 20 -c 1 -o gpurun_out/prof_z $CMD > gpurun_out/ncu2.log 2>&1
ls -la gpurun_out
