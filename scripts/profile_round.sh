#!/bin/bash
# Run ON THE GPU BOX (under gpurun): plain bench, ncu launch list of the same command, full captures of the
# two dominant kernels.  Outputs land in gpurun_out/; scripts/summarize_profiles.py (run in the dev container)
# turns them into the text summaries committed under profiles/.
set -x
TAG=${1:-r01}
CMD="python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-e2e --no-extras"
python bench.py > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
$CMD > gpurun_out/plain_${TAG}.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${TAG}.csv $CMD > gpurun_out/ncu_launches_${TAG}.log 2>&1
$CMD > gpurun_out/plain_${TAG}.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:z_alloc_fused -s 20 -c 1 -f -o gpurun_out/prof_z_${TAG} $CMD > gpurun_out/ncu_z_${TAG}.log 2>&1
$CMD > gpurun_out/plain_${TAG}.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:e_family -s 20 -c 1 -f -o gpurun_out/prof_e_${TAG} $CMD > gpurun_out/ncu_e_${TAG}.log 2>&1
$CMD > gpurun_out/plain_${TAG}.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:p_family -s 20 -c 1 -f -o gpurun_out/prof_p_${TAG} $CMD > gpurun_out/ncu_p_${TAG}.log 2>&1
ls -la gpurun_out
