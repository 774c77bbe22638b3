# usage: run_prof_k.sh <kernel regex> <out name> ; ncu full capture after a plain run exits 0
set -x
CMD="python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/plain_$2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:$1 -s 20 -c 1 -o gpurun_out/prof_$2 $CMD > gpurun_out/ncu_$2.log 2>&1
