#!/bin/bash
# Run ON THE GPU BOX: full ncu capture of one kernel (one launch, config 3).  usage: prof1.sh <kernel> <tag>
K=$1; TAG=${2:-x}
CMD="python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/plain_${TAG}.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:$K -s 20 -c 1 -f -o gpurun_out/prof_${K}_${TAG} $CMD > gpurun_out/ncu_${K}_${TAG}.log 2>&1
