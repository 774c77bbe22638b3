#!/bin/bash
# Run ON THE GPU BOX: full ncu captures of the three sweep kernels (one launch each, config 3).
TAG=${1:-x}
CMD="python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/plain_${TAG}.log 2>&1 || exit 1
for k in z_alloc_fused e_family p_family; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 20 -c 1 -f -o gpurun_out/prof_${k}_${TAG} $CMD > gpurun_out/ncu_${k}_${TAG}.log 2>&1
done
ls -la gpurun_out
