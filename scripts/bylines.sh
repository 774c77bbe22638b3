#!/bin/bash
# usage: bylines.sh <kernel> <tag> [top]  -- per-source-line table of a capture made by prof1.sh/prof3.sh (run in the dev container,
# with the SAME libsigner_gibbs.so that was profiled)
K=$1; TAG=$2; TOP=${3:-200}; LIB=${4:-/root/repo/signer_b200/libsigner_gibbs.so}
cd /root/repo
ncu -i gpurun_out/prof_${K}_${TAG}.ncu-rep --page source --csv > gpurun_out/sass_${K}_${TAG}.csv 2>/dev/null
(cd /tmp && rm -f *.cubin && cuobjdump -xelf all $LIB > /dev/null && nvdisasm -g -c /tmp/*.cubin > /tmp/sass_lines.txt)
python scripts/ncu_by_line.py gpurun_out/sass_${K}_${TAG}.csv /tmp/sass_lines.txt $K $TOP
