#!/bin/bash
# Run ON A MULTI-GPU BOX: N processes, one per GPU, make the stateless calls of scripts/e2e_short.py at the same time with
# the phase trace on -- shows which phase of a call grows when the ranks share the host.  usage: e2e_concurrent.sh N
N=${1:-4}
nproc; lscpu | grep -E "Socket|NUMA node\(s\)|Model name" | head -4
for g in $(seq 0 $((N-1))); do
  CUDA_VISIBLE_DEVICES=$g SIGNER_TRACE=2 python scripts/e2e_short.py > gpurun_out/e2e_conc_$g.log 2>&1 &
done
wait
for g in $(seq 0 $((N-1))); do echo "== gpu $g"; grep -B2 "10+10 keep=F  *rep [23]" gpurun_out/e2e_conc_$g.log | grep -v "^--" | head -6; done
