#!/bin/bash
# Run ON THE GPU BOX: bench with different batch-width gradings of z_alloc_fused's work list (SIGNER_Z_WIDTHS)
for w in ${WIDTHS:-"99999999:32" "2048:4" "1024:4" "1024:8" "512:8" "1024:4,512:8" "512:16" "2048:2,1024:4,512:8" "1024:8,512:16" "1024:2,512:4,384:8"} ; do
  SIGNER_Z_WIDTHS=$w python bench.py --no-cpu-baseline --no-e2e --no-extras --steps ${STEPS:-1000} --warmup 100 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('%-44s %8.1f sweeps/s  %s' % ('$w', d['value'], {k:(round(x,4) if x else x) for k,x in d['kernel_ms'].items()}))"
done
