#!/bin/bash
# Run ON THE GPU BOX: bench one variant library under several settings of one environment variable.
# usage: variants_env.sh <lib.so> <VAR> <value> [value ...]
LIBV=$1; VAR=$2; shift 2
cp signer_b200/libsigner_gibbs.so /tmp/base.so; cp $LIBV signer_b200/libsigner_gibbs.so
for v in "$@"; do
  env $VAR=$v python bench.py --no-cpu-baseline --no-e2e --no-extras --steps ${STEPS:-1000} --warmup 100 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('%-30s %8.1f sweeps/s  %s' % ('$VAR=$v', d['value'], {k:(round(x,4) if x else x) for k,x in d['kernel_ms'].items()}))"
done
cp /tmp/base.so signer_b200/libsigner_gibbs.so
