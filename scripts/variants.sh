#!/bin/bash
# Run ON THE GPU BOX: bench every build variant signer_b200/variants/*.so (made in the dev container with -D switches)
# by copying it over the product library of this scratch copy.  Prints sweeps/s and per-kernel ms.
cp signer_b200/libsigner_gibbs.so /tmp/base.so
for v in /tmp/base.so signer_b200/variants/*.so; do
  cp $v signer_b200/libsigner_gibbs.so
  python bench.py --no-cpu-baseline --no-e2e --no-extras --steps ${STEPS:-1000} --warmup 100 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('%-40s %8.1f sweeps/s  %s' % ('$v'.split('/')[-1], d['value'], {k:(round(x,4) if x else x) for k,x in d['kernel_ms'].items()}))"
done
cp /tmp/base.so signer_b200/libsigner_gibbs.so
