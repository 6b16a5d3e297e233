#!/bin/sh
# Run bench.py (short, no CPU baseline) for every library variant under fermat_b200/variants/ and print one line each.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for so in fermat_b200/variants/libfermat_b200_*.so; do
  name=$(basename $so .so | sed 's/libfermat_b200_//')
  FERMAT_B200_LIB=$PWD/$so python bench.py --steps ${STEPS:-16} --warmup 3 --no-cpu-baseline $EXTRA 2> gpurun_out/sweep_$name.err | python -c "
import json,sys
l=sys.stdin.read().strip().splitlines()
d=json.loads(l[-1]) if l else None
if d: print('%-10s %7.1f Msamples/s  %6.3f ms/pass  e2e %7.1f | trace %.3f ms  shade %.3f ms  shadow %.3f ms (per launch)' % ('$name', d['value'], d['ms_per_step'], d['e2e']['value'], d['kernels']['trace']['ms_per_launch'], d['kernels']['shade']['ms_per_launch'], d['kernels']['shadow']['ms_per_launch']))
else: print('$name FAILED')
" | tee -a gpurun_out/sweep.txt
done
