#!/bin/sh
mkdir -p gpurun_out
rm -f gpurun_out/sweep.txt
run() {
  name="$1"; shift
  env "$@" timeout 200 python bench.py --steps 16 --warmup 3 --no-cpu-baseline 2> gpurun_out/sweep_$name.err | python -c "
import json,sys
l=sys.stdin.read().strip().splitlines()
d=json.loads(l[-1]) if l else None
if d: print('%-24s %7.1f Msamples/s  e2e %7.1f | trace %.3f shade %.3f shadow %.3f | launches %d' % ('$name', d['value'], d['e2e']['value'], d['kernels']['trace']['ms_per_launch'], d['kernels']['shade']['ms_per_launch'], d['kernels']['shadow']['ms_per_launch'], d['gpu_launches']))
else: print('$name FAILED')" | tee -a gpurun_out/sweep.txt
}
for v in base rl2 rl4 st0 st16 sh7 sh5 thin16 base; do
run $v FERMAT_B200_LIB=$PWD/fermat_b200/variants/libfermat_b200_$v.so
done
