#!/bin/sh
mkdir -p gpurun_out
rm -f gpurun_out/sweep.txt
for v in base sh7; do
  FERMAT_B200_LIB=$PWD/fermat_b200/variants/libfermat_b200_$v.so python bench.py --steps 16 --warmup 3 --no-cpu-baseline 2> /dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('%-40s %7.1f Msamples/s  e2e %7.1f | trace %.3f shade %.3f shadow %.3f' % ('$v', d['value'], d['e2e']['value'], d['kernels']['trace']['ms_per_launch'], d['kernels']['shade']['ms_per_launch'], d['kernels']['shadow']['ms_per_launch']))" | tee -a gpurun_out/sweep.txt
done
for v in t128 t128s8; do for c in 3 4 6; do
  FB200_TRACE_CTAS=$c FERMAT_B200_LIB=$PWD/fermat_b200/variants/libfermat_b200_$v.so python bench.py --steps 16 --warmup 3 --no-cpu-baseline 2> /dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('%-40s %7.1f Msamples/s  e2e %7.1f | trace %.3f shade %.3f shadow %.3f' % ('$v ctas=$c', d['value'], d['e2e']['value'], d['kernels']['trace']['ms_per_launch'], d['kernels']['shade']['ms_per_launch'], d['kernels']['shadow']['ms_per_launch']))" | tee -a gpurun_out/sweep.txt
done; done
