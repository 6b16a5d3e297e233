#!/bin/sh
mkdir -p gpurun_out
rm -f gpurun_out/sweep.txt
for v in ap3thin; do
FERMAT_B200_LIB=$PWD/fermat_b200/variants/libfermat_b200_$v.so timeout 400 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "closest or shadow or big_scenes or deterministic or degenerate or render_matches or full_size or shards" > gpurun_out/r03_pytest_$v.log 2>&1; echo "pytest $v exit $?" | tee -a gpurun_out/r03_pytest_$v.log
tail -3 gpurun_out/r03_pytest_$v.log
done
run() {
  name="$1"; shift
  env "$@" timeout 200 python bench.py --steps 16 --warmup 3 --no-cpu-baseline 2> gpurun_out/sweep_$name.err | python -c "
import json,sys
l=sys.stdin.read().strip().splitlines()
d=json.loads(l[-1]) if l else None
if d: print('%-24s %7.1f Msamples/s  e2e %7.1f | trace %.3f shade %.3f shadow %.3f | launches %d' % ('$name', d['value'], d['e2e']['value'], d['kernels']['trace']['ms_per_launch'], d['kernels']['shade']['ms_per_launch'], d['kernels']['shadow']['ms_per_launch'], d['gpu_launches']))
else: print('$name FAILED')" | tee -a gpurun_out/sweep.txt
}
for v in ap apr aplist apseg ap3 ap3thin; do
run $v FERMAT_B200_LIB=$PWD/fermat_b200/variants/libfermat_b200_$v.so
done
for v in statsap3; do
FERMAT_B200_LIB=$PWD/fermat_b200/variants/libfermat_b200_$v.so timeout 300 python tools/trace_stats.py > gpurun_out/r03_trace_stats_$v.json 2> gpurun_out/r03_trace_stats_$v.txt
tail -10 gpurun_out/r03_trace_stats_$v.txt | cut -c1-400
done
