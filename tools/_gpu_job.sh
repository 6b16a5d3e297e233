python -m pytest tests -m gpu -q -x 2>&1 | tail -3
for cfg in "2 2" "3 1" "4 1" "1 0"; do
set -- $cfg
echo "subframes $1 ctas $2" | tee -a gpurun_out/sweep.txt
FB200_SUBFRAMES=$1 FB200_TRACE_CTAS=$2 python bench.py --steps 16 --warmup 3 --no-cpu-baseline 2>gpurun_out/sf.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('   %7.1f Msamples/s  %6.3f ms/pass e2e %7.1f | trace %.3f shade %.3f shadow %.3f finite %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['kernels']['trace']['ms_per_launch'], d['kernels']['shade']['ms_per_launch'], d['kernels']['shadow']['ms_per_launch'], d['finite']))" | tee -a gpurun_out/sweep.txt
tail -3 gpurun_out/sf.err | cut -c1-300
done
