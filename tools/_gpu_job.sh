python -m pytest tests -m gpu -q -x 2>&1 | tail -3
sh tools/sweep.sh
echo "greedy collapse:" | tee -a gpurun_out/sweep.txt
FB200_BVH_COLLAPSE=greedy python bench.py --steps 16 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('greedy     %7.1f Msamples/s  %6.3f ms/pass | trace %.3f shade %.3f shadow %.3f' % (d['value'], d['ms_per_step'], d['kernels']['trace']['ms_per_launch'], d['kernels']['shade']['ms_per_launch'], d['kernels']['shadow']['ms_per_launch']))" | tee -a gpurun_out/sweep.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_trace --launch-skip 8 -c 4 -f -o gpurun_out/r01d_trace python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_d.log 2>&1
tail -2 gpurun_out/ncu_d.log
