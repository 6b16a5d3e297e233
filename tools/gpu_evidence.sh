#!/bin/sh
# One gpurun call that produces the evidence set a round closes with (about 5 GPU-minutes on one B200):
#   /usr/local/graft/bin/gpurun --timeout 2400 -- 'sh tools/gpu_evidence.sh r04'
# writes gpurun_out/<tag>_*: GPU tests, bench N=1 (+ CPU baseline), reference arm, ncu launch list, ncu --set full of 12
# launches (bounces 0..2 of one sub-frame), DRAM traffic of one whole pass, per-bounce anatomy of the trace launches.
# Copy what is to be judged into profiles/ (tools/ncu_summary.py condenses the .ncu-rep: see profiles/README.md).
# Keep ncu --set full captures to a dozen launches: 72 launches with sources exceed gpurun's 64 MiB return limit.
TAG=${1:-rXX}
# (round 2: the sweeps of tools/sweep.py run first when SWEEP is set)
[ -n "$SWEEP" ] && python tools/sweep.py $TAG $SWEEP
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv > gpurun_out/${TAG}_gpu.txt
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/${TAG}_pytest_gpu.log
tail -3 gpurun_out/${TAG}_pytest_gpu.log
timeout 400 python bench.py --steps 24 --warmup 4 > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err
python -c "
import json
d=json.loads(open('gpurun_out/${TAG}_bench_n1.json').read().strip().splitlines()[-1]); print('N=1', d['value'], d['e2e']['value'], d['roofline']['frac'], d['cpu_baseline']['value'], d['clocks'])"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_reference_arm.json 2> gpurun_out/${TAG}_bench_reference_arm.err
# launch list (kernels run one at a time under the profiler: let a trace launch take all four CTA slots as in bench.py's spans)
FB200_TRACE_CTAS=4 timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/${TAG}_launches_bench_steps2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_list.log 2>&1
# 36 matching launches per sub-frame pass (trace, shade, shadow trace, accumulate x 9 bounces), host enqueue order: 144 = start of a pass
FB200_TRACE_CTAS=4 timeout 500 ncu --set full --clock-control none --import-source on -k regex:"k_trace|k_shade|k_accumulate" --launch-skip 144 -c 12 -f -o gpurun_out/${TAG}_full python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_full.log 2>&1
FB200_TRACE_CTAS=4 timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:"k_trace|k_shade|k_accumulate" --launch-skip 144 -c 72 --csv --log-file gpurun_out/${TAG}_traffic_pass.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_traffic.log 2>&1
if [ -f fermat_b200/variants/libfermat_b200_stats.so ]; then       # tools/build_variants.sh stats:"-DFB_TRACE_STATS=1"
  FERMAT_B200_LIB=$PWD/fermat_b200/variants/libfermat_b200_stats.so timeout 300 python tools/trace_stats.py > gpurun_out/${TAG}_trace_anatomy.json 2> gpurun_out/${TAG}_trace_anatomy.txt
  tail -10 gpurun_out/${TAG}_trace_anatomy.txt | cut -c1-200
fi
