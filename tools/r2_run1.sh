#!/bin/sh
# round 2, GPU run 1: the new parity tests, the bench, and the sweeps round 1 left on host evidence + CTA-size variants
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv > gpurun_out/r2a_gpu.txt; nproc >> gpurun_out/r2a_gpu.txt
timeout 1200 python -m pytest tests -x -q -m gpu -s > gpurun_out/r2a_pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/r2a_pytest_gpu.log
grep -a "passed\|failed\|rel L2\|hits differ\|Error\|error" gpurun_out/r2a_pytest_gpu.log | tail -20
timeout 400 python bench.py --steps 24 --warmup 4 > gpurun_out/r2a_bench_n1.json 2> gpurun_out/r2a_bench_n1.err
python tools/sweep.py r2a base far:FB200_SHADOW_ORDER=far auto:FB200_SHADOW_ORDER=auto opt0:FB200_BVH_OPT=0 opt0far:FB200_BVH_OPT=0:FB200_SHADOW_ORDER=far sbvh:FB200_BVH_BUILDER=sbvh \
  s3c1:FB200_SUBFRAMES=3:FB200_TRACE_CTAS=1 s3c2:FB200_SUBFRAMES=3:FB200_TRACE_CTAS=2 s4c1:FB200_SUBFRAMES=4:FB200_TRACE_CTAS=1 \
  t128:lib=t128:FB200_TRACE_CTAS=4 t128s0:lib=t128s0:FB200_TRACE_CTAS=4 t128s3:lib=t128:FB200_TRACE_CTAS=3:FB200_SUBFRAMES=3 t64:lib=t64:FB200_TRACE_CTAS=8 t64c6s3:lib=t64:FB200_TRACE_CTAS=6:FB200_SUBFRAMES=3 \
  sp1:lib=sp1 sp2:lib=sp2 base2
