#!/bin/sh
# one GPU-box visit: GPU test suite, kernel-variant sweep, extras (LBVH, post pipeline, 1024-spp parity)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv > gpurun_out/r02_gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02_pytest_gpu.log
tail -5 gpurun_out/r02_pytest_gpu.log
rm -f gpurun_out/sweep.txt
STEPS=16 timeout 600 sh tools/sweep.sh
timeout 900 python tools/extras_gpu.py --sections lbvh,post,parity > gpurun_out/r02_extras.json 2> gpurun_out/r02_extras.err
tail -c 1500 gpurun_out/r02_extras.err
