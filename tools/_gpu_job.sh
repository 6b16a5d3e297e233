#!/bin/sh
mkdir -p gpurun_out
FB200_TRACE_CTAS=4 timeout 800 ncu --set full --clock-control none --import-source on -k regex:"k_trace|k_shade|k_accumulate" --launch-skip 144 -c 72 -f -o gpurun_out/r03_full_pass python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r03_ncu_full_pass.log 2>&1
ls -la gpurun_out/r03_full_pass.ncu-rep
