#!/bin/sh
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_psfpt.py -x -q -m gpu > gpurun_out/r03_pytest_psfpt.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/r03_pytest_psfpt.log
grep -v "^  \|allocating\|settings" gpurun_out/r03_pytest_psfpt.log | tail -30
FB200_TRACE_CTAS=4 timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:"k_trace|k_shade|k_accumulate" --launch-skip 144 -c 72 --csv --log-file gpurun_out/r03_traffic_pass.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r03_ncu_traffic.log 2>&1
wc -l gpurun_out/r03_traffic_pass.csv
