#!/bin/sh
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu -s > gpurun_out/r2f_pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/r2f_pytest_gpu.log
grep -a "passed\|failed\|rel L2\|spp\|Error" gpurun_out/r2f_pytest_gpu.log | tail -12
for i in 1 2 3 4 5 6 7 8 9 10; do timeout 120 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "directional" 2>&1 | tail -1; done | sort | uniq -c
timeout 600 python bench.py --steps 24 --warmup 4 > gpurun_out/r2f_bench_n1.json 2> gpurun_out/r2f_bench_n1.err; echo "bench exit $?"; tail -3 gpurun_out/r2f_bench_n1.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2f_bench_n1.json').read().strip().splitlines()[-1])
print('N=1', d['value'], 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'], 'cpu', d['cpu_baseline']['value'], d['clocks'])
print('strong', d.get('strong_c5'))
PY
