#!/bin/sh
# closing run of a round on one GPU: all GPU tests, the bench, the reference arm
TAG=${1:-r2final}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/${TAG}_pytest_gpu.log
tail -3 gpurun_out/${TAG}_pytest_gpu.log
timeout 400 python bench.py --steps 24 --warmup 4 > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err
python -c "
import json
d=json.loads(open('gpurun_out/${TAG}_bench_n1.json').read().strip().splitlines()[-1]); print('N=1', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['cpu_baseline']['value'], d['clocks'], d['strong_c5']['value'])"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_reference_arm.json 2> gpurun_out/${TAG}_bench_reference_arm.err
tail -c 300 gpurun_out/${TAG}_bench_reference_arm.json
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
