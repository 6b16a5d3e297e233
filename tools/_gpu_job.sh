#!/bin/sh
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 4 --steps 16 --warmup 4 > gpurun_out/r03_bench_n4.json 2> gpurun_out/r03_bench_n4.err
python -c "
import json
d=json.loads(open('gpurun_out/r03_bench_n4.json').read().strip().splitlines()[-1]); print('N=4', d['value'], d['e2e']['value'], d['ms_per_step'], d['finite'], d['config']['workload'])"
tail -2 gpurun_out/r03_bench_n4.err | cut -c1-200
