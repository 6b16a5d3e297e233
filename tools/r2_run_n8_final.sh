#!/bin/sh
# final N-GPU numbers of a round: parity of the gathered frame + bench (weak + strong_c5)
N=${1:-8}
TAG=${2:-r2u}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR tools/multigpu_check.py --res 1600 900 2> gpurun_out/mg_check.err | tail -1 | tee gpurun_out/${TAG}_multigpu_check_n$N.json
timeout 600 $TR bench.py --gpus $N --steps 24 --warmup 4 > gpurun_out/${TAG}_bench_n$N.json 2> gpurun_out/${TAG}_bench_n$N.err; echo "bench exit $?"; tail -2 gpurun_out/${TAG}_bench_n$N.err
python - <<PY
import json
d=json.loads(open('gpurun_out/${TAG}_bench_n$N.json').read().strip().splitlines()[-1])
print('N=$N', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], d['clocks'])
print('strong', d.get('strong_c5'))
PY
