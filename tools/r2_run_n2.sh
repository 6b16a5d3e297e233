#!/bin/sh
# N-GPU checks of the product's frame gather: parity vs the unsharded render, bench (weak + strong_c5), CLI -gpus N
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR tools/multigpu_check.py --res 1600 900 2> gpurun_out/mg_check.err | tail -1
[ "$N" -le 2 ] && timeout 300 $TR tools/multigpu_check.py --res 1000 700 --passes 3 2>> gpurun_out/mg_check.err | tail -1
tail -3 gpurun_out/mg_check.err
timeout 600 $TR bench.py --gpus $N --steps 24 --warmup 4 > gpurun_out/r2h_bench_n$N.json 2> gpurun_out/r2h_bench_n$N.err; echo "bench exit $?"; tail -2 gpurun_out/r2h_bench_n$N.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2h_bench_n$N.json').read().strip().splitlines()[-1])
print('N=$N', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], d['clocks'])
print('strong', d.get('strong_c5'))
PY
python -c "import fermat_b200 as fb, shutil; shutil.copy(fb.resolve_scene('scenes/_cache/bathroom2.fbs'), '/tmp/bathroom2.fbs')"
cd gpurun_out
timeout 300 ../fermat_b200/fermat_pt -pt -i /tmp/bathroom2.fbs -r 800 450 -bounces 8 -passes 3 -o cli_n1 2>&1 | tail -1
timeout 300 ../fermat_b200/fermat_pt -pt -i /tmp/bathroom2.fbs -r 800 450 -bounces 8 -passes 3 -o cli_nN -gpus $N 2>&1 | tail -2
cmp cli_n1.pfm cli_nN.pfm && echo "CLI: -gpus $N image identical to the 1-GPU image (pfm)"; cmp cli_n1.tga cli_nN.tga && echo "CLI: tga identical"
rm -f cli_n1.pfm cli_nN.pfm cli_nN.tga; mv cli_n1.tga r2h_cli_bathroom2_800x450.tga
