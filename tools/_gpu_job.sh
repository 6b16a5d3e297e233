python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 16 --warmup 3 > gpurun_out/r01g_bench_n2.json 2> gpurun_out/r01g_bench_n2.err
python -c "
import json
d=json.loads(open('gpurun_out/r01g_bench_n2.json').read().strip().splitlines()[-1]); print('N=2', d['value'], d['e2e']['value'], d['finite'])"
tail -2 gpurun_out/r01g_bench_n2.err | cut -c1-200
python bench.py --res 3840 2160 --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r01g_bench_4k_n1.json 2> gpurun_out/r01g_bench_4k_n1.err
python -c "
import json
d=json.loads(open('gpurun_out/r01g_bench_4k_n1.json').read().strip().splitlines()[-1]); print('4K N=1', d['value'], d['e2e']['value'], d['ms_per_step'], d['finite'])"
timeout 200 python tools/run_configs.py --only C1_gpu > gpurun_out/configs_r01_c1gpu.json 2> gpurun_out/configs_r01_c1gpu.err
grep -o '"Msamples_per_s_device": [0-9.]*\|"rel_l2_composited": [0-9.e-]*' gpurun_out/configs_r01_c1gpu.err
