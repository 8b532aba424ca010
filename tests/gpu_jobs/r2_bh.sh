#!/bin/bash
# round 2, call BH (8 GPUs): the driver's scaling command at N = 8 with the final schedule
mkdir -p gpurun_out; cd "$(dirname "$0")/../.."
nvidia-smi -L | wc -l
timeout -s KILL 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29708 bench.py --gpus 8 --steps 20 --warmup 3 --no-aux --no-stock --no-strong --no-modes --no-cpu-baseline --no-kernel-profile > gpurun_out/r2bh_bench_n8.json 2> gpurun_out/r2bh_bench_n8.err; echo "rc=$?"
python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/r2bh_bench_n8.json') if l.startswith('{')][-1])
    print({k:d[k] for k in ('value','ms_per_step','n_gpus')}); print('e2e', d['e2e']['value'], d['e2e']['ms_per_step'])
except Exception as e: print('parse failed', e)
PY
tail -2 gpurun_out/r2bh_bench_n8.err | cut -c1-300
