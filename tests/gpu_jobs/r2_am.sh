#!/bin/bash
# round 2, call AM (8 GPUs): the driver's scaling command at N = 8 and N = 4 (NCCL all-reduces inside the iteration graph)
mkdir -p gpurun_out; cd "$(dirname "$0")/../.."
nvidia-smi -L | wc -l
for n in 8 4; do
echo "== bench N=$n"; timeout -s KILL 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2970$n bench.py --gpus $n --steps 20 --warmup 3 --no-aux --no-stock --no-strong --no-modes --no-cpu-baseline --no-kernel-profile > gpurun_out/r2am_bench_n$n.json 2> gpurun_out/r2am_bench_n$n.err; echo "rc=$?"; python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/r2am_bench_n$n.json') if l.startswith('{')][-1])
    print({k:d[k] for k in ('value','ms_per_step','n_gpus')}); print('e2e', d['e2e']['value'], d['e2e']['ms_per_step'])
except Exception as e: print('parse failed', e)
PY
tail -2 gpurun_out/r2am_bench_n$n.err | cut -c1-300
done
echo "== bench N=1"; timeout -s KILL 200 python bench.py --no-aux --no-stock --no-strong --no-modes --no-cpu-baseline --no-kernel-profile > gpurun_out/r2am_bench_n1.json 2> gpurun_out/r2am_bench_n1.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2am_bench_n1.json'))
print({k:d[k] for k in ('value','ms_per_step')})
PY
