#!/bin/bash
# round 2, call AU (2 GPUs): NCCL parity test and N = 2 / N = 1 bench after the schedule changes (priorities, preparation streams, D(real) beside the recurrence)
mkdir -p gpurun_out; cd "$(dirname "$0")/../.."
nvidia-smi -L | wc -l
B="--no-aux --no-stock --no-strong --no-modes --no-cpu-baseline --no-kernel-profile"
echo "== pytest nccl 2 ranks"; timeout -s KILL 420 python -m pytest tests/test_gpu_dist_nccl.py -m gpu -q -x -s -p no:cacheprovider > gpurun_out/r2au_pytest_nccl.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/r2au_pytest_nccl.log | cut -c1-400
echo "== bench N=2"; timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 20 --warmup 3 $B > gpurun_out/r2au_bench_n2.json 2> gpurun_out/r2au_bench_n2.err; echo "rc=$?"; python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/r2au_bench_n2.json') if l.startswith('{')][-1])
    print({k:d[k] for k in ('value','ms_per_step','n_gpus')}); print('e2e', d['e2e']['value'], d['e2e']['ms_per_step'])
except Exception as e: print('parse failed', e)
PY
tail -2 gpurun_out/r2au_bench_n2.err | cut -c1-300
echo "== bench N=1"; timeout -s KILL 300 python bench.py $B > gpurun_out/r2au_bench_n1.json 2> gpurun_out/r2au_bench_n1.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2au_bench_n1.json'))
print({k:d[k] for k in ('value','ms_per_step')}, 'e2e', d['e2e']['value'])
PY
echo "== parity subset"; timeout -s KILL 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "graph_replay or golden or train_iter" 2>&1 | tail -2 | cut -c1-300
