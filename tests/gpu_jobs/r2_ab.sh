#!/bin/bash
# round 2, call AB (2 GPUs): NCCL parity test, N = 2 bench with the all-reduces inside the graph (split Adam, early exchange), N = 1 on the same box
mkdir -p gpurun_out; cd "$(dirname "$0")/../.."
nvidia-smi -L
echo "== pytest nccl 2 ranks"; timeout -s KILL 420 python -m pytest tests/test_gpu_dist_nccl.py -m gpu -q -x -s -p no:cacheprovider > gpurun_out/r2ab_pytest_nccl.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/r2ab_pytest_nccl.log | cut -c1-400
for ng in 1 0; do echo "== bench N=2 nccl_in_graph=$ng"; TGB200_NCCL_GRAPH=$ng timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2961$ng bench.py --gpus 2 --no-aux --no-stock --no-strong --no-modes --no-cpu-baseline --no-kernel-profile > gpurun_out/r2ab_bench_n2_ng$ng.json 2> gpurun_out/r2ab_bench_n2_ng$ng.err; echo "rc=$?"; python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/r2ab_bench_n2_ng$ng.json') if l.startswith('{')][-1])
    print({k:d[k] for k in ('value','ms_per_step','n_gpus')}); print('e2e', d['e2e']['value'], d['e2e']['ms_per_step'])
except Exception as e: print('parse failed', e)
PY
done
echo "== bench N=1"; timeout -s KILL 300 python bench.py --no-aux --no-stock --no-strong --no-modes --no-cpu-baseline --no-kernel-profile > gpurun_out/r2ab_bench_n1.json 2> gpurun_out/r2ab_bench_n1.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2ab_bench_n1.json'))
print({k:d[k] for k in ('value','ms_per_step')})
PY
echo "== reference arm under torchrun N=2"; timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29633 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 2>/dev/null | tail -1 | cut -c1-300
