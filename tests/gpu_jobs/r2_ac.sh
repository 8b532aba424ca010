#!/bin/bash
# round 2, call AC (2 GPUs): full suite, N = 1 bench, NCCL test, N = 2 bench (pipelined tail exchange with per-bucket Adam, batched BN reductions)
mkdir -p gpurun_out; cd "$(dirname "$0")/../.."
echo "== pytest all gpu"; timeout -s KILL 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2ac_pytest_all.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/r2ac_pytest_all.log | cut -c1-300
echo "== bench N=1"; timeout -s KILL 300 python bench.py --no-aux --no-stock --no-strong --no-modes --no-cpu-baseline --no-kernel-profile > gpurun_out/r2ac_bench_n1.json 2> gpurun_out/r2ac_bench_n1.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2ac_bench_n1.json'))
print({k:d[k] for k in ('value','ms_per_step')}, 'e2e', d['e2e']['ms_per_step'])
PY
echo "== bench N=2"; timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29615 bench.py --gpus 2 --no-aux --no-stock --no-strong --no-modes --no-cpu-baseline --no-kernel-profile > gpurun_out/r2ac_bench_n2.json 2> gpurun_out/r2ac_bench_n2.err; echo "rc=$?"; python - <<'PY'
import json
try:
    d=json.loads([l for l in open('gpurun_out/r2ac_bench_n2.json') if l.startswith('{')][-1])
    print({k:d[k] for k in ('value','ms_per_step','n_gpus')}); print('e2e', d['e2e']['value'], d['e2e']['ms_per_step'])
except Exception as e: print('parse failed', e)
PY
tail -3 gpurun_out/r2ac_bench_n2.err | cut -c1-300
grep -n "col_reduce_v4" gpurun_out/r2ac_timeline_step.txt 2>/dev/null | head -3
