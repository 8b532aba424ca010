#!/bin/bash
# round 2, call L (2 GPUs): NCCL parity test, NCCL inside the CUDA graph vs graph segments, small-GRU accumulator split
mkdir -p gpurun_out; cd "$(dirname "$0")/../.."
nvidia-smi -L
echo "== pytest small gru"; timeout -s KILL 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_tf32.py -m gpu -q -x -k gru -p no:cacheprovider 2>&1 | tail -3
echo "== pytest nccl 2 ranks"; timeout -s KILL 900 python -m pytest tests/test_gpu_dist_nccl.py -m gpu -q -x -s -p no:cacheprovider > gpurun_out/r2l_pytest_nccl.log 2>&1; echo "rc=$?"; tail -25 gpurun_out/r2l_pytest_nccl.log | cut -c1-600
for ng in 1 0; do echo "== bench N=2 nccl_in_graph=$ng"; TGB200_NCCL_GRAPH=$ng timeout -s KILL 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --no-aux --no-stock --no-strong --no-modes --no-cpu-baseline --no-kernel-profile > gpurun_out/r2l_bench_n2_ng$ng.json 2> gpurun_out/r2l_bench_n2_ng$ng.err; echo "rc=$?"; tail -3 gpurun_out/r2l_bench_n2_ng$ng.err | cut -c1-300; python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/r2l_bench_n2_ng$ng.json') if l.startswith('{')][-1])
    print({k:d[k] for k in ('value','ms_per_step','n_gpus')}); print('e2e', d['e2e']['value'], d['e2e']['ms_per_step'])
except Exception as e: print('parse failed', e)
PY
done
echo "== bench N=1"; timeout -s KILL 900 python bench.py --no-aux --no-stock --no-strong --no-modes --no-cpu-baseline > gpurun_out/r2l_bench_n1.json 2> gpurun_out/r2l_bench_n1.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2l_bench_n1.json'))
print({k:d[k] for k in ('value','ms_per_step')}); print({k:v['ms_per_step'] for k,v in d['roofline']['families'].items()})
PY
