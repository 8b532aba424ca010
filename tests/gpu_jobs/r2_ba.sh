#!/bin/bash
# round 2, call BA (2 GPUs): final code - NCCL test, N = 2 bench under torchrun, N = 1 on the same box, ncu --set full of the wide projection GEMM
mkdir -p gpurun_out; cd "$(dirname "$0")/../.."
nvidia-smi -L | wc -l
B="--no-aux --no-stock --no-strong --no-modes --no-cpu-baseline --no-kernel-profile"
echo "== pytest nccl 2 ranks"; timeout -s KILL 420 python -m pytest tests/test_gpu_dist_nccl.py -m gpu -q -x -s -p no:cacheprovider > gpurun_out/r2ba_pytest_nccl.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/r2ba_pytest_nccl.log | cut -c1-400
echo "== bench N=2"; timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 20 --warmup 3 $B > gpurun_out/r2ba_bench_n2.json 2> gpurun_out/r2ba_bench_n2.err; echo "rc=$?"; python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/r2ba_bench_n2.json') if l.startswith('{')][-1])
    print({k:d[k] for k in ('value','ms_per_step','n_gpus')}); print('e2e', d['e2e']['value'], d['e2e']['ms_per_step'])
except Exception as e: print('parse failed', e)
PY
echo "== bench N=1"; timeout -s KILL 300 python bench.py $B > gpurun_out/r2ba_bench_n1.json 2> gpurun_out/r2ba_bench_n1.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2ba_bench_n1.json'))
print({k:d[k] for k in ('value','ms_per_step')}, 'e2e', d['e2e']['value'])
PY
echo "== ncu full gemm (forward side)"; CUDA_VISIBLE_DEVICES=0 timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tf32_kernel -s 3 -c 8 -o gpurun_out/r2ba_ncu_gemm_fwd -f python bench.py --steps 2 --warmup 3 $B > gpurun_out/r2ba_ncu_gemm_fwd.log 2>&1; echo "rc=$?"
ls -la gpurun_out/r2ba_ncu_*.ncu-rep
