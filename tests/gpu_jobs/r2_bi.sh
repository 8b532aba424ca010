#!/bin/bash
# round 2, call BI (1 GPU): TemporalBlock residual add + final ReLU in conv2's GEMM epilogue, fused backward head (tg_tcn_res_bwd)
mkdir -p gpurun_out; cd "$(dirname "$0")/../.."
B="--no-aux --no-stock --no-strong --no-modes --no-cpu-baseline --no-kernel-profile"
echo "== pytest"; timeout -s KILL 300 python -m pytest tests/test_gpu_tf32.py tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "two_tap or graph_replay or golden or train_iter or fp64 or full_size" 2>&1 | tail -3 | cut -c1-300
for name in fused unfused; do
  if [ $name = unfused ]; then export TGB200_TCN_FUSED_ADD=0; fi
  timeout -s KILL 200 python bench.py $B > gpurun_out/r2bi_bench_$name.json 2> gpurun_out/r2bi_bench_$name.err; echo "rc=$?"
  python - <<PY
import json
d=json.load(open('gpurun_out/r2bi_bench_$name.json'))
print('$name', {k:d[k] for k in ('value','ms_per_step','launches_per_step')}, 'e2e', d['e2e']['value'])
PY
done
