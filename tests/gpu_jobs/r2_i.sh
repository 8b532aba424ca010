#!/bin/bash
# round 2, call I: K-split backward with bulk DSMEM copies + mbarriers; compile-time-unrolled MMA issue (H = 300); stand-alone encoders
mkdir -p gpurun_out; cd "$(dirname "$0")/../.."
echo "== trace"; timeout -s KILL 180 python tests/trace_gru.py > gpurun_out/r2i_trace.log 2>&1; echo "rc=$?"; grep -E "median|step 16|step  2" gpurun_out/r2i_trace.log
echo "== pytest gru tc"; timeout -s KILL 600 python -m pytest tests/test_gpu_tf32.py -m gpu -q -x -p no:cacheprovider > gpurun_out/r2i_pytest_tf32.log 2>&1; echo "rc=$?"; tail -12 gpurun_out/r2i_pytest_tf32.log
echo "== pytest all gpu"; timeout -s KILL 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2i_pytest_all.log 2>&1; echo "rc=$?"; tail -12 gpurun_out/r2i_pytest_all.log
echo "== bench"; timeout -s KILL 900 python bench.py --no-aux --no-stock --no-strong --no-modes --no-cpu-baseline > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err; echo "rc=$?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2i_bench.json'))
print({k:d[k] for k in ('value','ms_per_step')}); print('e2e', d['e2e']); print('fam', json.dumps(d['roofline']['families'])[:1500])
PY
cp gpurun_out/kernels_by_shape.txt gpurun_out/r2i_kernels_by_shape.txt
