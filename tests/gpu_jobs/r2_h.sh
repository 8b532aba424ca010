#!/bin/bash
# round 2, call H: MMA microbenchmark (fixed), K-split backward recurrence kernel
mkdir -p gpurun_out; cd "$(dirname "$0")/../.."
echo "== mma ubench"; timeout -s KILL 120 tests/ubench/bin/mma_issue > gpurun_out/r2h_mma_issue.txt 2>&1; echo "rc=$?"; cat gpurun_out/r2h_mma_issue.txt
echo "== trace (bwd K-split)"; timeout -s KILL 180 python tests/trace_gru.py > gpurun_out/r2h_trace.log 2>&1; echo "rc=$?"; grep -E "median|step 16|step  2" gpurun_out/r2h_trace.log
echo "== pytest gru tc"; timeout -s KILL 600 python -m pytest tests/test_gpu_tf32.py -m gpu -q -x -p no:cacheprovider > gpurun_out/r2h_pytest_tf32.log 2>&1; echo "rc=$?"; tail -12 gpurun_out/r2h_pytest_tf32.log
echo "== pytest all gpu"; timeout -s KILL 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2h_pytest_all.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/r2h_pytest_all.log
echo "== bench"; timeout -s KILL 900 python bench.py --no-aux --no-stock --no-strong --no-modes --no-cpu-baseline > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err; echo "rc=$?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2h_bench.json'))
print({k:d[k] for k in ('value','ms_per_step')}); print('e2e', d['e2e']); print('fam', json.dumps(d['roofline']['families'])[:1500])
PY
