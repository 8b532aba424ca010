#!/bin/bash
# round 2, call AL (1 GPU): full suite incl. Speech2Gesture; smoke; bench aux with the S2G leg
mkdir -p gpurun_out; cd "$(dirname "$0")/../.."
echo "== pytest all gpu"; timeout -s KILL 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2al_pytest_all.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/r2al_pytest_all.log | cut -c1-300
echo "== smoke"; timeout -s KILL 300 python __graft_entry__.py smoke 2>&1 | grep -v Warn | tail -3
echo "== bench"; timeout -s KILL 600 python bench.py --no-stock --no-strong --no-modes --no-cpu-baseline --no-kernel-profile > gpurun_out/r2al_bench.json 2> gpurun_out/r2al_bench.err; echo "rc=$?"; tail -2 gpurun_out/r2al_bench.err | cut -c1-300; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2al_bench.json'))
print({k:d[k] for k in ('value','ms_per_step','launches_per_step')}); print('e2e', d['e2e']['value'], d['e2e']['ms_per_step']); print('aux', d['aux'])
PY
