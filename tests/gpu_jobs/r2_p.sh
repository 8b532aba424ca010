#!/bin/bash
# round 2, call P (1 GPU): fused D v2 + alignment fix + fast intrinsics; ncu of v2; full suite; bench
mkdir -p gpurun_out; cd "$(dirname "$0")/../.."
echo "== pytest fused D"; timeout -s KILL 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -k "fused_stack" 2>&1 | tail -8 | cut -c1-400
echo "== fused D in fp32 mode"; TGB200_MODE=fp32 timeout -s KILL 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -k "fused_stack" 2>&1 | tail -3 | cut -c1-400
echo "== pytest all gpu"; timeout -s KILL 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2p_pytest_all.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/r2p_pytest_all.log | cut -c1-300
echo "== bench"; timeout -s KILL 600 python bench.py --no-aux --no-stock --no-strong --no-modes --no-cpu-baseline --no-kernel-profile > gpurun_out/r2p_bench.json 2> gpurun_out/r2p_bench.err; echo "rc=$?"; tail -2 gpurun_out/r2p_bench.err | cut -c1-300; python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2p_bench.json'))
    print({k:d[k] for k in ('value','ms_per_step','launches_per_step')}); print('e2e', d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['per_step_ms'])
except Exception as e: print('parse failed', e)
PY
echo "== ncu"; timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:"dgru_stack|dconv_stack" -c 4 -o gpurun_out/r2p_ncu_dfused -f python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "fused_stack and 128" > gpurun_out/r2p_ncu.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/r2p_ncu.log | cut -c1-200
