#!/bin/bash
# round 2, call R (1 GPU): fused D v3 (mma.sync per-clip GEMMs in fast mode), full-vocabulary parity cases, bench, timeline
mkdir -p gpurun_out; cd "$(dirname "$0")/../.."
echo "== pytest fused D"; timeout -s KILL 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -p no:cacheprovider -k "fused_stack" 2>&1 | grep -E "fused vs|passed|failed|Error|assert" | head -12 | cut -c1-300
echo "== fused D in fp32 mode"; TGB200_MODE=fp32 timeout -s KILL 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -p no:cacheprovider -k "fused_stack" 2>&1 | grep -E "fused vs|passed|failed|Error|assert" | head -12 | cut -c1-300
echo "== pytest all gpu"; timeout -s KILL 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2r_pytest_all.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/r2r_pytest_all.log | cut -c1-300
echo "== fast-mode errors"; timeout -s KILL 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -k "fast_mode_train" -s 2>&1 | grep -E "rel-L2|passed|failed" | head -20 | cut -c1-600
echo "== bench"; timeout -s KILL 600 python bench.py --no-aux --no-stock --no-strong --no-modes --no-cpu-baseline --no-kernel-profile > gpurun_out/r2r_bench.json 2> gpurun_out/r2r_bench.err; echo "rc=$?"; tail -2 gpurun_out/r2r_bench.err | cut -c1-300; python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2r_bench.json'))
    print({k:d[k] for k in ('value','ms_per_step','launches_per_step')}); print('e2e', d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['per_step_ms'])
except Exception as e: print('parse failed', e)
PY
echo "== timeline"; timeout -s KILL 300 python tests/timeline_step.py gpurun_out/r2r_timeline.csv > gpurun_out/r2r_timeline.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/r2r_timeline.log
python tests/timeline_to_txt.py gpurun_out/r2r_timeline.json gpurun_out/r2r_timeline_step.txt && python tests/analyze_timeline.py gpurun_out/r2r_timeline_step.txt > gpurun_out/r2r_timeline_step_ownership.txt; head -30 gpurun_out/r2r_timeline_step_ownership.txt
rm -f gpurun_out/r2r_timeline.json
grep -n "dgru_stack\|dconv_stack" gpurun_out/r2r_timeline_step.txt | cut -c1-120
