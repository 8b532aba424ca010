#!/bin/bash
# round 2, call AA (1 GPU): TCN GEMM with 16 epilogue warps, dropout fused into the GRU forward; trace, tests, bench, timeline
mkdir -p gpurun_out; cd "$(dirname "$0")/../.."
echo "== pytest kernels"; timeout -s KILL 300 python -m pytest tests/test_gpu_tf32.py -m gpu -q -x -p no:cacheprovider -k "conv1_wgrad or two_tap" 2>&1 | tail -4 | cut -c1-300
echo "== pytest all gpu"; timeout -s KILL 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2aa_pytest_all.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/r2aa_pytest_all.log | cut -c1-300
echo "== bench"; timeout -s KILL 600 python bench.py --no-aux --no-stock --no-strong --no-modes --no-cpu-baseline --no-kernel-profile > gpurun_out/r2aa_bench.json 2> gpurun_out/r2aa_bench.err; echo "rc=$?"; tail -2 gpurun_out/r2aa_bench.err | cut -c1-300; python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2aa_bench.json'))
    print({k:d[k] for k in ('value','ms_per_step','launches_per_step')}); print('e2e', d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['per_step_ms'])
except Exception as e: print('parse failed', e)
PY
echo "== timeline"; timeout -s KILL 300 python tests/timeline_step.py gpurun_out/r2aa_timeline.csv > gpurun_out/r2aa_timeline.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/r2aa_timeline.log
python tests/timeline_to_txt.py gpurun_out/r2aa_timeline.json gpurun_out/r2aa_timeline_step.txt && python tests/analyze_timeline.py gpurun_out/r2aa_timeline_step.txt > gpurun_out/r2aa_timeline_step_ownership.txt; head -36 gpurun_out/r2aa_timeline_step_ownership.txt
rm -f gpurun_out/r2aa_timeline.json
