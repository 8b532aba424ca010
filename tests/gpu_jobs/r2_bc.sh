#!/bin/bash
# round 2, call BC (1 GPU): direction duplication folded into the head data-gradient GEMM
mkdir -p gpurun_out; cd "$(dirname "$0")/../.."
B="--no-aux --no-stock --no-strong --no-modes --no-cpu-baseline --no-kernel-profile"
echo "== pytest"; timeout -s KILL 600 python -m pytest tests/test_gpu_tf32.py tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "padded_pitch or gemm_tf32_plain or graph_replay or golden or train_iter or fp64 or full_size" 2>&1 | tail -4 | cut -c1-300
run() { # name, env...
  local name=$1; shift
  env "$@" timeout -s KILL 300 python bench.py $B > gpurun_out/r2bc_bench_$name.json 2> gpurun_out/r2bc_bench_$name.err; echo "rc=$?"
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2bc_bench_$name.json'))
    print('$name', {k:d[k] for k in ('value','ms_per_step')}, 'e2e', d['e2e']['value'])
except Exception as e: print('parse failed', e)
PY
}
echo "== bench"
run default X=1
run default2 X=1
echo "== timeline"; timeout -s KILL 300 python tests/timeline_step.py gpurun_out/r2bc_timeline.csv > gpurun_out/r2bc_timeline.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/r2bc_timeline.log
python tests/timeline_to_txt.py gpurun_out/r2bc_timeline.json gpurun_out/r2bc_timeline_step.txt && python tests/analyze_timeline.py gpurun_out/r2bc_timeline_step.txt > gpurun_out/r2bc_timeline_step_ownership.txt; head -12 gpurun_out/r2bc_timeline_step_ownership.txt
rm -f gpurun_out/r2bc_timeline.json
