#!/bin/bash
# round 2, call AS (1 GPU): two bias streams, GRU masks drawn at the GRU input; full GPU suite; bench
mkdir -p gpurun_out; cd "$(dirname "$0")/../.."
B="--no-aux --no-stock --no-strong --no-modes --no-cpu-baseline --no-kernel-profile"
run() { # name, env...
  local name=$1; shift
  env "$@" timeout -s KILL 300 python bench.py $B > gpurun_out/r2as_bench_$name.json 2> gpurun_out/r2as_bench_$name.err; echo "rc=$?"
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2as_bench_$name.json'))
    print('$name', {k:d[k] for k in ('value','ms_per_step')}, 'e2e', d['e2e']['value'], d['e2e'].get('per_step_ms'))
except Exception as e: print('parse failed', e)
PY
}
echo "== bench"
run default X=1
echo "== pytest all gpu"; timeout -s KILL 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2as_pytest_all.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/r2as_pytest_all.log | cut -c1-300
run default2 X=1
echo "== timeline"; timeout -s KILL 300 python tests/timeline_step.py gpurun_out/r2as_timeline.csv > gpurun_out/r2as_timeline.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/r2as_timeline.log
python tests/timeline_to_txt.py gpurun_out/r2as_timeline.json gpurun_out/r2as_timeline_step.txt && python tests/analyze_timeline.py gpurun_out/r2as_timeline_step.txt > gpurun_out/r2as_timeline_step_ownership.txt; head -12 gpurun_out/r2as_timeline_step_ownership.txt
rm -f gpurun_out/r2as_timeline.json
