#!/bin/bash
# round 2, call AO (1 GPU): per-block preparation streams (TCN masks / weight norm), D(real) pass forked beside the generator's recurrence
mkdir -p gpurun_out; cd "$(dirname "$0")/../.."
B="--no-aux --no-stock --no-strong --no-modes --no-cpu-baseline --no-kernel-profile"
echo "== pytest parity subset"; timeout -s KILL 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "graph_replay or golden or train_iter" 2>&1 | tail -4 | cut -c1-300
run() { # name, env...
  local name=$1; shift
  env "$@" timeout -s KILL 300 python bench.py $B > gpurun_out/r2ao_bench_$name.json 2> gpurun_out/r2ao_bench_$name.err; echo "rc=$?"
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2ao_bench_$name.json'))
    print('$name', {k:d[k] for k in ('value','ms_per_step')}, 'e2e', d['e2e']['value'])
except Exception as e: print('parse failed', e)
PY
}
echo "== bench"
run top TGB200_DREAL_AT=top
run concat TGB200_DREAL_AT=concat
run gru0 TGB200_DREAL_AT=gru0
run gru1 TGB200_DREAL_AT=gru1
best=$(python - <<'PY'
import json
b=None
for n in ('top','concat','gru0','gru1'):
    try:
        d=json.load(open('gpurun_out/r2ao_bench_%s.json'%n))
        if b is None or d['ms_per_step']<b[1]: b=(n,d['ms_per_step'])
    except Exception: pass
print(b[0] if b else 'top')
PY
)
echo "== timeline $best"; TGB200_DREAL_AT=$best timeout -s KILL 300 python tests/timeline_step.py gpurun_out/r2ao_timeline.csv > gpurun_out/r2ao_timeline.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/r2ao_timeline.log
python tests/timeline_to_txt.py gpurun_out/r2ao_timeline.json gpurun_out/r2ao_timeline_step.txt && python tests/analyze_timeline.py gpurun_out/r2ao_timeline_step.txt > gpurun_out/r2ao_timeline_step_ownership.txt; head -12 gpurun_out/r2ao_timeline_step_ownership.txt
rm -f gpurun_out/r2ao_timeline.json
