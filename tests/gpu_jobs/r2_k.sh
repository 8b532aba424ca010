#!/bin/bash
# round 2, call K: nearest-TF32 operands inside the recurrence, text chain captured before the audio chain, early scalar read-back
mkdir -p gpurun_out; cd "$(dirname "$0")/../.."
echo "== pytest all gpu"; timeout -s KILL 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2k_pytest_all.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/r2k_pytest_all.log
echo "== fast-mode errors"; timeout -s KILL 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -k "fast_mode" -s 2>&1 | grep -E "rel|worst|err|passed|failed" | head -20
for wf in 0 1; do echo "== bench wav_first=$wf"; TGB200_WAV_FIRST=$wf timeout -s KILL 900 python bench.py --no-aux --no-stock --no-strong --no-modes --no-cpu-baseline --no-kernel-profile > gpurun_out/r2k_bench_wf$wf.json 2> gpurun_out/r2k_bench_wf$wf.err; echo "rc=$?"; python - <<PY
import json
d=json.load(open('gpurun_out/r2k_bench_wf$wf.json'))
print({k:d[k] for k in ('value','ms_per_step')}); print('e2e', d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['per_step_ms'])
PY
done
