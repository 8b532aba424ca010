#!/bin/bash
# round 2, call AN (1 GPU): 240-column GEMM tiles for the N = 1800 projections; CUDA_DEVICE_MAX_CONNECTIONS A/B; timeline
mkdir -p gpurun_out; cd "$(dirname "$0")/../.."
B="--no-aux --no-stock --no-strong --no-modes --no-cpu-baseline --no-kernel-profile"
echo "== pytest gemm"; timeout -s KILL 300 python -m pytest tests/test_gpu_tf32.py -m gpu -q -x -p no:cacheprovider -k "gemm_tf32_plain" 2>&1 | tail -4 | cut -c1-300
echo "== prof_gemm bn240"; timeout -s KILL 120 python tests/prof_gemm.py 2>&1 | tail -3
echo "== prof_gemm bn128"; TGB200_NO_BN240=1 timeout -s KILL 120 python tests/prof_gemm.py 2>&1 | tail -3
run() { # name, env...
  local name=$1; shift
  env "$@" timeout -s KILL 300 python bench.py $B > gpurun_out/r2an_bench_$name.json 2> gpurun_out/r2an_bench_$name.err; echo "rc=$?"
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2an_bench_$name.json'))
    print('$name', {k:d[k] for k in ('value','ms_per_step')}, 'e2e', d['e2e']['value'])
except Exception as e: print('parse failed', e)
PY
}
echo "== bench"
run default X=1
run conn32 CUDA_DEVICE_MAX_CONNECTIONS=32
run conn32_bn128 CUDA_DEVICE_MAX_CONNECTIONS=32 TGB200_NO_BN240=1
run bn128 TGB200_NO_BN240=1
echo "== timeline conn32"; CUDA_DEVICE_MAX_CONNECTIONS=32 timeout -s KILL 300 python tests/timeline_step.py gpurun_out/r2an_timeline.csv > gpurun_out/r2an_timeline.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/r2an_timeline.log
python tests/timeline_to_txt.py gpurun_out/r2an_timeline.json gpurun_out/r2an_timeline_step.txt && python tests/analyze_timeline.py gpurun_out/r2an_timeline_step.txt > gpurun_out/r2an_timeline_step_ownership.txt; head -12 gpurun_out/r2an_timeline_step_ownership.txt
rm -f gpurun_out/r2an_timeline.json
