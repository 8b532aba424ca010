#!/bin/bash
# round 2, call S (1 GPU): clip-group two-tap GEMM (gemm_tcn.cu), A/B bench against the two-accumulator tiles
mkdir -p gpurun_out; cd "$(dirname "$0")/../.."
echo "== pytest two-tap"; timeout -s KILL 300 python -m pytest tests/test_gpu_tf32.py -m gpu -q -x -p no:cacheprovider -k "two_tap" 2>&1 | tail -6 | cut -c1-300
echo "== pytest all gpu"; timeout -s KILL 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2s_pytest_all.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/r2s_pytest_all.log | cut -c1-300
for old in 1 0; do echo "== bench two_acc=$old"; if [ $old = 1 ]; then export TGB200_TCN_TWO_ACC=1; else unset TGB200_TCN_TWO_ACC; fi
timeout -s KILL 600 python bench.py --no-aux --no-stock --no-strong --no-modes --no-cpu-baseline > gpurun_out/r2s_bench_old$old.json 2> gpurun_out/r2s_bench_old$old.err; echo "rc=$?"; tail -2 gpurun_out/r2s_bench_old$old.err | cut -c1-300; python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2s_bench_old$old.json'))
    print({k:d[k] for k in ('value','ms_per_step','launches_per_step')}); print('e2e', d['e2e']['value'], d['e2e']['ms_per_step'])
    print({k:(v['ms_per_step'], v.get('tflops')) for k,v in d['roofline']['families'].items()})
except Exception as e: print('parse failed', e)
PY
grep "taps2" gpurun_out/kernels_by_shape.txt | head -4
done
