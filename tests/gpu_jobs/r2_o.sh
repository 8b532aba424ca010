#!/bin/bash
# round 2, call O (1 GPU): fused ConvDiscriminator kernels v2 (cp.async staging, no global loads in the step loops), float4 BatchNorm kernels
mkdir -p gpurun_out; cd "$(dirname "$0")/../.."
echo "== pytest fused D"; timeout -s KILL 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -k "fused_stack" 2>&1 | tail -15 | cut -c1-400
echo "== pytest all gpu"; timeout -s KILL 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2o_pytest_all.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/r2o_pytest_all.log | cut -c1-300
for df in 1; do echo "== bench d_fused=$df"; TGB200_D_FUSED=$df timeout -s KILL 600 python bench.py --no-aux --no-stock --no-strong --no-modes --no-cpu-baseline --no-kernel-profile > gpurun_out/r2o_bench_df$df.json 2> gpurun_out/r2o_bench_df$df.err; echo "rc=$?"; tail -2 gpurun_out/r2o_bench_df$df.err | cut -c1-300; python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2o_bench_df$df.json'))
    print({k:d[k] for k in ('value','ms_per_step','launches_per_step')}); print('e2e', d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['per_step_ms'])
except Exception as e: print('parse failed', e)
PY
done
echo "== timeline"; timeout -s KILL 300 python tests/timeline_step.py gpurun_out/r2o_timeline.csv > gpurun_out/r2o_timeline.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/r2o_timeline.log
python tests/timeline_to_txt.py gpurun_out/r2o_timeline.json gpurun_out/r2o_timeline_step.txt && python tests/analyze_timeline.py gpurun_out/r2o_timeline_step.txt > gpurun_out/r2o_timeline_step_ownership.txt; head -45 gpurun_out/r2o_timeline_step_ownership.txt
rm -f gpurun_out/r2o_timeline.json
grep -n "dgru_stack\|dconv_stack\|bn_bwd\|col_reduce\|affine_lrelu" gpurun_out/r2o_timeline_step.txt | cut -c1-120
