#!/bin/bash
# round 2, call J: re-run the two failed tests, step timeline + ownership, ncu --set full of the recurrence kernels, full bench
mkdir -p gpurun_out; cd "$(dirname "$0")/../.."
echo "== pytest failed ones"; timeout -s KILL 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_zz_autoencoder.py -m gpu -q -p no:cacheprovider -k "standalone or batch128_steps" 2>&1 | tail -5
echo "== timeline"; timeout -s KILL 300 python tests/timeline_step.py gpurun_out/r2j_timeline.csv > gpurun_out/r2j_timeline.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/r2j_timeline.log
python tests/timeline_to_txt.py gpurun_out/r2j_timeline.json gpurun_out/r2j_timeline_step.txt && python tests/analyze_timeline.py gpurun_out/r2j_timeline_step.txt > gpurun_out/r2j_timeline_step_ownership.txt; head -30 gpurun_out/r2j_timeline_step_ownership.txt
rm -f gpurun_out/r2j_timeline.json
echo "== ncu gru fwd"; timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:gru_fwd_cl_kernel -s 3 -c 1 -o gpurun_out/r2j_ncu_gru_fwd -f python tests/trace_gru.py > gpurun_out/r2j_ncu_fwd.log 2>&1; echo "rc=$?"
echo "== ncu gru bwd"; timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:gru_bwd_ks_kernel -s 3 -c 1 -o gpurun_out/r2j_ncu_gru_bwd -f python tests/trace_gru.py > gpurun_out/r2j_ncu_bwd.log 2>&1; echo "rc=$?"
ls -la gpurun_out/*.ncu-rep
echo "== bench"; time (timeout -s KILL 900 python bench.py > gpurun_out/r2j_bench.json 2> gpurun_out/r2j_bench.err); echo "rc=$?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2j_bench.json'))
print({k:d[k] for k in ('value','ms_per_step')}); print('e2e', d['e2e']); print('modes', d['modes']); print('strong', d['strong_scaling']); print('stock', {k:v for k,v in d['gpu_stock_baseline'].items() if k!='what'})
PY
cp gpurun_out/kernels_by_shape.txt gpurun_out/r2j_kernels_by_shape.txt
