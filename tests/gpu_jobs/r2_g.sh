#!/bin/bash
# round 2, call G: tcgen05.mma issue / execution microbenchmark; the new bench.py end to end
mkdir -p gpurun_out; cd "$(dirname "$0")/../.."
echo "== mma ubench"; timeout -s KILL 120 tests/ubench/bin/mma_issue > gpurun_out/r2g_mma_issue.txt 2>&1; echo "rc=$?"; cat gpurun_out/r2g_mma_issue.txt
echo "== trace"; timeout -s KILL 180 python tests/trace_gru.py > gpurun_out/r2g_trace.log 2>&1; echo "rc=$?"; grep -E "median|step 16" gpurun_out/r2g_trace.log
echo "== bench"; time (timeout -s KILL 900 python bench.py > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err); echo "rc=$?"; grep -E "Error|error" gpurun_out/r2g_bench.err | head; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2g_bench.json'))
print({k:d[k] for k in ('value','ms_per_step')}); print('e2e', d['e2e']); print('modes', d['modes']); print('strong', d['strong_scaling']); print('stock', d['gpu_stock_baseline'])
print('aux', d['aux']); print('cpu', d['cpu_baseline']); print('fam', json.dumps(d['roofline']['families'])[:1800])
PY
cp gpurun_out/kernels_by_shape.txt gpurun_out/r2g_kernels_by_shape.txt
echo "== bench reference arm"; time (timeout -s KILL 900 python bench.py --impl reference --steps 5 --warmup 1 2>&1 | grep -E "impl" | cut -c1-400)
