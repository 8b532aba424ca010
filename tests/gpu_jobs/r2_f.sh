#!/bin/bash
# round 2, call F: accumulator rotation A/B in the forward recurrence, prefetching discriminator GRU, the new bench.py end to end
mkdir -p gpurun_out; cd "$(dirname "$0")/../.."
for n in 1 2 3; do echo "== trace nacc=$n"; TGB200_GRU_NACC=$n timeout -s KILL 180 python tests/trace_gru.py > gpurun_out/r2f_trace_nacc$n.log 2>&1; echo "rc=$?"; grep -E "median|step 16" gpurun_out/r2f_trace_nacc$n.log | head -4; done
echo "== nacc=3 correctness"; TGB200_GRU_NACC=3 timeout -s KILL 300 python -m pytest tests/test_gpu_tf32.py -m gpu -q -x -k gru_layer_tensor_core -p no:cacheprovider 2>&1 | tail -2
echo "== pytest all gpu"; timeout -s KILL 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2f_pytest_all.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/r2f_pytest_all.log
echo "== bench"; timeout -s KILL 900 python bench.py > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; echo "rc=$?"; grep -E "Elapsed|Error|error" gpurun_out/r2f_bench.err | head; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2f_bench.json'))
print({k:d[k] for k in ('value','ms_per_step')}); print('e2e', d['e2e']); print('modes', d['modes']); print('strong', d['strong_scaling']); print('stock', d['gpu_stock_baseline'])
print('aux', d['aux']); print('cpu', d['cpu_baseline']); print('fam', json.dumps(d['roofline']['families'])[:1500])
PY
cp gpurun_out/kernels_by_shape.txt gpurun_out/r2f_kernels_by_shape.txt
echo "== bench reference arm"; timeout -s KILL 900 python bench.py --impl reference --steps 5 --warmup 1 2>&1 | grep -E "impl|Elapsed" | cut -c1-400
