#!/bin/bash
# round 2, call AZ (1 GPU): evidence run with the final code - full GPU suite, full default bench, reference arm, ncu launch list, ncu --set full of the changed GEMM
mkdir -p gpurun_out; cd "$(dirname "$0")/../.."
echo "== pytest all gpu"; timeout -s KILL 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2az_pytest_all.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/r2az_pytest_all.log | cut -c1-300
echo "== bench (default arguments)"; time (timeout -s KILL 900 python bench.py > gpurun_out/r2az_bench.json 2> gpurun_out/r2az_bench.err); echo "rc=$?"; tail -2 gpurun_out/r2az_bench.err | cut -c1-300
cp gpurun_out/kernels_by_shape.txt gpurun_out/r2az_kernels_by_shape.txt
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2az_bench.json'))
print({k:d[k] for k in ('value','ms_per_step','launches_per_step','gpu_launches')}); print('e2e', d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['per_step_ms'])
print('modes', d['modes']); print('stock', {k:v for k,v in d['gpu_stock_baseline'].items() if k!='what'}); print('strong', d['strong_scaling'])
print('roofline', {k:v for k,v in d['roofline'].items() if k not in ('families','note')})
print('cpu', d['cpu_baseline']); print('clocks', d.get('clocks'))
print('aux', {k:(v if not isinstance(v,dict) else {kk:vv for kk,vv in v.items() if kk in ('value','ms_per_step','unit')}) for k,v in d.get('aux',{}).items()})
PY
echo "== reference arm"; time (timeout -s KILL 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2az_bench_reference.json 2> gpurun_out/r2az_bench_reference.err); tail -c 500 gpurun_out/r2az_bench_reference.json
echo "== ncu launch list"; timeout -s KILL 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2az_launches.csv python bench.py --steps 2 --warmup 3 --no-aux --no-stock --no-strong --no-modes --no-cpu-baseline --no-kernel-profile > gpurun_out/r2az_ncu_launches.log 2>&1; echo "rc=$?"; wc -l gpurun_out/r2az_launches.csv
for k in gemm_tf32_kernel; do
  echo "== ncu full $k"; timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 30 -c 12 -o gpurun_out/r2az_ncu_$k -f python bench.py --steps 2 --warmup 3 --no-aux --no-stock --no-strong --no-modes --no-cpu-baseline --no-kernel-profile > gpurun_out/r2az_ncu_$k.log 2>&1; echo "rc=$?"
done
ls -la gpurun_out/r2az_ncu_*.ncu-rep
