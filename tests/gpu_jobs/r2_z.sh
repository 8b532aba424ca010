#!/bin/bash
# round 2, call Z (1 GPU): evidence run - full default bench, ncu launch list of the same command, ncu --set full of the new kernels
mkdir -p gpurun_out; cd "$(dirname "$0")/../.."
echo "== bench (default arguments)"; time (timeout -s KILL 900 python bench.py > gpurun_out/r2z_bench.json 2> gpurun_out/r2z_bench.err); echo "rc=$?"; tail -2 gpurun_out/r2z_bench.err | cut -c1-300
cp gpurun_out/kernels_by_shape.txt gpurun_out/r2z_kernels_by_shape.txt
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2z_bench.json'))
print({k:d[k] for k in ('value','ms_per_step','launches_per_step')}); print('e2e', d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['per_step_ms'])
print('modes', d['modes']); print('stock', {k:v for k,v in d['gpu_stock_baseline'].items() if k!='what'}); print('strong', d['strong_scaling'])
print('roofline', {k:v for k,v in d['roofline'].items() if k not in ('families','note')})
print('cpu', d['cpu_baseline'])
PY
echo "== reference arm"; time (timeout -s KILL 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2z_bench_reference.json 2> gpurun_out/r2z_bench_reference.err); tail -c 600 gpurun_out/r2z_bench_reference.json
echo "== ncu launch list"; timeout -s KILL 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2z_launches.csv python bench.py --steps 2 --warmup 3 --no-aux --no-stock --no-strong --no-modes --no-cpu-baseline --no-kernel-profile > gpurun_out/r2z_ncu_launches.log 2>&1; echo "rc=$?"; wc -l gpurun_out/r2z_launches.csv
for k in gemm_tcn_kernel dgru_stack_fwd dgru_stack_bwd gru_fwd_cl gru_bwd_ks; do
  echo "== ncu full $k"; timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 6 -c 1 -o gpurun_out/r2z_ncu_$k -f python bench.py --steps 2 --warmup 3 --no-aux --no-stock --no-strong --no-modes --no-cpu-baseline --no-kernel-profile > gpurun_out/r2z_ncu_$k.log 2>&1; echo "rc=$?"
done
ls -la gpurun_out/r2z_ncu_*.ncu-rep
