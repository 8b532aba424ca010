#!/bin/bash
# round 2, call BF (1 GPU): D(real) forked between a layer's projection and its recurrence ('preN') against behind the second recurrence ('gru1')
mkdir -p gpurun_out; cd "$(dirname "$0")/../.."
B="--no-aux --no-stock --no-strong --no-modes --no-cpu-baseline --no-kernel-profile"
for name in gru1 pre1 pre2 pre0; do
  TGB200_DREAL_AT=$name timeout -s KILL 300 python bench.py $B > gpurun_out/r2bf_bench_$name.json 2> gpurun_out/r2bf_bench_$name.err; echo "rc=$?"
  python - <<PY
import json
d=json.load(open('gpurun_out/r2bf_bench_$name.json'))
print('$name', {k:d[k] for k in ('value','ms_per_step')}, 'e2e', d['e2e']['value'])
PY
done
