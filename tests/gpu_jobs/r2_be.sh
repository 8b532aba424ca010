#!/bin/bash
# round 2, call BE (1 GPU): head weight gradient on the tensor cores (152-float pitch); full GPU suite; bench; timeline
mkdir -p gpurun_out; cd "$(dirname "$0")/../.."
B="--no-aux --no-stock --no-strong --no-modes --no-cpu-baseline --no-kernel-profile"
echo "== pytest all gpu"; timeout -s KILL 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2be_pytest_all.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/r2be_pytest_all.log | cut -c1-300
for name in a b; do
  timeout -s KILL 300 python bench.py $B > gpurun_out/r2be_bench_$name.json 2> gpurun_out/r2be_bench_$name.err; echo "rc=$?"
  python - <<PY
import json
d=json.load(open('gpurun_out/r2be_bench_$name.json'))
print('$name', {k:d[k] for k in ('value','ms_per_step')}, 'e2e', d['e2e']['value'])
PY
done
echo "== timeline"; timeout -s KILL 300 python tests/timeline_step.py gpurun_out/r2be_timeline.csv > gpurun_out/r2be_timeline.log 2>&1; echo "rc=$?"
python tests/timeline_to_txt.py gpurun_out/r2be_timeline.json gpurun_out/r2be_timeline_step.txt && python tests/analyze_timeline.py gpurun_out/r2be_timeline_step.txt > gpurun_out/r2be_timeline_step_ownership.txt; head -8 gpurun_out/r2be_timeline_step_ownership.txt
rm -f gpurun_out/r2be_timeline.json
