#!/bin/bash
# Runs every GPU test node in its own process (a CUDA fault in one test cannot poison the next); logs under gpurun_out/.
out=${1:-gpurun_out/each.log}
mkdir -p "$(dirname "$out")"; : > "$out"
for t in $(python -m pytest tests/test_gpu_parity.py -m gpu --collect-only -q -p no:cacheprovider 2>/dev/null | grep '::'); do
  echo "=== $t" >> "$out"
  CUDA_LAUNCH_BLOCKING=1 timeout 300 python -m pytest "$t" -q -x -p no:cacheprovider 2>&1 | grep -vE "^$|FutureWarning|WeightNorm.apply|warnings summary|Docs:" | tail -45 >> "$out"
done
grep -E "^===|passed|failed" "$out"
