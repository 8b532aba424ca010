#!/bin/bash
# round 2, call AD (1 GPU): compute-sanitizer over the kernels added after the GRU work (fused discriminator, clip-group TCN GEMM, split-K)
mkdir -p gpurun_out; cd "$(dirname "$0")/../.."
SAN=/usr/local/cuda/bin/compute-sanitizer
for tool in memcheck racecheck synccheck; do
  echo "== $tool fused D (B = 21)"; timeout -s KILL 600 $SAN --tool $tool --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "fused_stack and 21" > gpurun_out/r2ad_${tool}_dfused.log 2>&1; echo "rc=$?"; grep -E "ERROR SUMMARY|passed|failed|RACECHECK SUMMARY" gpurun_out/r2ad_${tool}_dfused.log | tail -3
done
for tool in memcheck racecheck; do
  echo "== $tool two-tap GEMM (B = 3) + split-K + fused dropout"; timeout -s KILL 600 $SAN --tool $tool --print-limit 20 python -m pytest tests/test_gpu_tf32.py -m gpu -q -x -p no:cacheprovider -k "(two_tap and 3-) or (gemm_tf32_plain and 1000) or (fused_dropout and 5)" > gpurun_out/r2ad_${tool}_gemm.log 2>&1; echo "rc=$?"; grep -E "ERROR SUMMARY|passed|failed|RACECHECK SUMMARY" gpurun_out/r2ad_${tool}_gemm.log | tail -3
done
