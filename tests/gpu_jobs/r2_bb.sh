#!/bin/bash
# round 2, call BB (1 GPU): compute-sanitizer memcheck / racecheck over the new GEMM modes (256-row tiles, 240-column tiles, one-wave split-K with
# accumulation, accumulating-tap data gradient, padded-pitch head), small shapes
mkdir -p gpurun_out; cd "$(dirname "$0")/../.."
T="tests/test_gpu_tf32.py"
IDS="$T::test_gemm_tf32_plain[300-1900-64] $T::test_gemm_tf32_plain[256-960-96] $T::test_gemm_tf32_plain[300-320-1024] $T::test_conv_dgrad_tf32_matches_conv_transpose[3-200-16-32-15-6] $T::test_conv_dgrad_tf32_matches_conv_transpose[5-100-8-16-5-5] $T::test_conv_dgrad_tf32_matches_conv_transpose[2-61-4-8-7-3] $T::test_gemm_tf32_padded_pitch_head[77]"
for tool in memcheck racecheck; do
  echo "== $tool"; timeout -s KILL 900 compute-sanitizer --tool $tool --print-limit 20 python -m pytest $IDS -m gpu -q -x -p no:cacheprovider > gpurun_out/r2bb_sanitizer_$tool.log 2>&1; echo "rc=$?"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" gpurun_out/r2bb_sanitizer_$tool.log | tail -4
done
