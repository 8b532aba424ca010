#!/bin/bash
# round 2, call C: where does the recurrence step go - per-group landing stamps, multicast vs unicast all-gather; sanitizer on the cluster kernels
mkdir -p gpurun_out; cd "$(dirname "$0")/../.."
echo "== trace multicast"; timeout -s KILL 180 python tests/trace_gru.py > gpurun_out/r2c_trace_mc.log 2>&1; echo "rc=$?"; cat gpurun_out/r2c_trace_mc.log
echo "== trace unicast"; TGB200_GRU_UNICAST=1 timeout -s KILL 180 python tests/trace_gru.py > gpurun_out/r2c_trace_uc.log 2>&1; echo "rc=$?"; cat gpurun_out/r2c_trace_uc.log
echo "== unicast correctness"; TGB200_GRU_UNICAST=1 timeout -s KILL 300 python -m pytest tests/test_gpu_tf32.py -m gpu -q -x -k gru_layer_tensor_core -p no:cacheprovider 2>&1 | tail -3
echo "== memcheck"; timeout -s KILL 600 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_tf32.py -m gpu -q -x -k "gru_layer_tensor_core and 3-34" -p no:cacheprovider > gpurun_out/r2c_memcheck.log 2>&1; echo "rc=$?"; tail -12 gpurun_out/r2c_memcheck.log
echo "== racecheck"; timeout -s KILL 600 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_gpu_tf32.py -m gpu -q -x -k "gru_layer_tensor_core and 3-34" -p no:cacheprovider > gpurun_out/r2c_racecheck.log 2>&1; echo "rc=$?"; tail -12 gpurun_out/r2c_racecheck.log
