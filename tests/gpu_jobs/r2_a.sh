#!/bin/bash
# round 2, call A: validate the cluster GRU kernels, compare with the legacy kernels, then the whole GPU suite + bench
mkdir -p gpurun_out; cd "$(dirname "$0")/../.."
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2a_smi.txt 2>&1
echo "== trace cluster" ; timeout -s KILL 180 python tests/trace_gru.py > gpurun_out/r2a_trace_cluster.log 2>&1; echo "rc=$?"; tail -30 gpurun_out/r2a_trace_cluster.log
echo "== trace legacy" ; TGB200_GRU_LEGACY=1 timeout -s KILL 180 python tests/trace_gru.py > gpurun_out/r2a_trace_legacy.log 2>&1; echo "rc=$?"; grep -E "median|resident" gpurun_out/r2a_trace_legacy.log
echo "== pytest gru tc"; timeout -s KILL 600 python -m pytest tests/test_gpu_tf32.py -m gpu -q -x -k gru_layer_tensor_core -p no:cacheprovider > gpurun_out/r2a_pytest_gru.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/r2a_pytest_gru.log
echo "== pytest all gpu"; timeout -s KILL 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2a_pytest_all.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/r2a_pytest_all.log
echo "== bench"; timeout -s KILL 600 python bench.py > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "rc=$?"; cat gpurun_out/r2a_bench.json | head -c 3000
