#!/bin/bash
# round 2, call E: forward recurrent weights in tensor memory; register-resident discriminator GRU
mkdir -p gpurun_out; cd "$(dirname "$0")/../.."
( nvidia-smi topo -m; lscpu | grep -i -E "numa|model name|^CPU\(s\)"; for d in /sys/bus/pci/devices/*; do [ -f $d/numa_node ] && echo "$(basename $d) $(cat $d/class) $(cat $d/numa_node)"; done | grep " 0x0302" ) > gpurun_out/r2e_topo.txt 2>&1
echo "== trace cluster" ; timeout -s KILL 180 python tests/trace_gru.py > gpurun_out/r2e_trace_cluster.log 2>&1; echo "rc=$?"; tail -24 gpurun_out/r2e_trace_cluster.log
echo "== pytest kernels"; timeout -s KILL 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k gru -p no:cacheprovider 2>&1 | tail -3
echo "== pytest gru tc"; timeout -s KILL 600 python -m pytest tests/test_gpu_tf32.py -m gpu -q -x -p no:cacheprovider > gpurun_out/r2e_pytest_tf32.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/r2e_pytest_tf32.log
echo "== pytest all gpu"; timeout -s KILL 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2e_pytest_all.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/r2e_pytest_all.log
echo "== bench"; timeout -s KILL 600 python bench.py --no-aux > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err; echo "rc=$?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2e_bench.json'))
print({k:d[k] for k in ('value','ms_per_step')}, d['e2e'], d['roofline']['by_kernel_ms_per_step'], d['roofline']['by_kernel_tflops'])
PY
cp gpurun_out/kernels_by_shape.txt gpurun_out/r2e_kernels_by_shape.txt
