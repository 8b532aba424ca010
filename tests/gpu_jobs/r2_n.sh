#!/bin/bash
# round 2, call N: ncu --set full with source of the fused discriminator kernels
mkdir -p gpurun_out; cd "$(dirname "$0")/../.."
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:"dgru_stack|dconv_stack" -c 4 -o gpurun_out/r2n_ncu_dfused -f python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "fused_stack" > gpurun_out/r2n_ncu.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/r2n_ncu.log | cut -c1-200
ls -la gpurun_out/r2n_ncu_dfused.ncu-rep
