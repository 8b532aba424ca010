#!/bin/bash
mkdir -p gpurun_out; cd "$(dirname "$0")/../.."
echo "== new"; timeout -s KILL 120 python tests/trace_gemm_tcn.py 2>&1 | grep -v Warn | tail -4
echo "== old"; TGB200_TCN_TWO_ACC=1 timeout -s KILL 120 python tests/trace_gemm_tcn.py 2>&1 | grep -v Warn | tail -4
