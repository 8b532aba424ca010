#!/bin/bash
# round 2, call BJ (1 GPU): full GPU suite on the final code
mkdir -p gpurun_out; cd "$(dirname "$0")/../.."
timeout -s KILL 200 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2bj_pytest_all.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/r2bj_pytest_all.log | cut -c1-300
