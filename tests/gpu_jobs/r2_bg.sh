#!/bin/bash
# round 2, call BG (1 GPU): __graft_entry__.smoke() of the final code
mkdir -p gpurun_out; cd "$(dirname "$0")/../.."
timeout -s KILL 300 python __graft_entry__.py smoke 2>&1 | tail -4 | cut -c1-300
