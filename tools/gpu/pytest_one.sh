#!/bin/bash
# run one pytest selection with full failure output: bash tools/gpu/pytest_one.sh <pytest args>
mkdir -p gpurun_out
timeout 1500 python -m pytest "$@" -q -x 2>&1 | grep -v Warning | tail -60 | tee gpurun_out/r02_one.log
