#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -12 | tee gpurun_out/r02_pytest_gpu_full_v2.log
