#!/bin/bash
# ViT kernels + model parity on the B200
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_vit_kernels_gpu.py tests/test_vit_model_gpu.py -q -m gpu 2>&1 | tail -80 > gpurun_out/pytest_vit.log
cat gpurun_out/pytest_vit.log
