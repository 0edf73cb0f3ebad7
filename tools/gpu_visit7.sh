#!/bin/bash
# Visit 7: tcgen05 attention bring-up
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_vit_kernels_gpu.py -x -q -k attention 2>&1 | tail -15 > gpurun_out/pytest_attn.log; cat gpurun_out/pytest_attn.log
timeout 300 python tools/attn_prof.py legacy 2>&1 | tail -12 | tee gpurun_out/attn_prof.log
