#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_histogram_gpu.py tests/test_vit_model_gpu.py tests/test_engine_ft_gpu.py -q -x 2>&1 | grep -v Warning | tail -30 | tee gpurun_out/r02_pytest_hist_v2.log
rm -f gpurun_out/hist_repl_sweep.jsonl
MEMB_HIST_REPLICAS=8 timeout 300 python tools/hist_repl_sweep.py 2>&1 | tail -40
