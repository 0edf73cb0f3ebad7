#!/bin/bash
# round 2, visit 2: calibrated ViT parity (report mode first, then the verdict), full GPU suite, new bench line
mkdir -p gpurun_out
MEMB_PARITY_REPORT=1 timeout 900 python -m pytest tests/test_vit_model_gpu.py -x -q -s 2>&1 | grep -v "^$" | tail -40 | tee gpurun_out/r02_vit_parity_report_v1.log
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -25 | tee gpurun_out/r02_pytest_gpu_full_v1.log
timeout 900 python bench.py > gpurun_out/r02_bench_pretrain_v1.json 2> gpurun_out/r02_bench_pretrain_v1.err; cut -c1-300 gpurun_out/r02_bench_pretrain_v1.json; tail -3 gpurun_out/r02_bench_pretrain_v1.err
