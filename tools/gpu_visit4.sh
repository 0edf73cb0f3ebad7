#!/bin/bash
# Visit 4: CTA-pair GEMM bring-up: parity tests, shape timings (pair vs single-CTA), model tests.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gemm_gpu.py -x -q 2>&1 | tail -25 > gpurun_out/pytest_gemm.log; cat gpurun_out/pytest_gemm.log
echo "== pair"; timeout 300 python tools/gemm_prof.py 2>&1 | tee gpurun_out/gemm_prof_pair.log
echo "== single"; MEMB_GEMM_SINGLE_CTA=1 timeout 300 python tools/gemm_prof.py 2>&1 | tee gpurun_out/gemm_prof_single.log
timeout 900 python -m pytest tests/test_vit_model_gpu.py tests/test_engine_gpu.py tests/test_vit_kernels_gpu.py -x -q 2>&1 | tail -8 > gpurun_out/pytest_vit.log; cat gpurun_out/pytest_vit.log
timeout 600 python bench.py --workload pretrain --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_pretrain.json 2> gpurun_out/bench_pretrain.err; cat gpurun_out/bench_pretrain.json; tail -3 gpurun_out/bench_pretrain.err
