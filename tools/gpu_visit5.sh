#!/bin/bash
# Visit 5: fp16-pair dVAE conv bring-up
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_dvae_gpu.py -x -q 2>&1 | tail -25 > gpurun_out/pytest_dvae.log; cat gpurun_out/pytest_dvae.log
timeout 600 python tools/dvae_probe.py 2>&1 | tail -12 | tee gpurun_out/dvae_probe.log
timeout 600 python bench.py --workload pretrain --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_pretrain.json 2> gpurun_out/bench_pretrain.err; cat gpurun_out/bench_pretrain.json; tail -3 gpurun_out/bench_pretrain.err
