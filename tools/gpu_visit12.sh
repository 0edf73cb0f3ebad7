#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_vit_model_gpu.py -x -q -k large 2>&1 | tail -8
timeout 900 python bench.py --model large --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_pretrain_large.json 2> gpurun_out/bench_pretrain_large.err; cut -c1-700 gpurun_out/bench_pretrain_large.json; tail -5 gpurun_out/bench_pretrain_large.err
