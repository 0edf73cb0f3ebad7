#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_engine_gpu.py tests/test_vit_model_gpu.py -x -q 2>&1 | tail -5
timeout 600 python bench.py --workload pretrain --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_pretrain.json 2> gpurun_out/bench_pretrain.err; cut -c1-300 gpurun_out/bench_pretrain.json; python -c "
import json; d=json.load(open('gpurun_out/bench_pretrain.json')); print(d['value'], d['ms_per_step'], d['e2e'], d['gpu_launches'])"; tail -3 gpurun_out/bench_pretrain.err
