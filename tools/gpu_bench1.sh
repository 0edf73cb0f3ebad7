#!/bin/bash
# pretraining-step bench + launch list
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv
timeout 900 python -m pytest tests/test_vit_model_gpu.py tests/test_engine_gpu.py -q -m gpu 2>&1 | tail -5
timeout 900 python bench.py --workload pretrain --steps 10 --warmup 3 > gpurun_out/bench_pretrain.json 2> gpurun_out/bench_pretrain.err; cat gpurun_out/bench_pretrain.json; tail -20 gpurun_out/bench_pretrain.err
timeout 600 python bench.py --impl reference --workload pretrain --steps 2 --warmup 1 > gpurun_out/bench_pretrain_ref.json 2>gpurun_out/bench_pretrain_ref.err; cat gpurun_out/bench_pretrain_ref.json; tail -3 gpurun_out/bench_pretrain_ref.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r1_pretrain_launches.csv python bench.py --workload pretrain --steps 1 --warmup 3 --no-cpu-baseline --batch 128 > gpurun_out/ncu_pretrain.log 2>&1; tail -3 gpurun_out/ncu_pretrain.log
ls -la gpurun_out
