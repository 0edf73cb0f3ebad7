#!/bin/bash
# Visit 6 (session re-entry): full gpu suite, pretrain bench, launch list for fresh shares
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
timeout 600 python bench.py --workload pretrain --steps 10 --warmup 3 > gpurun_out/bench_pretrain.json 2> gpurun_out/bench_pretrain.err; cat gpurun_out/bench_pretrain.json; tail -5 gpurun_out/bench_pretrain.err
timeout 300 python bench.py --workload histogram > gpurun_out/bench_hist.json 2> gpurun_out/bench_hist.err; cat gpurun_out/bench_hist.json; tail -3 gpurun_out/bench_hist.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r1_pretrain_launches.csv python bench.py --workload pretrain --steps 1 --warmup 3 --no-cpu-baseline --batch 128 > gpurun_out/ncu_pretrain.log 2>&1; tail -3 gpurun_out/ncu_pretrain.log
ls -la gpurun_out
