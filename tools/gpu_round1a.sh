#!/bin/bash
# First GPU visit: parity tests, bench, sweep, ncu launch list + full capture of the scatter kernel.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
timeout 600 python bench.py --workload histogram --steps 50 --warmup 10 > gpurun_out/bench_hist.json 2> gpurun_out/bench_hist.err; cat gpurun_out/bench_hist.json; tail -5 gpurun_out/bench_hist.err
timeout 300 python bench.py --impl reference --workload histogram --steps 5 --warmup 2 > gpurun_out/bench_hist_ref.json 2>&1; cat gpurun_out/bench_hist_ref.json
timeout 600 python tools/hist_sweep.py > gpurun_out/hist_sweep.log 2>&1; tail -70 gpurun_out/hist_sweep.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r1_hist_launches.csv python bench.py --workload histogram --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:hist_scatter_global -s 3 -c 2 -o gpurun_out/r1_hist_scatter python bench.py --workload histogram --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
