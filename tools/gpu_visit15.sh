#!/bin/bash
# validation visit after container re-creation: full GPU suite, smoke, both bench workloads, reference arm
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/pytest_gpu_full_v5.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
timeout 600 python bench.py > gpurun_out/bench_pretrain_v6.json 2> gpurun_out/bench_pretrain_v6.err; cut -c1-400 gpurun_out/bench_pretrain_v6.json; tail -3 gpurun_out/bench_pretrain_v6.err
timeout 300 python bench.py --workload histogram > gpurun_out/bench_hist_v3.json 2> gpurun_out/bench_hist_v3.err; cut -c1-300 gpurun_out/bench_hist_v3.json
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_v1.json 2> gpurun_out/bench_ref_v1.err; cut -c1-300 gpurun_out/bench_ref_v1.json
