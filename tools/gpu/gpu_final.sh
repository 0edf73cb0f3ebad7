#!/bin/bash
# round-end check: full GPU suite, smoke, default bench, reference arm
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2 | tee gpurun_out/pytest_gpu_full_v8.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/bench_pretrain_v13.json 2> gpurun_out/bench_pretrain_v13.err; cut -c1-160 gpurun_out/bench_pretrain_v13.json; tail -2 gpurun_out/bench_pretrain_v13.err
