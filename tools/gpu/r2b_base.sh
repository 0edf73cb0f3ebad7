#!/bin/bash
# round 2 (session 2) baseline: full GPU suite, smoke, default bench line
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | grep -v Warning | tail -6 | tee gpurun_out/r02_pytest_gpu_full_s2_v0.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/r02_bench_pretrain_s2_v0.json 2> gpurun_out/r02_bench_pretrain_s2_v0.err; cat gpurun_out/r02_bench_pretrain_s2_v0.json; tail -3 gpurun_out/r02_bench_pretrain_s2_v0.err
