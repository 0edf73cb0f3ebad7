#!/bin/bash
# round-end check: full GPU suite, smoke, default bench, reference arm
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2 | tee gpurun_out/r02_pytest_gpu_full_final.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/r02_bench_pretrain_final.json 2> gpurun_out/r02_bench_pretrain_final.err; cut -c1-160 gpurun_out/r02_bench_pretrain_final.json; tail -2 gpurun_out/r02_bench_pretrain_final.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference_final.json 2> gpurun_out/r02_bench_reference_final.err; cut -c1-200 gpurun_out/r02_bench_reference_final.json
