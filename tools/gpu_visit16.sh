#!/bin/bash
# full validation: GPU suite, smoke, every bench workload + reference arms
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2 | tee gpurun_out/pytest_gpu_full_v7.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 600 python bench.py > gpurun_out/bench_pretrain_v12.json 2> gpurun_out/bench_pretrain_v12.err; cut -c1-120 gpurun_out/bench_pretrain_v12.json; tail -2 gpurun_out/bench_pretrain_v12.err
timeout 300 python bench.py --workload histogram > gpurun_out/bench_hist_v5.json 2> /dev/null; cut -c1-120 gpurun_out/bench_hist_v4.json
timeout 300 python bench.py --workload histogram --sensor 240x180 > gpurun_out/bench_hist_240x180_v2.json 2> /dev/null; cut -c1-120 gpurun_out/bench_hist_240x180_v1.json
timeout 300 python bench.py --workload event_pipeline > gpurun_out/bench_evpipe_v6.json 2> /dev/null; cut -c1-120 gpurun_out/bench_evpipe_v5.json
timeout 300 python bench.py --workload raw_histogram > gpurun_out/bench_rawhist_v2.json 2> /dev/null; cut -c1-120 gpurun_out/bench_rawhist_v1.json
timeout 400 python bench.py --workload finetune > gpurun_out/bench_finetune_v1.json 2> /dev/null; cut -c1-120 gpurun_out/bench_finetune_v1.json
