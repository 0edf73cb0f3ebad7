#!/bin/bash
# round 2, visit 1: replicated-plane histogram (correctness + replica sweep), tokenizer per-layer profile at 128 images, baseline bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_histogram_gpu.py -x -q 2>&1 | tail -3 | tee gpurun_out/r02_pytest_hist_v1.log
rm -f gpurun_out/hist_repl_sweep.jsonl
for k in 2 4 8 16; do MEMB_HIST_REPLICAS=$k timeout 300 python tools/hist_repl_sweep.py 2>&1 | tail -24; done
timeout 600 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed,dram__throughput.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:"conv_f16x2|im2col" --csv --log-file gpurun_out/r02_dvae_layers_b128.csv python tools/dvae_conv_prof.py 128 > /dev/null 2>&1
tail -40 gpurun_out/r02_dvae_layers_b128.csv | cut -c1-200
timeout 600 python bench.py > gpurun_out/r02_bench_pretrain_v0.json 2> gpurun_out/r02_bench_pretrain_v0.err; cut -c1-400 gpurun_out/r02_bench_pretrain_v0.json; tail -2 gpurun_out/r02_bench_pretrain_v0.err
