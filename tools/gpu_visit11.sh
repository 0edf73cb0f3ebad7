#!/bin/bash
# Visit 11: ncu --set full of the two roofline kernels (DRAM traffic per launch)
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_pair_kernel -s 2 -c 1 -f -o gpurun_out/r1_gemm_fc1_pair python tools/gemm_prof.py fc1_gelu > gpurun_out/ncu_fc1.log 2>&1; tail -2 gpurun_out/ncu_fc1.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:hist_scatter_global -s 5 -c 1 -f -o gpurun_out/r1_hist_scatter_v1 python bench.py --workload histogram --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_hist.log 2>&1; tail -2 gpurun_out/ncu_hist.log
ls -la gpurun_out/*.ncu-rep
