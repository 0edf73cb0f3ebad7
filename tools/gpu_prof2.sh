#!/bin/bash
mkdir -p gpurun_out
python tools/gemm_prof.py 2>&1 | tee gpurun_out/gemm_prof.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 2 -c 1 -o gpurun_out/r1_gemm_qkv python tools/gemm_prof.py qkv_store > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 2 -c 1 -o gpurun_out/r1_gemm_fc1 python tools/gemm_prof.py fc1_gelu > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tf32x3 -s 3 -c 1 -o gpurun_out/r1_conv_l2 python tools/dvae_probe.py > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
