#!/bin/bash
# Visit 3: full GPU suite on the current tree, GEMM shape timings, launch list of one pretrain step,
# ncu --set full of the fc1 / qkv GEMMs and the dVAE conv.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
timeout 300 python tools/gemm_prof.py 2>&1 | tee gpurun_out/gemm_prof.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r1_pretrain_launches.csv python bench.py --workload pretrain --steps 1 --warmup 1 --no-cpu-baseline --batch 128 > gpurun_out/ncu_pretrain.log 2>&1; tail -3 gpurun_out/ncu_pretrain.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 2 -c 1 -o gpurun_out/r1_gemm_fc1 python tools/gemm_prof.py fc1_gelu > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 2 -c 1 -o gpurun_out/r1_gemm_qkv python tools/gemm_prof.py qkv_store > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tf32x3 -s 3 -c 1 -o gpurun_out/r1_conv_l2 python tools/dvae_probe.py > /dev/null 2>&1
ls -la gpurun_out
