#!/bin/bash
# Visit 8: ncu of the tcgen05 attention kernels
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_fwd_tc -s 1 -c 1 -f -o gpurun_out/r1_attn_fwd_v3 python tools/attn_prof.py ncu > gpurun_out/ncu_attn_fwd.log 2>&1; tail -3 gpurun_out/ncu_attn_fwd.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_bwd_tc -s 1 -c 1 -f -o gpurun_out/r1_attn_bwd_v3 python tools/attn_prof.py ncu > gpurun_out/ncu_attn_bwd.log 2>&1; tail -3 gpurun_out/ncu_attn_bwd.log
ls -la gpurun_out/*.ncu-rep
