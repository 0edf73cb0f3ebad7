#!/bin/bash
# library table, per-kernel launch shares of the pretraining step, ncu --set full of the tokenizer's convolution launches (DRAM traffic)
mkdir -p gpurun_out
timeout 900 python tools/lib_compare.py 2>&1 | grep -v Warning | tail -30
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02_pretrain_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-histogram > /dev/null 2>&1
python tools/launch_shares.py gpurun_out/r02_pretrain_launches.csv 45 > gpurun_out/r02_pretrain_launch_shares_v1.txt; head -30 gpurun_out/r02_pretrain_launch_shares_v1.txt
rm -f gpurun_out/r02_pretrain_launches.csv
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_f16x2 -s 28 -c 14 -o gpurun_out/r02_conv_f16x2_full python tools/dvae_conv_prof.py 128 > /dev/null 2>&1
ncu -i gpurun_out/r02_conv_f16x2_full.ncu-rep --page raw --csv --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct 2>/dev/null > gpurun_out/r02_conv_f16x2_full_keys.csv
cut -c1-250 gpurun_out/r02_conv_f16x2_full_keys.csv | tail -16
ls -la gpurun_out/r02_conv_f16x2_full.ncu-rep
