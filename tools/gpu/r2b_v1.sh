#!/bin/bash
# session 2, visit 1: tokenizer/ViT overlap probe; ncu --set full + source of the first tokenizer layer and the head
mkdir -p gpurun_out
timeout 600 python tools/overlap_probe.py 2>&1 | tail -4
SIDE_PRIO=-1 timeout 600 python tools/overlap_probe.py 2>&1 | tail -4
for spec in "l1:0" "head:13"; do
  name=${spec%%:*}; skip=${spec##*:}
  timeout 600 ncu --set full --import-source on --clock-control none --profile-from-start off -k regex:conv_f16x2 --launch-skip $skip --launch-count 1 \
    -o gpurun_out/r02_ncu_conv_$name -f python tools/dvae_conv_prof.py 128 > /dev/null 2>&1
  ls -la gpurun_out/r02_ncu_conv_$name.ncu-rep
done
