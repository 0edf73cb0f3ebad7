#!/bin/bash
# ncu --set full + source of one tokenizer convolution launch: bash tools/gpu/dvae_ncu.sh <name> <launch index in the pass>
# (0 = act0, 1..3 = act1..act3, 4..12 = residual blocks, 13 = head; see tools/dvae_conv_prof.py)
mkdir -p gpurun_out
name=${1:-l1}; skip=${2:-0}
timeout 600 ncu --set full --import-source on --clock-control none --profile-from-start off -k regex:conv_f16x2 --launch-skip $skip --launch-count 1 \
  -o gpurun_out/r02_ncu_conv_$name -f python tools/dvae_conv_prof.py 128 > /dev/null 2>&1
ls -la gpurun_out/r02_ncu_conv_$name.ncu-rep
