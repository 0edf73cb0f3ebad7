#!/bin/bash
# BASELINE config 4: ViT-L/16 MEM pretraining across 8 x B200 (weak scaling, batch 128 per GPU)
mkdir -p gpurun_out
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --model large --steps 8 --warmup 3 > gpurun_out/scale_large_n8.json 2> gpurun_out/scale_large_n8.err; cut -c1-200 gpurun_out/scale_large_n8.json; tail -2 gpurun_out/scale_large_n8.err
