#!/bin/bash
# 8- and 4-GPU runs of the pretraining bench on one box (weak scaling, batch 128 per GPU)
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/scale_n8.json 2> gpurun_out/scale_n8.err; cut -c1-200 gpurun_out/scale_n8.json; tail -3 gpurun_out/scale_n8.err
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/scale_n4.json 2> gpurun_out/scale_n4.err; cut -c1-200 gpurun_out/scale_n4.json; tail -3 gpurun_out/scale_n4.err
