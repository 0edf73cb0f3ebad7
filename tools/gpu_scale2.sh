#!/bin/bash
# 1 -> 2 GPU scaling of the pretraining bench on one box (+ the DP parity check)
mkdir -p gpurun_out
timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/scale_n1.json 2> gpurun_out/scale_n1.err; cut -c1-400 gpurun_out/scale_n1.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/scale_n2.json 2> gpurun_out/scale_n2.err; cut -c1-400 gpurun_out/scale_n2.json; tail -3 gpurun_out/scale_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/dp_check.py > gpurun_out/dp2_check.log 2>&1; tail -4 gpurun_out/dp2_check.log
