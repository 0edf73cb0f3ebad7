#!/bin/bash
# 2 GPUs: how many NCCL CTAs does the overlapped gradient exchange need / how much do they cost the backward GEMMs?
mkdir -p gpurun_out
for c in default 2 4 8; do
  if [ $c = default ]; then unset NCCL_MAX_CTAS; export NCCL_DEBUG=INFO; else export NCCL_MAX_CTAS=$c; unset NCCL_DEBUG; fi
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline --no-histogram > gpurun_out/r02_scale_n2_ctas_$c.json 2> gpurun_out/r02_scale_n2_ctas_$c.err
  python - <<PY
import json
d=json.load(open('gpurun_out/r02_scale_n2_ctas_$c.json'))
print('NCCL_MAX_CTAS=$c', d['value'], d['ms_per_step'], d['breakdown_ms'], d['e2e']['value'])
PY
done
grep -i "channels\|nvls\|ctas" gpurun_out/r02_scale_n2_ctas_default.err | head -12
