#!/bin/bash
mkdir -p gpurun_out
for b in all 32e6 2e6; do
  MEMB_DP_BUCKET=$b timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline --no-histogram > gpurun_out/r02_scale_n2_bucket_$b.json 2> gpurun_out/r02_scale_n2_bucket_$b.err
  python - <<PY
import json
d=json.load(open('gpurun_out/r02_scale_n2_bucket_$b.json'))
print('MEMB_DP_BUCKET=$b', d['value'], d['ms_per_step'], d['breakdown_ms'], d['e2e']['value'])
PY
done
