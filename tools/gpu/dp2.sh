#!/bin/bash
# 2 GPUs: DP parity (both wire formats) and the pretrain bench with fp32 / bf16 gradient wire
mkdir -p gpurun_out
for w in fp32 bf16; do
  MEMB_DP_WIRE=$w timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tools/dp_check.py 2>&1 | grep "rank" | tee -a gpurun_out/r02_dp2_check_s2.log
done
for w in fp32 bf16; do
  MEMB_DP_WIRE=$w timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline --no-histogram > gpurun_out/r02_scale_n2_$w.json 2> gpurun_out/r02_scale_n2_$w.err
  python - <<PY
import json
d=json.load(open('gpurun_out/r02_scale_n2_$w.json'))
print('$w', d['value'], d['ms_per_step'], d['breakdown_ms'], d['e2e']['value'])
PY
done
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-histogram | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('n1', d['value'], d['ms_per_step'], d['breakdown_ms'], d['e2e']['value'])"
