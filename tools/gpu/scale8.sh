#!/bin/bash
# 8 GPUs of one box: the driver's scaling command for N = 8 (and N = 1 on the same box for the ratio)
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29620 bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu-baseline --no-histogram > gpurun_out/r02_scale_n8.json 2> gpurun_out/r02_scale_n8.err
tail -2 gpurun_out/r02_scale_n8.err
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-histogram > gpurun_out/r02_scale_n1_samebox.json 2> gpurun_out/r02_scale_n1_samebox.err
python - <<'PY'
import json
for n in ("n8", "n1_samebox"):
    d = json.load(open(f"gpurun_out/r02_scale_{n}.json"))
    print(n, d["value"], d["ms_per_step"], d["breakdown_ms"], d["e2e"]["value"], d["clocks"]["sm_mhz"])
PY
