#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_vit_kernels_gpu.py tests/test_vit_model_gpu.py tests/test_engine_gpu.py tests/test_engine_ft_gpu.py -q -x 2>&1 | grep -v Warning | tail -8 | tee gpurun_out/r02_s2_pytest_v5.log
timeout 900 python bench.py --no-histogram --no-cpu-baseline > gpurun_out/r02_bench_pretrain_s2_v2.json 2> gpurun_out/r02_bench_pretrain_s2_v2.err; tail -3 gpurun_out/r02_bench_pretrain_s2_v2.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_pretrain_s2_v2.json'))
print(d['value'], d['ms_per_step'], d['breakdown_ms'], d['e2e']['value'], d['clocks'])
PY
