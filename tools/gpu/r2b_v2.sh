#!/bin/bash
# session 2, visit 2: SORT histogram strategy (parity + timing), tokenizer with shared pad rows + relaxed cluster arrive
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_histogram_gpu.py tests/test_dvae_gpu.py -q -x 2>&1 | grep -v Warning | tail -8 | tee gpurun_out/r02_s2_pytest_v2.log
for k in uniform edge hot; do for s in 7 6 1; do timeout 120 python tools/hist_one.py $s $k 10000000; done; done
timeout 120 python tools/hist_one.py 7 uniform 10000000 1280x720
timeout 120 python tools/hist_one.py 6 uniform 10000000 1280x720
timeout 120 python tools/hist_one.py 7 uniform 1000000
timeout 120 python tools/hist_one.py 1 uniform 1000000
timeout 900 python bench.py > gpurun_out/r02_bench_pretrain_s2_v1.json 2> gpurun_out/r02_bench_pretrain_s2_v1.err; cut -c1-250 gpurun_out/r02_bench_pretrain_s2_v1.json; tail -3 gpurun_out/r02_bench_pretrain_s2_v1.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_pretrain_s2_v1.json'))
print(d['breakdown_ms'], d['e2e']['value'])
print({k:v['us'] for k,v in d['roofline']['per_layer'].items()})
print({k:(v['us_per_step'], v['roofline']['frac'], v['parity_vs_oracle']) for k,v in d['histogram']['streams'].items()})
PY
