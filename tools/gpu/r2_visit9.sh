#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_dvae_gpu.py tests/test_engine_gpu.py tests/test_histogram_gpu.py -q -x 2>&1 | grep -v Warning | tail -6
for s in 0; do for k in uniform edge hot; do timeout 120 python tools/hist_one.py $s $k 10000000; done; done
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/r02_bench_pretrain_v2.json 2> gpurun_out/r02_bench_pretrain_v2.err; tail -2 gpurun_out/r02_bench_pretrain_v2.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_pretrain_v2.json'))
print({k:d[k] for k in ['value','ms_per_step','breakdown_ms']}, d['step_tensor_util']['frac'], d['e2e']['value'])
r=d['roofline']; print(r['achieved'], r['frac'], r['frac_of_scheme_ceiling'], r['step_share_ms'])
for k,v in r['per_layer'].items(): print('  ',k,v)
for k,v in d['histogram']['streams'].items(): print(k, v['value'], v['us_per_step'], v['roofline']['frac'], v['parity_vs_oracle'])
PY
