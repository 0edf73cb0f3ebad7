#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_histogram_gpu.py -q -x 2>&1 | grep -v Warning | tail -3
timeout 600 python tools/hist_dbg.py | tail -3
timeout 600 python bench.py --workload histogram --dist uniform | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('uniform', d['value'], d['ms_per_step'], d['roofline']['frac'])"
for k in edge8 edge50 hot; do timeout 300 python bench.py --workload histogram --dist $k --no-cpu-baseline | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$k', d['value'], d['ms_per_step'], d['roofline']['frac'], d['parity_vs_oracle'])"; done
timeout 300 python bench.py --workload histogram --dist edge8 --sensor 1280x720 --no-cpu-baseline | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('edge8 1280x720', d['value'], d['ms_per_step'], d['roofline']['frac'], d['parity_vs_oracle'])"
