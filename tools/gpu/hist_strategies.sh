#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_histogram_gpu.py -q -x 2>&1 | grep -v Warning | tail -5
for k in uniform edge hot; do timeout 120 python tools/hist_one.py 7 $k 10000000; done
timeout 120 python tools/hist_one.py 7 uniform 10000000 1280x720
timeout 120 python tools/hist_one.py 7 uniform 1000000
timeout 120 python tools/hist_one.py 7 uniform 3000000
for k in uniform edge; do
  timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum --clock-control none -k regex:"hist_sort" --csv --log-file gpurun_out/sort_$k.csv python tools/hist_one.py 7 $k 10000000 > /dev/null 2>&1
  python - <<PY
import csv,io
rows=[l for l in open("gpurun_out/sort_$k.csv") if l.startswith('"')]
rd=list(csv.reader(io.StringIO("".join(rows)))); h=rd[0]; ki=h.index("Kernel Name"); vi=h.index("Metric Value"); mi=h.index("Metric Name")
from collections import defaultdict
d=defaultdict(list)
for r in rd[1:]: d[(r[ki][:40], r[mi])].append(float(r[vi].replace(",","")))
for a,v in d.items(): print("$k", a, "n=%d median %.1f"%(len(v), sorted(v)[len(v)//2]))
PY
done
for k in uniform edge8 edge50 hot; do timeout 300 python bench.py --workload histogram --dist $k --no-cpu-baseline | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$k', d['value'], d['ms_per_step'], d['roofline']['frac'], d['parity_vs_oracle'])"; done
