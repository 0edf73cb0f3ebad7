#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_dvae_train_gpu.py -q -x -s 2>&1 | grep -v Warning | tail -40 | tee gpurun_out/r02_pytest_dvae_train_v1.log
timeout 900 python -m pytest tests/test_histogram_gpu.py tests/test_vit_model_gpu.py tests/test_engine_ft_gpu.py -q -x 2>&1 | grep -v Warning | tail -6
for s in 0 6; do for k in uniform edge hot; do timeout 120 python tools/hist_one.py $s $k 10000000; done; done
for k in uniform edge; do
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"hist_" --csv --log-file gpurun_out/hyb_$k.csv python tools/hist_one.py 0 $k 10000000 > /dev/null 2>&1
  python - <<PY
import csv,io
rows=[l for l in open("gpurun_out/hyb_$k.csv") if l.startswith('"')]
rd=list(csv.reader(io.StringIO("".join(rows)))); h=rd[0]; ki=h.index("Kernel Name"); vi=h.index("Metric Value")
from collections import defaultdict
d=defaultdict(list)
for r in rd[1:]: d[r[ki][:50]].append(float(r[vi].replace(",",""))/1000)
for a,v in d.items(): print("$k", a, "n=%d median %.1f us"%(len(v), sorted(v)[len(v)//2]))
PY
done
