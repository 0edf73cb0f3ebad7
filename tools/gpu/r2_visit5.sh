#!/bin/bash
mkdir -p gpurun_out
for k in uniform edge; do
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"hist_" --csv --log-file gpurun_out/hyb_$k.csv python tools/hist_one.py 6 $k 10000000 > /dev/null 2>&1
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
for t in 1 100 400; do for k in uniform edge; do MEMB_HYB_TILE_GRANULES=$t timeout 120 python tools/hist_one.py 6 $k 10000000 | sed "s/^/tile_granules=$t /"; done; done
timeout 120 python tools/hist_one.py 6 uniform 1000000
timeout 120 python tools/hist_one.py 1 uniform 1000000
