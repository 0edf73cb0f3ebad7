#!/bin/bash
# per-layer table of one 128-image tokenizer pass: time, tensor-pipe activity, L2 and DRAM throughput (ncu, one pass)
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed,dram__throughput.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off -k regex:"conv_f16x2|im2col" --csv --log-file gpurun_out/r02_dvae_layers_b128_v2.csv python tools/dvae_conv_prof.py 128 > /dev/null 2>&1
python - <<'PY'
import csv
rows = {}
lines = [l for l in open("gpurun_out/r02_dvae_layers_b128_v2.csv") if l.startswith('"')]
for d in csv.DictReader(lines):
    rows.setdefault(int(d["ID"]), {"kernel": d["Kernel Name"].split("(")[0]})[d["Metric Name"]] = float(d["Metric Value"].replace(",", ""))
names = ["im2col", "act0", "act1", "act2", "act3"] + [f"res{j}_{k}" for j in range(3) for k in range(3)] + ["head"]
out = ["# one 128-image tokenizer pass under ncu (isolated launches, unthrottled clocks): layer, us, tensor-pipe active %, L2 %, DRAM %, DRAM read MB, DRAM write MB"]
tot = 0.0
for (i, r), n in zip(sorted(rows.items()), names):
    us = r["gpu__time_duration.sum"] / 1e3
    tot += us
    out.append(f"{n:8s} {us:8.1f} us  tensor {r['sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_elapsed']:5.1f}  l2 {r['lts__throughput.avg.pct_of_peak_sustained_elapsed']:5.1f}  "
               f"dram {r['dram__throughput.avg.pct_of_peak_sustained_elapsed']:5.1f}  read {r['dram__bytes_read.sum'] / 1e6:8.1f}  write {r['dram__bytes_write.sum'] / 1e6:8.1f}")
out.append(f"total    {tot:8.1f} us")
open("gpurun_out/r02_dvae_layers_b128_v2.txt", "w").write("\n".join(out) + "\n")
print("\n".join(out))
PY
