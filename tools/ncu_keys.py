#!/usr/bin/env python
"""Key metrics of one kernel from an ncu report: python tools/ncu_keys.py X.ncu-rep"""
import csv, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, val = rows[0], rows[-1]
keys = ["gpu__time_duration.sum", "sm__throughput.avg.pct", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct",
        "lts__t_bytes.sum", "lts__throughput.avg.pct", "l1tex__throughput.avg.pct", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma", "sm__inst_executed_pipe_alu.avg.pct", "sm__pipe_tensor", "sm__inst_executed_pipe_tensor",
        "smsp__issue_active.avg.pct", "smsp__inst_executed.sum ", "issue_stalled", "sm__warps_active.avg.pct", "launch__registers_per_thread ",
        "local_op_ld.sum ", "sm__cycles_elapsed.avg ", "sm__cycles_active.avg ", "smsp__cycles_active.avg ", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum ",
        "tmem", "uniform"]
for h, v in zip(hdr, val):
    hh = h + " "
    if any(k in hh for k in keys) and "device__" not in h:
        try:
            fv = float(v.replace(",", ""))
            if fv == 0: continue
        except ValueError:
            pass
        print(f"{h:110s} {v}")
