"""Top stall sites of an ncu source-page CSV (ncu -i X.ncu-rep --page source --csv)."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
i_src, i_samp, i_exec = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stall = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
body = rows[2:]
tot = sum(int(r[i_samp] or 0) for r in body)
print("total samples", tot, "instructions", sum(int(r[i_exec] or 0) for r in body))
top = sorted(range(len(body)), key=lambda k: -int(body[k][i_samp] or 0))[: int(sys.argv[2]) if len(sys.argv) > 2 else 25]
for k in sorted(top):
    r = body[k]
    st = sorted(((int(r[i] or 0), h) for i, h in stall), reverse=True)[:3]
    print(f"{k:5d} {int(r[i_samp]):6d} {100 * int(r[i_samp]) / tot:5.1f}%  x{r[i_exec]:>8s}  {r[i_src].strip()[:70]:70s} {[(h[6:], n) for n, h in st if n]}")
