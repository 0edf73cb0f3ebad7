"""GPU: run one histogram configuration a few times (for ncu launch lists): python tools/hist_one.py <strategy> <kind> <n> [WxH]"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mem_b200.process_data import histogram
from oracle.make_golden import synth_events
from tools.hist_sweep import hot_pixel_events, timeit
s, kind, n = int(sys.argv[1]), sys.argv[2], int(sys.argv[3])
W, H = (int(v) for v in (sys.argv[4] if len(sys.argv) > 4 else "640x480").split("x"))
rng = np.random.default_rng(0)
ev = hot_pixel_events(rng, n, H, W) if kind == "hot" else synth_events(rng, n, H, W, kind)
d = torch.from_numpy(ev).cuda()
ms = timeit(lambda: histogram(d, H, W, strategy=s, check=False), iters=10)
print(f"strategy {s} {kind} n={n} {W}x{H}: {ms * 1e3:.1f} us")
