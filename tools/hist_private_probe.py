"""Launch the PRIVATE rasteriser at 240x180 (for ncu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mem_b200.process_data import histogram
from oracle.make_golden import synth_events
rng = np.random.default_rng(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
d = torch.from_numpy(synth_events(rng, n, 180, 240, "edge")).cuda()
for _ in range(4):
    histogram(d, 180, 240, strategy=4, check=False)
torch.cuda.synchronize()
