"""GPU: the replicated-plane GLOBAL strategy (MEMB_HIST_GLOBAL_REPL) against plain GLOBAL on the config-2 top
sizes; the replica count comes from MEMB_HIST_REPLICAS (read once per process), so run once per count:
    for k in 2 4 8 16; do MEMB_HIST_REPLICAS=$k python tools/hist_repl_sweep.py; done
Appends JSON lines to gpurun_out/hist_repl_sweep.jsonl."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mem_b200.process_data import histogram  # noqa: E402
from oracle.make_golden import synth_events  # noqa: E402
from tools.hist_sweep import hot_pixel_events, timeit  # noqa: E402


def main():
    k = int(os.environ.get("MEMB_HIST_REPLICAS", "8"))
    rng = np.random.default_rng(0)
    out = open(os.path.join(ROOT, "gpurun_out", "hist_repl_sweep.jsonl"), "a")
    for (W, H) in [(640, 480), (1280, 720)]:
        for kind in ["uniform", "edge", "hot"]:
            for n in [1_000_000, 10_000_000]:
                ev = hot_pixel_events(rng, n, H, W) if kind == "hot" else synth_events(rng, n, H, W, kind)
                d = torch.from_numpy(ev).cuda()
                ref = histogram(d, H, W, strategy=1, check=False)
                for s in (1, 5, 6):
                    got = histogram(d, H, W, strategy=s, check=False)
                    assert torch.equal(got, ref), (W, H, kind, n, s)
                    ms = timeit(lambda: histogram(d, H, W, strategy=s, check=False))
                    row = {"sensor": f"{W}x{H}", "kind": kind, "n": n, "strategy": s, "replicas": k if s == 5 else 1,
                           "us": ms * 1e3, "gev_s": n / ms / 1e6, "gb_s": (32 * n + 3 * H * W) / ms / 1e6}
                    print(row, flush=True)
                    out.write(json.dumps(row) + "\n")


if __name__ == "__main__":
    main()
