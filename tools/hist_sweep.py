"""GPU: time every rasteriser strategy over the BASELINE config-2 sweep (10k-10M events,
240x180 and 640x480, uniform / edge / hot-pixel) plus the ragged training batch.
Writes gpurun_out/hist_sweep.json and prints a table."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mem_b200.process_data import histogram, histogram_batch  # noqa: E402
from oracle.make_golden import synth_events  # noqa: E402

NAMES = {0: "auto", 1: "global", 2: "global_agg", 3: "tile", 4: "private", 6: "hybrid", 7: "sort"}


def timeit(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    ts = []
    for _ in range(iters):
        flush.zero_()                      # evict L2 between iterations
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def hot_pixel_events(rng, n, H, W):
    ev = synth_events(rng, n, H, W, "uniform")
    hot = rng.random(n) < 0.01
    hx, hy = rng.integers(0, W, 16), rng.integers(0, H, 16)
    pick = rng.integers(0, 16, n)
    ev[hot, 0], ev[hot, 1] = hx[pick[hot]], hy[pick[hot]]
    return ev


ONLY = [int(v) for v in os.environ.get("SWEEP_STRATEGIES", "").split(",") if v]


def main():
    rng = np.random.default_rng(0)
    rows = []
    for (W, H) in [(240, 180), (640, 480)]:
        for kind in ["uniform", "edge", "hot"]:
            for n in [10_000, 30_000, 100_000, 1_000_000, 10_000_000]:
                ev = hot_pixel_events(rng, n, H, W) if kind == "hot" else synth_events(rng, n, H, W, kind)
                d = torch.from_numpy(ev).cuda()
                for s in (0, 1, 2, 3, 4, 6, 7):
                    if s == 3 and n > 1_000_000 and W == 640:
                        continue
                    if s == 4 and W == 640:            # sensor does not fit a tile: PRIVATE is GLOBAL there
                        continue
                    if s in (6, 7) and (W != 640 or n < 1_000_000):   # large-sensor, long-stream strategies
                        continue
                    if ONLY and s not in ONLY:
                        continue
                    ms = timeit(lambda: histogram(d, H, W, strategy=s, check=False))
                    rows.append({"sensor": f"{W}x{H}", "kind": kind, "n": n, "strategy": NAMES[s], "ms": ms,
                                 "gev_s": n / ms / 1e6, "gb_s": (32 * n + 3 * H * W) / ms / 1e6})
                    print(rows[-1], flush=True)
    # ragged training batch: 128 streams x 30000 events
    for (W, H) in [(341, 256), (240, 180)]:
        B, per = 128, 30000
        ev = synth_events(rng, B * per, H, W, "edge", frac=True)
        off = torch.arange(B + 1, dtype=torch.int64, device="cuda") * per
        d = torch.from_numpy(ev).cuda()
        for C in (2, 3):
            for s in (1, 2, 3):
                ms = timeit(lambda: histogram_batch(d, off, H, W, channels=C, strategy=s, check=False,
                                                    max_stream_len=per))
                rows.append({"sensor": f"{W}x{H}", "kind": f"batch128x30k_C{C}", "n": B * per, "strategy": NAMES[s],
                             "ms": ms, "gev_s": B * per / ms / 1e6, "gb_s": (32 * B * per + C * H * W * B) / ms / 1e6})
                print(rows[-1], flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "hist_sweep.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
