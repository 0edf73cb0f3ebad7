"""GPU: randomised cross-check of every rasteriser strategy against the oracle (bug hunting aid)."""
import sys, numpy as np, torch
sys.path.insert(0, "/root/repo")
from mem_b200.process_data import histogram
from oracle.histogram_ref import event_hist_ref
from oracle.make_golden import synth_events
bad_total = 0
for seed in range(6):
    for (H, W) in [(180, 240), (480, 640), (256, 341)]:
        rng = np.random.default_rng(seed * 1000 + H)
        for kind in ("uniform", "edge", "hot"):
            for n in (5_000, 20_000, 300_000, 1_200_000):
                ev = synth_events(rng, n, H, W, kind, frac=(kind == "edge"))
                want = event_hist_ref(ev, H, W)
                d = torch.from_numpy(ev).cuda()
                for s in (0, 1, 5, 6):
                    got = histogram(d, H, W, strategy=s).cpu().numpy()
                    bad = int((got != want).sum())
                    if bad:
                        bad_total += 1
                        rows = np.unique(np.argwhere(got != want)[:, 0])
                        print(f"MISMATCH seed={seed} {W}x{H} {kind} n={n} s={s}: {bad} values, rows {rows[:4]}..{rows[-4:]}, sums {int(got.sum())} vs {int(want.sum())}", flush=True)
print("done, mismatching configurations:", bad_total)
