"""Where does end-to-end step time go?  Host-side enqueue time per step (GPU never waited on), pinned H2D
bandwidth, and per-step wall times of train_one_epoch.  python tools/e2e_diag.py"""
import os, sys, time, io, contextlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from benchmarks import pretrain as bp
from mem_b200 import engine_for_pretraining, utils, _lib

dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
model, vae, opt = bp.build(torch, dev)
B = 128
batches = [bp.synth_batch(torch, B, i, dev) for i in range(4)]
devb = [(s, im, mk.to(dev)) for s, im, mk in batches]
hostb = [((s.cpu().pin_memory(), im.cpu().pin_memory(), mk.pin_memory()), None) for s, im, mk in batches]
step = bp.Step(torch, model, vae, opt, 1)
for i in range(4):
    step(*devb[i % 4])
torch.cuda.synchronize()
# host enqueue time: 3 steps back to back, no sync
t = []
for i in range(4):
    t0 = time.perf_counter(); l0 = _lib.launch_count(); step(*devb[i % 4]); t.append((time.perf_counter() - t0) * 1e3)
    nl = _lib.launch_count() - l0
torch.cuda.synchronize()
print("host enqueue ms per step (no sync):", [round(x, 2) for x in t], "libmemb launches/step", nl, "cpus", os.cpu_count(), flush=True)
# pinned H2D bandwidth
src = hostb[0][0][0]
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
dst = torch.empty_like(src, device=dev)
dst.copy_(src, non_blocking=True); torch.cuda.synchronize()
e0.record(); [dst.copy_(src, non_blocking=True) for _ in range(5)]; e1.record(); torch.cuda.synchronize()
print("pinned H2D GB/s:", round(5 * src.numel() * 4 / (e0.elapsed_time(e1) * 1e-3) / 1e9, 1), flush=True)
scaler = utils.NativeScalerWithGradNormCount()
for rep in range(3):
    loader = [hostb[i % 4] for i in range(10)]
    with contextlib.redirect_stdout(io.StringIO()):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        engine_for_pretraining.train_one_epoch(model, vae, loader, opt, dev, 0, scaler, 1.0)
        torch.cuda.synchronize(); dt = (time.perf_counter() - t0) * 1e3 / 10
    print(f"train_one_epoch rep {rep}: {dt:.2f} ms/step", flush=True)
