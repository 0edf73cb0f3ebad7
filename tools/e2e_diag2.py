"""Phase timing inside train_one_epoch: wraps the calls it makes with host clocks and CUDA events."""
import os, sys, time, io, contextlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from benchmarks import pretrain as bp
from mem_b200 import engine_for_pretraining as eng, utils

dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
model, vae, opt = bp.build(torch, dev)
B = 128
batches = [bp.synth_batch(torch, B, i, dev) for i in range(4)]
hostb = [((s.cpu().pin_memory(), im.cpu().pin_memory(), mk.pin_memory()), None) for s, im, mk in batches]
scaler = utils.NativeScalerWithGradNormCount()
if "--prestep" in sys.argv:
    devb = [(s, im, mk.to(dev)) for s, im, mk in batches]
    stp = bp.Step(torch, model, vae, opt, 1)
    for i in range(6):
        stp(*devb[i % 4])
    torch.cuda.synchronize()
    print("prestep done; reserved GB", torch.cuda.memory_reserved() / 1e9)
NS = int(sys.argv[sys.argv.index('--steps') + 1]) if '--steps' in sys.argv else 8
log = []
def wrap(obj, name, tag):
    f = getattr(obj, name)
    def g(*a, **k):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter(); e0.record(); r = f(*a, **k); e1.record(); t1 = time.perf_counter()
        log.append((tag, t0, t1, e0, e1)); return r
    setattr(obj, name, g)
wrap(vae, "get_codebook_indices", "dvae")
wrap(eng, "pretrain_step", "vit")
wrap(eng._StepStats, "read", "read")
wrap(eng._PrefetchToDevice, "_load", "load")
wrap(opt, "step", "opt")
for rep in range(6):
    log.clear()
    loader = [hostb[i % 4] for i in range(NS)]
    with contextlib.redirect_stdout(io.StringIO()):
        torch.cuda.synchronize(); t00 = time.perf_counter()
        eng.train_one_epoch(model, vae, loader, opt, dev, 0, scaler, 1.0)
        torch.cuda.synchronize(); dt = (time.perf_counter() - t00) * 1e3 / NS
    print(f"rep {rep}: {dt:.2f} ms/step; reserved GB {torch.cuda.memory_reserved() / 1e9:.2f}")
    ends = [t1 for tag, t0, t1, e0, e1 in log if tag == "read"]
    print("   per-step ms:", [round(1e3 * (b - a), 1) for a, b in zip([t00] + ends[:-1], ends)])
    slow = [(tag, round(1e3*(t1-t0),2), round(e0.elapsed_time(e1),2)) for tag, t0, t1, e0, e1 in log if tag != "read" and (1e3*(t1-t0) > 6 or e0.elapsed_time(e1) > 24)]
    print("   slow phases:", slow)
    if rep == 99:
        for tag, t0, t1, e0, e1 in (log if "--quiet" not in sys.argv else [x for x in log if x[0] == "read"]):
            print(f"  {tag:5s} host start {1e3*(t0-t00):8.2f} dur {1e3*(t1-t0):7.2f} | gpu {e0.elapsed_time(e1):7.2f} ms")

# ---- statement-level timing of _load in this context
if "--loadprobe" in sys.argv:
    import types
    acc = {}
    def tick(label, t0):
        t1 = time.perf_counter(); acc.setdefault(label, []).append(round((t1 - t0) * 1e3, 3)); return t1
    def _load(self, item):
        t = time.perf_counter()
        batch, rest = item[0], item[1:]
        samples, images, mask = batch
        n_masked = int(mask.ne(0).sum()) if not mask.is_cuda else None
        t = tick("ne.sum", t)
        slot = self.slots[self.k % 2]; self.k += 1
        cur = torch.cuda.current_stream(self.device)
        dst = tuple(self._stage(slot, i, x) for i, x in enumerate((samples, images, mask)))
        t = tick("stage", t)
        free = torch.cuda.Event(); free.record(cur)
        t = tick("free.record", t)
        with torch.cuda.stream(self.side):
            self.side.wait_event(free)
            t = tick("wait_event", t)
            order = [2, 1, 0] if "--rev" in sys.argv else [0, 1, 2]
            srcs = (samples, images, mask)
            for i in order:
                dst[i].copy_(srcs[i], non_blocking=True)
                t = tick("copy", t)
            if "--again" in sys.argv:
                for i in order:
                    dst[i].copy_(srcs[i], non_blocking=True)
                    t = tick("copy", t)
            ev = torch.cuda.Event(); ev.record(self.side)
            t = tick("ev.record", t)
        return dst, n_masked, ev, rest
    eng._PrefetchToDevice._load = _load
    loader = [hostb[i % 4] for i in range(NS)]
    with contextlib.redirect_stdout(io.StringIO()):
        eng.train_one_epoch(model, vae, loader, opt, dev, 0, scaler, 1.0)
    for k, v in acc.items():
        print(k, v)
if "--pinprobe" in sys.argv:
    for bi, (b, _) in enumerate(hostb):
        print("batch", bi, [(tuple(x.shape), x.dtype, x.is_pinned(), x.is_contiguous(), x.stride()) for x in b])
    fresh = torch.empty(128, 2, 224, 224, device=dev)
    for j in range(3):
        for name, x in (("samples", hostb[0][0][0]), ("images", hostb[0][0][1])):
            torch.cuda.synchronize(); t0 = time.perf_counter(); fresh.copy_(x, non_blocking=True); t1 = time.perf_counter(); torch.cuda.synchronize()
            print(name, "copy_ host ms", round((t1 - t0) * 1e3, 3))
