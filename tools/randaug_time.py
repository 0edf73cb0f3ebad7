"""GPU: time memb_event_randaug on one training batch (128 x [3,224,224] float32 in / out, two operations per sample drawn
like the reference's EventRandAugment(magnitude=20)); per-operation timing with the whole batch on one operation."""
import contextlib, io, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mem_b200 import transforms as T
from oracle.make_golden import synth_event_image

B, H, W = 128, 224, 224
rng = np.random.default_rng(0)
x = torch.from_numpy(np.stack([synth_event_image(rng, H, W) for _ in range(B)]).astype(np.float32) / np.float32(255)).cuda()
with contextlib.redirect_stdout(io.StringIO()):
    aug = T.EventRandAugment(small=False, magnitude=20)


def timed(ops, iters=20):
    for _ in range(3):
        T.apply_ops(x, ops, out_float=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        T.apply_ops(x, ops, out_float=True)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


torch.manual_seed(0)
ops = aug.draw_batch(B, H, W)
us = timed(ops)
print(f"EventRandAugment batch {B} x 3x{H}x{W}, 2 drawn ops per sample: {us:.1f} us per call "
      f"({B * 3 * H * W * 8 / us / 1e3:.0f} GB/s of float32 in + out)")
for name in T.ALL:
    mag = {"Posterize": 5.0, "Solarize": 128.0, "Rotate": 14.0, "TranslateX": 30.0, "TranslateY": 30.0}.get(name, 0.21)
    one = np.zeros((B, 1), dtype=T.OP_DTYPE)
    one[:, 0] = T.encode_op(name, mag)
    print(f"  {name:13s} {timed(one, 10):7.1f} us")
