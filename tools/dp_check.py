"""GPU box, torchrun --nproc-per-node N: data-parallel parity.  Every rank runs its shard of a global batch through
train_one_epoch (bucketed NCCL all-reduce overlapped with backward + FlatAdamW) and, separately, the whole global
batch through a single-process copy of the same model; after one step the two weight sets must agree (the DP step
averages per-rank mean losses like DDP, so shards are built with equal masked counts per rank)."""
import contextlib
import io
import os
import sys
from types import SimpleNamespace

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mem_b200 import engine_for_pretraining, optim_factory, registry, utils  # noqa: E402
from mem_b200 import modeling_pretrain  # noqa: F401,E402
from mem_b200.vae_model import DiscreteVAE  # noqa: E402
from oracle import dvae_ref, engine_ref, vit_ref  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dev = torch.device("cuda", local)


def build():
    torch.manual_seed(0)
    cfg = dict(vit_ref.TINY, depth=4)
    m = registry.create_model("pt_vit", **cfg)
    m.load_state_dict(vit_ref.synth_state_dict(m.state_dict(), seed=31))
    v = DiscreteVAE(**engine_ref.TINY_VAE)
    v.load_state_dict(dvae_ref.synth_state_dict(v.state_dict(), seed=32, head_gain=4.0))
    m.to(dev), v.to(dev)
    with contextlib.redirect_stdout(io.StringIO()):
        o = optim_factory.create_optimizer(SimpleNamespace(opt="adamw", weight_decay=0.05, lr=1e-3, opt_eps=1e-8), m)
    return m, v, o


Bper = 4
img = dvae_ref.synth_images(Bper * world, 2, 112, 112, seed=5)
g = torch.Generator().manual_seed(1)
mask = torch.zeros(Bper * world, 49, dtype=torch.long)
for b in range(Bper * world):
    mask[b, torch.randperm(49, generator=g)[:18]] = 1          # equal masked count per sample -> per rank
mask = mask.view(-1, 7, 7)
scaler = utils.NativeScalerWithGradNormCount()
# --- DP run
m_dp, vae, o_dp = build()
sl = slice(rank * Bper, (rank + 1) * Bper)
with contextlib.redirect_stdout(io.StringIO()):
    s_dp = engine_for_pretraining.train_one_epoch(m_dp, vae, [((img[sl], img[sl], mask[sl]), None)], o_dp, dev, 0, scaler, 1.0)
# --- single-process run on the global batch (world size hidden from the engine)
m_1, vae1, o_1 = build()
real = utils.get_world_size
utils.get_world_size = lambda: 1
with contextlib.redirect_stdout(io.StringIO()):
    s_1 = engine_for_pretraining.train_one_epoch(m_1, vae1, [((img, img, mask), None)], o_1, dev, 0, scaler, 1.0)
utils.get_world_size = real
worst, off, total = 0.0, 0, 0
for (n, a), (_, b) in zip(m_dp.state_dict().items(), m_1.state_dict().items()):
    if a.is_floating_point():
        d = (a - b).abs()
        worst = max(worst, d.max().item())
        off += int((d > 2e-4).sum())
        total += d.numel()
red = m_dp._memb_reducer
wire = str(red.wire_dtype).replace("torch.", "")
# fp32 wire: every weight agrees.  bf16 wire: each rank's gradient is rounded before the sum, so an element whose per-rank
# gradients nearly cancel can change sign, and the first AdamW step moves it by lr the other way (2 * lr apart): such elements
# must stay a small fraction, the global statistics must agree.
weights_ok = worst < 2e-4 if red.wire_dtype == torch.float32 else (off <= 2e-3 * total and worst <= 2.5e-3)
ok = weights_ok and abs(s_dp["grad_norm"] - s_1["grad_norm"]) < 2e-2 * s_1["grad_norm"] and abs(s_dp["loss"] - s_1["loss"]) < 1e-3
print(f"[rank {rank}] wire {wire}; dp loss {s_dp['loss']:.5f} vs single {s_1['loss']:.5f}; grad_norm {s_dp['grad_norm']:.4f} vs {s_1['grad_norm']:.4f}; "
      f"max weight diff after 1 step {worst:.2e}, {off} of {total} elements beyond 2e-4; buckets launched {red.launched}; "
      f"{'OK' if ok else 'MISMATCH'}", flush=True)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
