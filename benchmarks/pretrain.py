"""bench.py workload: one ViT-B/16 MEM pretraining step per "step" (BASELINE.json config 3).

step = dVAE tokens (fp32-faithful) -> masked ViT forward -> 8192-way CE -> backward -> [DP: bucketed NCCL
all-reduce overlapped with backward] -> global-norm clip + AdamW, on batch 128 per GPU of synthetic
2x224x224 event histograms (rasterised on the GPU by the histogram kernel from synthetic event streams,
then normalised by the per-image maximum like the reference's NormalizeEvent, transforms.py:233-237).

`value`  : samples/s with the batch already resident in HBM (CUDA events, max over ranks).
`e2e`    : the same through engine_for_pretraining.train_one_epoch with a HOST (pinned) data loader:
           H2D copies of samples / images / masks and the D2H read of the step statistics are inside.
`roofline`: the dominant kernel of the step, `conv16::conv_f16x2` (the dVAE tokenizer's convolutions: 16 launches
           per step, ~40 % of the step's launch time, profiles/r02_pretrain_launch_shares_*.txt).  Every conv launch of
           the timed region is bracketed by CUDA events on the launching stream; achieved = algorithmic FLOPs
           (2*B*OH*OW*Cout*K per launch, 3.1 TFLOP per 128-image step) / summed launch time.  The kernel carries every
           fp32 operand as an fp16 hi/lo pair and issues three tensor-core MMAs per algorithmic product, so the
           roofline of its algorithmic FLOP rate is the measured peak / 3 (`peak`, `frac`); `frac_of_measured_peak` is the
           plain ratio to MEASURED_PEAKS.json and `executed_mma_tflops` the rate of MMA work actually issued.
`gemm_fc1`: the largest ViT GEMM launch (fc1: [B*197, 768] x [3072, 768]^T, bf16 tcgen05) timed alone, cold L2.
`step_tensor_util`: ViT bf16 FLOPs / whole step time / measured sustained bf16 peak (dVAE time is inside the
           step, its FLOPs are not counted: SURVEY.md 8d).
"""
from __future__ import annotations

import os
import random
import time

import numpy as np

CFG = dict(model="beit_base_patch16_224_8k_vocab", img=224, patch=16, in_chans=2, depth=12, dim=768, heads=12, mlp=4,
           vocab=8192, num_mask=75, min_mask=16, drop_path=0.1, init_values=0.1, lr=5e-4, wd=0.05, max_norm=1.0,
           events_per_sample=30000)
DVAE = dict(input_H=224, input_W=224, num_tokens=8192, codebook_dim=32, num_layers=4, num_resnet_blocks=3, hidden_dim=384,
            channels=2)
MODELS = {"base": dict(name="beit_base_patch16_224_8k_vocab", dim=768, depth=12, heads=12, init_values=0.1),
          "large": dict(name="beit_large_patch16_224_8k_vocab", dim=1024, depth=24, heads=16, init_values=1e-5)}


def vit_flops_per_sample(dim, depth, heads, n_masked, C=2, N=197, vocab=8192, patch=16, mlp=4):
    """fwd+bwd (3x fwd) FLOPs of the masked ViT step, SURVEY.md 8(d)."""
    blk = 2 * N * dim * 3 * dim + 2 * 2 * heads * N * N * (dim // heads) + 2 * N * dim * dim + 2 * 2 * N * dim * mlp * dim
    fwd = depth * blk + 2 * n_masked * dim * vocab + 2 * (N - 1) * (C * patch * patch) * dim
    return 3.0 * fwd


def synth_batch(torch, B, seed, device, cfg=CFG):
    """(samples, images, masks int64 [B,14,14]) on `device`: event streams -> histogram kernel -> /max."""
    from mem_b200.masking_generator import MaskingGenerator
    from mem_b200.process_data import histogram_batch
    H = W = cfg["img"]
    n = cfg["events_per_sample"]
    g = torch.Generator(device=device).manual_seed(seed)
    ev = torch.empty(B * n, 4, dtype=torch.float64, device=device)
    # edge-like streams: events clustered on a few line segments + uniform noise
    seg = torch.randint(0, 12, (B * n,), generator=g, device=device)
    base = torch.rand(B, 12, 4, generator=g, device=device, dtype=torch.float64)
    sample = torch.arange(B, device=device).repeat_interleave(n)
    p0 = base[sample, seg]
    s = torch.rand(B * n, generator=g, device=device, dtype=torch.float64)
    x = (p0[:, 0] + (p0[:, 2] - 0.5) * s) * W + torch.randn(B * n, generator=g, device=device, dtype=torch.float64)
    y = (p0[:, 1] + (p0[:, 3] - 0.5) * s) * H + torch.randn(B * n, generator=g, device=device, dtype=torch.float64)
    noise = torch.rand(B * n, generator=g, device=device) < 0.2
    x = torch.where(noise, torch.rand(B * n, generator=g, device=device, dtype=torch.float64) * W, x)
    y = torch.where(noise, torch.rand(B * n, generator=g, device=device, dtype=torch.float64) * H, y)
    ev[:, 0] = x.clamp_(0, W - 1).floor_()
    ev[:, 1] = y.clamp_(0, H - 1).floor_()
    ev[:, 2] = torch.rand(B * n, generator=g, device=device, dtype=torch.float64) * 3e5
    ev[:, 3] = torch.randint(0, 2, (B * n,), generator=g, device=device).double() * 2 - 1
    offsets = torch.arange(B + 1, device=device, dtype=torch.int64) * n
    hist = histogram_batch(ev, offsets, H, W, channels=cfg["in_chans"], max_stream_len=n, check=False)   # uint8 [B,H,W,2]
    img = hist.permute(0, 3, 1, 2).float()
    img = img / img.amax(dim=(1, 2, 3), keepdim=True).clamp_min(1.0)
    random.seed(seed)
    gen = MaskingGenerator((H // cfg["patch"],) * 2, cfg["num_mask"], min_num_patches=cfg["min_mask"])
    masks = torch.from_numpy(np.stack([gen() for _ in range(B)])).long()
    img = img.contiguous()   # NCHW-contiguous, what a DataLoader's default collate hands the engine
    return img, img.clone(), masks


class Step:
    """The device-resident step, exactly what train_one_epoch does between the H2D copies and the stats read."""

    def __init__(self, torch, model, vae, opt, world):
        from mem_b200.engine_for_pretraining import _reducer_for
        from mem_b200.vit_engine import pretrain_step
        self.torch, self.model, self.vae, self.opt = torch, model, vae, opt
        self.pretrain_step = pretrain_step
        self.reducer = _reducer_for(model) if world > 1 else None
        opt.grad_divisor = float(world)
        self.ev = None

    def __call__(self, samples, images, masks, marks=None):
        torch = self.torch
        if marks is not None:
            marks[0].record()
        tokens = self.vae.get_codebook_indices(images)
        if marks is not None:
            marks[1].record()
        self.opt.zero_grad()
        stats = self.pretrain_step(self.model, samples, masks.flatten(1), tokens, cap=samples.shape[0] * CFG["num_mask"],
                                   bucket_hook=self.reducer.hook if self.reducer else None)
        if self.reducer:
            self.reducer.finish()
        if marks is not None:
            marks[2].record()
        self.opt.step(max_norm=CFG["max_norm"])
        if marks is not None:
            marks[3].record()
        return stats


def build(torch, device, size="base", drop_path=None):
    from mem_b200 import modeling_pretrain, optim_factory, registry  # noqa: F401
    from mem_b200.vae_model import DiscreteVAE
    from types import SimpleNamespace
    m = MODELS[size]
    torch.manual_seed(0)
    model = registry.create_model(m["name"], pretrained=False, drop_path_rate=CFG["drop_path"] if drop_path is None else drop_path,
                                  drop_block_rate=None, use_shared_rel_pos_bias=True, use_abs_pos_emb=False,
                                  init_values=m["init_values"], in_chans=CFG["in_chans"]).to(device)
    vae = DiscreteVAE(**DVAE).to(device)
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        opt = optim_factory.create_optimizer(SimpleNamespace(opt="adamw", weight_decay=CFG["wd"], lr=CFG["lr"], opt_eps=1e-8), model)
    model.train()
    return model, vae, opt


def main(args, rank, local_rank, world, ClockSampler, measured_peaks):
    if args.impl == "reference":
        return reference_arm(args, rank)
    import torch
    import torch.distributed as dist
    assert torch.cuda.is_available(), "bench.py (our arm) needs a GPU; there is no CPU fallback"
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    from mem_b200 import _lib, engine_for_pretraining, utils
    steps = args.steps or 20
    warm = max(args.warmup if args.warmup is not None else 5, 3)
    B = args.batch
    size = getattr(args, "model", "base")
    m = MODELS[size]
    model, vae, opt = build(torch, device, size)
    n_batches = 4
    batches = [synth_batch(torch, B, 1000 * rank + i, device) for i in range(n_batches)]
    dev_batches = [(s, im, mk.to(device)) for s, im, mk in batches]
    host_batches = [((s.cpu().pin_memory(), im.cpu().pin_memory(), mk.pin_memory()), None) for s, im, mk in batches]
    step = Step(torch, model, vae, opt, world)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def maxreduce(ms):
        if world > 1:
            t = torch.tensor([ms], device=device, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return ms

    for i in range(warm):
        step(*dev_batches[i % n_batches])
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    # ---- device-resident timed region
    marks = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(steps)]
    conv_launches = []
    vae._tokenizer().launch_timer = conv_launches       # (layer, FLOPs, start, end) per conv_f16x2 launch
    l0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(steps):
        stats = step(*dev_batches[i % n_batches], marks=marks[i])
    e1.record()
    barrier()
    launches = _lib.launch_count() - l0
    vae._tokenizer().launch_timer = None
    ms_step = maxreduce(e0.elapsed_time(e1)) / steps
    seg = np.array([[mk[j].elapsed_time(mk[j + 1]) for j in range(3)] for mk in marks]).mean(0)
    n_masked = float(stats[2].item()) / B
    loss_last = float(stats[0].item() / max(stats[2].item(), 1.0))
    # ---- end to end through the public API (host batches, pinned)
    scaler = utils.NativeScalerWithGradNormCount()
    import contextlib
    import io
    e2e_steps = max(4, min(steps, 12))
    loader = [host_batches[i % n_batches] for i in range(e2e_steps)]
    with contextlib.redirect_stdout(io.StringIO()):
        engine_for_pretraining.train_one_epoch(model, vae, loader[:3], opt, device, 0, scaler, CFG["max_norm"])
        barrier()
        e0.record()
        out_stats = engine_for_pretraining.train_one_epoch(model, vae, loader, opt, device, 0, scaler, CFG["max_norm"])
        e1.record()
        barrier()
    ms_e2e = maxreduce(e0.elapsed_time(e1)) / e2e_steps
    clocks = sampler.stop() if rank == 0 else None
    # ---- dominant kernel (tokenizer convolutions, timed inside the steps above) and the largest ViT GEMM alone
    roof = conv_roofline(conv_launches, steps, measured_peaks) if rank == 0 else None
    fc1 = gemm_roofline(torch, model, B, m, measured_peaks) if rank == 0 else None
    if world > 1:
        dist.barrier()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    hbm, tf_burst, tf_sust, how = measured_peaks()
    flops = vit_flops_per_sample(m["dim"], m["depth"], m["heads"], n_masked)
    value = B * world / (ms_step * 1e-3)
    h2d = sum(t.numel() * t.element_size() for t in host_batches[0][0])
    line = {"metric": "ViT-B/16 MEM pretrain samples/s" if size == "base" else "ViT-L/16 MEM pretrain samples/s",
            "value": round(value, 1), "unit": "samples/s", "n_gpus": world, "steps": steps, "warmup": warm,
            "ms_per_step": round(ms_step, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"{m['name']} MEM pretraining step: random-init dVAE tokenizer (hidden 384, 3 res blocks, 8192 "
                                   f"tokens, fp32-faithful), {CFG['num_mask']} blockwise-masked patches, CE over 8192, AdamW, clip "
                                   f"{CFG['max_norm']}, drop_path {CFG['drop_path']}; synthetic 2x224x224 event histograms",
                       "batch_per_gpu": B, "global_batch": B * world, "parallelism": f"dp{world}",
                       "l2_policy": f"inputs rotate over {n_batches} resident batches; per-step activations (>3 GB) exceed the 126 MB L2"},
            "e2e": {"value": round(B * world / (ms_e2e * 1e-3), 1), "unit": "samples/s", "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": 16, "ms_per_step": round(ms_e2e, 3), "api": "engine_for_pretraining.train_one_epoch"},
            "gpu_launches": int(launches),
            "breakdown_ms": {"dvae_tokens": round(float(seg[0]), 3), "vit_fwd_ce_bwd_allreduce": round(float(seg[1]), 3),
                             "clip_adamw": round(float(seg[2]), 3)},
            "step_tensor_util": {"vit_bf16_gflop_per_sample": round(flops / 1e9, 2), "achieved_tflops": round(flops * B / (ms_step * 1e-3) / 1e12, 1),
                                 "peak_tflops_sustained": tf_sust, "frac": round(flops * B / (ms_step * 1e-3) / 1e12 / tf_sust, 4)},
            "masked_per_sample": round(n_masked, 2), "loss_last": round(loss_last, 4),
            "e2e_stats": {k: round(float(v), 5) for k, v in out_stats.items()},
            "clocks": clocks, "roofline": roof, "gemm_fc1": fc1}
    if not args.no_cpu_baseline and world == 1:
        line["cpu_baseline"] = cpu_baseline(sample_steps=2)
    if world > 1:
        dist.destroy_process_group()
    return line


def conv_roofline(conv_launches, steps, measured_peaks):
    """Roofline record of conv16::conv_f16x2 from the per-launch CUDA events of the timed region."""
    hbm, tf_burst, tf_sust, how = measured_peaks()
    per_layer = {}
    for layer, fl, e0, e1 in conv_launches:
        acc = per_layer.setdefault(layer, [0.0, 0.0, 0])
        acc[0] += fl; acc[1] += e0.elapsed_time(e1); acc[2] += 1
    n = sum(v[2] for v in per_layer.values())
    if n == 0:
        return None
    fl_total = sum(v[0] for v in per_layer.values())
    ms_total = sum(v[1] for v in per_layer.values())
    ach = fl_total / (ms_total * 1e-3) / 1e12
    layers = {k: {"launches_per_step": v[2] // steps, "us": round(v[1] / v[2] * 1e3, 1),
                  "tflops": round(v[0] / (v[1] * 1e-3) / 1e12, 1)} for k, v in per_layer.items()}
    # The kernel issues three tensor-core MMAs per algorithmic product (lo*hi + hi*lo + hi*hi keep 22 significand bits per
    # operand: the fp32-faithful tokens north_star asks for), so the roofline that bounds its ALGORITHMIC FLOP rate is the
    # measured tensor peak / 3 ("frac vs peak / 3"); the ratio to the plain measured peak is kept beside it.
    return {"bound": "tensor", "kernel": "conv16::conv_f16x2 (dVAE tokenizer convolutions: implicit GEMM on CTA-pair tcgen05, fp16 hi/lo "
                                         "operand pairs, 3 MMAs per fp32-faithful product)",
            "achieved": round(ach, 1), "peak": round(tf_sust / 3.0, 1), "unit": "TFLOP/s", "frac": round(3.0 * ach / tf_sust, 4),
            "measured_peak": tf_sust, "frac_of_measured_peak": round(ach / tf_sust, 4),
            "executed_mma_tflops": round(3.0 * ach, 1),
            "peak_source": how + " (MEASURED_PEAKS.json bf16_tflops_sustained / 3: the kernel is timed inside the long step; fp16 and "
                                 "bf16 share the tensor-pipe rate; three MMAs per algorithmic product)",
            "scheme_ceiling": "peak / 3: lo*hi + hi*lo + hi*hi per product (22 significand bits per operand)",
            "algorithmic_flops_per_launch": fl_total / n, "launch_ms": round(ms_total / n, 4),
            "launches_per_step": n // steps, "step_share_ms": round(ms_total / steps, 3),
            "timing": "CUDA events around every conv_f16x2 launch of the timed region, on the launching stream",
            "per_layer": layers, "traffic": _traffic("conv_f16x2_B128_mean"),
            "traffic_note": "mean DRAM read+write bytes per launch over the 16 launches of one 128-image tokenizer pass, "
                            "ncu --set full capture (profiles/ncu_traffic.json names the commit)"}


def gemm_roofline(torch, model, B, m, measured_peaks):
    """Time the dominant GEMM launch (fc1 forward, bias+GELU epilogue) alone, cold L2 between launches."""
    from mem_b200 import ops
    from mem_b200._lib import EPI_BIAS_GELU
    from mem_b200.vit_engine import engine_of
    hbm, tf_burst, tf_sust, how = measured_peaks()
    dev = next(model.parameters()).device
    M, K, N = B * 197, m["dim"], 4 * m["dim"]
    flat = engine_of(model).flat()
    a = torch.randn(M, K, device=dev).bfloat16()
    w = flat.w16("blocks.0.mlp.fc1.weight")
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    pre = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    bias = model.blocks[0].mlp.fc1.bias
    for _ in range(3):
        ops.gemm(a, w, out=out, epilogue=EPI_BIAS_GELU, bias=bias, d2=pre)
    times = []
    for _ in range(10):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.gemm(a, w, out=out, epilogue=EPI_BIAS_GELU, bias=bias, d2=pre)
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    ms = float(np.mean(times))
    fl = 2.0 * M * N * K
    ach = fl / (ms * 1e-3) / 1e12
    return {"bound": "tensor", "kernel": f"gemm_pair_kernel<256, BIAS_GELU> (CTA-pair tcgen05, bf16) fc1 [{M}x{K}]x[{N}x{K}]^T", "achieved": round(ach, 1),
            "peak": tf_burst, "unit": "TFLOP/s", "frac": round(ach / tf_burst, 4), "peak_source": how + " (MEASURED_PEAKS.json bf16_tflops, burst)",
            "algorithmic_flops_per_launch": fl, "launch_ms": round(ms, 4), "l2": "flushed (256 MB memset) before every timed launch",
            "traffic": _traffic("gemm_fc1_bias_gelu_B128") if (M, N, K) == (128 * 197, 3072, 768) else None}


def _traffic(key):
    """DRAM bytes per launch from the committed `ncu --set full` capture (profiles/ncu_traffic.json)."""
    import json
    try:
        root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        return json.load(open(os.path.join(root, "profiles", "ncu_traffic.json")))[key]["bytes"]
    except Exception:
        return None


# ----------------------------------------------------------------------------- CPU arms (oracle port)
def _cpu_step_fn(B):
    """The reference step restated in fp32 PyTorch on the host cores (oracle/engine_ref.py semantics at full size)."""
    import torch
    from mem_b200 import modeling_pretrain, registry  # noqa: F401
    from mem_b200.vae_model import DiscreteVAE
    from oracle import dvae_ref, engine_ref, vit_ref
    torch.manual_seed(0)
    model = registry.create_model("beit_base_patch16_224_8k_vocab", use_shared_rel_pos_bias=True, use_abs_pos_emb=False,
                                  init_values=0.1, in_chans=2)
    vit_sd = {k: (v.detach().clone().requires_grad_(True) if v.is_floating_point() else v) for k, v in model.state_dict().items()}
    vae_sd = {k: v.detach() for k, v in DiscreteVAE(**DVAE).state_dict().items()}
    names = [k for k, v in vit_sd.items() if v.is_floating_point()]
    groups = [g for g in engine_ref.param_groups([(n, vit_sd[n]) for n in names], CFG["wd"]) if g["params"]]
    opt = torch.optim.AdamW(groups, lr=CFG["lr"], betas=(0.9, 0.95), eps=1e-8)
    img = dvae_ref.synth_images(B, 2, 224, 224, seed=3)
    g = torch.Generator().manual_seed(1)
    mask = torch.zeros(B, 196, dtype=torch.bool)
    for b in range(B):
        mask[b, torch.randperm(196, generator=g)[:75]] = True

    def step():
        with torch.no_grad():
            tokens = dvae_ref.codebook_indices(img, vae_sd, 4, 3)
        loss, acc, _ = vit_ref.mem_loss(img, mask, tokens, vit_sd, 12, 16)
        opt.zero_grad()
        loss.backward()
        torch.nn.utils.clip_grad_norm_([vit_sd[n] for n in names], CFG["max_norm"])
        opt.step()
        return loss.item()
    return step


def cpu_baseline(sample_steps=2, B=8):
    import torch
    step = _cpu_step_fn(B)
    step()
    t0 = time.perf_counter()
    for _ in range(sample_steps):
        step()
    dt = (time.perf_counter() - t0) / sample_steps
    return {"value": round(B / dt, 3), "unit": "samples/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{sample_steps} steps of batch {B} (ViT-B/16 + dVAE tokenizer + AdamW, fp32) through oracle/engine_ref.py "
                      f"semantics on torch CPU kernels ({torch.get_num_threads()} threads)"}


def reference_arm(args, rank):
    if rank != 0:
        return
    import torch
    torch.set_num_threads(os.cpu_count() or 1)
    B = 8
    steps = args.steps or 3
    warm = args.warmup if args.warmup is not None else 1
    steps, warm = min(steps, 6), min(max(warm, 1), 2)       # bounded sample: ~2 s of CPU per step
    step = _cpu_step_fn(B)
    for _ in range(warm):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    value = B / dt
    sample = (f"{steps} steps of batch {B} per step (the reference's CPU-runnable config 1 shape) of the same ViT-B/16 MEM step, fp32, "
              f"torch CPU kernels on {torch.get_num_threads()} threads; oracle port of engine_for_pretraining.train_one_epoch")
    line = {"impl": "reference", "metric": "ViT-B/16 MEM pretrain samples/s", "value": round(value, 3), "unit": "samples/s",
            "n_gpus": args.gpus, "steps": steps, "warmup": warm, "ms_per_step": round(dt * 1e3, 1), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "beit_base_patch16_224_8k_vocab MEM pretraining step (CPU sample: batch 8 per step)", "batch_per_gpu": B},
            "cpu_baseline": {"value": round(value, 3), "unit": "samples/s", "cores": torch.get_num_threads(), "kind": "port", "sample": sample},
            "e2e": {"value": round(value, 3), "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    return line
