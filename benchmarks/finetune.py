"""bench.py workload `finetune`: ViT-B/16 classification step on N-Cars-shaped 2-class synthetic histograms
(BASELINE.json config 5: modeling_finetune path, no masking).

step = ft_vit forward (mean-pool head) -> CrossEntropyLoss -> backward -> global-norm clip + AdamW, batch 128 per GPU of
synthetic 3x224x224 event histograms.  `value`: batch resident in HBM (CUDA events); `e2e`: through
engine_for_finetuning.train_one_epoch with pinned host batches (H2D of samples and targets, D2H of the loss).
"""
from __future__ import annotations

import json
import os
import time

import numpy as np

from . import pretrain as bp

FT = dict(img_size=(224, 224), patch_size=(16, 16), in_chans=3, num_classes=2, embed_dim=768, depth=12, num_heads=12,
          mlp_ratio=4, init_values=0.1, use_rel_pos_bias=True, use_abs_pos_emb=False, use_mean_pooling=True, drop_path_rate=0.1)
METRIC = "ViT-B/16 finetune samples/s (ft_vit, 2 classes)"


def flops_per_sample(dim=768, depth=12, heads=12, C=3, N=197, patch=16, mlp=4, classes=2):
    blk = 2 * N * dim * 3 * dim + 2 * 2 * heads * N * N * (dim // heads) + 2 * N * dim * dim + 2 * 2 * N * dim * mlp * dim
    return 3.0 * (depth * blk + 2 * (N - 1) * (C * patch * patch) * dim + 2 * dim * classes)


def _cpu_step_fn(B):
    import torch
    from mem_b200 import modeling_finetune, registry  # noqa: F401
    from oracle import dvae_ref, engine_ref, vit_ref
    torch.manual_seed(0)
    model = registry.create_model("ft_vit", **FT)
    sd = {k: (v.detach().clone().requires_grad_(True) if v.is_floating_point() else v) for k, v in model.state_dict().items()}
    names = [k for k, v in sd.items() if v.is_floating_point()]
    groups = [g for g in engine_ref.param_groups([(n, sd[n]) for n in names], 0.05) if g["params"]]
    opt = torch.optim.AdamW(groups, lr=5e-4, betas=(0.9, 0.95), eps=1e-8)
    img = dvae_ref.synth_images(B, 3, 224, 224, seed=3)
    tgt = torch.randint(0, 2, (B,), generator=torch.Generator().manual_seed(1))

    def step():
        loss = torch.nn.functional.cross_entropy(vit_ref.classify_logits(img, sd, 12, 16), tgt)
        opt.zero_grad()
        loss.backward()
        torch.nn.utils.clip_grad_norm_([sd[n] for n in names], 1.0)
        opt.step()
        return loss.item()
    return step


def _cpu_line(steps, warm, B=8):
    import torch
    step = _cpu_step_fn(B)
    for _ in range(warm):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    return B / dt, dt, {"value": round(B / dt, 3), "unit": "samples/s", "cores": torch.get_num_threads(), "kind": "port",
                        "sample": f"{steps} steps of batch {B} (ft_vit ViT-B/16 forward + CE + backward + AdamW, fp32) through "
                                  f"oracle/vit_ref.py on torch CPU kernels ({torch.get_num_threads()} threads)"}


def main(args, rank, local_rank, world, ClockSampler, measured_peaks):
    if args.impl == "reference":
        if rank != 0:
            return
        import torch
        torch.set_num_threads(os.cpu_count() or 1)
        steps, warm = min(args.steps or 3, 6), min(max(args.warmup if args.warmup is not None else 1, 1), 2)
        value, dt, cb = _cpu_line(steps, warm)
        print(json.dumps({"impl": "reference", "metric": METRIC, "value": round(value, 3), "unit": "samples/s", "n_gpus": args.gpus,
                          "steps": steps, "warmup": warm, "ms_per_step": round(dt * 1e3, 1), "higher_is_better": True,
                          "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                          "config": {"workload": "ft_vit ViT-B/16 classification step (CPU sample: batch 8 per step)", "batch_per_gpu": 8},
                          "cpu_baseline": cb, "e2e": {"value": round(value, 3), "unit": "samples/s", "h2d_bytes_per_step": 0,
                                                      "d2h_bytes_per_step": 0}, "gpu_launches": 0}))
        return
    import torch
    import torch.distributed as dist
    assert torch.cuda.is_available(), "bench.py (our arm) needs a GPU; there is no CPU fallback"
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    from types import SimpleNamespace
    import contextlib
    import io
    from mem_b200 import _lib, engine_for_finetuning as eft, modeling_finetune, optim_factory, registry, utils  # noqa: F401
    from mem_b200.vit_engine import engine_of
    steps = args.steps or 20
    warm = max(args.warmup if args.warmup is not None else 5, 3)
    B = args.batch
    torch.manual_seed(0)
    model = registry.create_model("ft_vit", **FT).to(device)
    with contextlib.redirect_stdout(io.StringIO()):
        opt = optim_factory.create_optimizer(SimpleNamespace(opt="adamw", weight_decay=0.05, lr=5e-4, opt_eps=1e-8), model)
    model.train()
    cfg = dict(bp.CFG, in_chans=3)
    batches = []
    for i in range(4):
        img, _, _ = bp.synth_batch(torch, B, 2000 * rank + i, device, cfg)
        tgt = torch.randint(0, 2, (B,), device=device, generator=torch.Generator(device=device).manual_seed(i))
        batches.append((img, tgt))
    host = [(x.cpu().pin_memory(), t.cpu().pin_memory()) for x, t in batches]
    crit = torch.nn.CrossEntropyLoss()
    opt.grad_divisor = float(world)

    def step(x, t):
        loss = crit(model(x), t)
        loss.backward()
        if world > 1:
            dist.all_reduce(engine_of(model).flat().grad)
        opt.step(max_norm=1.0)
        opt.zero_grad()
        return loss

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

    opt.zero_grad()
    for i in range(warm):
        step(*batches[i % 4])
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    l0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(steps):
        loss = step(*batches[i % 4])
    e1.record()
    barrier()
    launches = _lib.launch_count() - l0
    ms_step = maxreduce(e0.elapsed_time(e1)) / steps
    scaler = utils.NativeScalerWithGradNormCount()
    e2e_steps = max(4, min(steps, 12))
    loader = [host[i % 4] for i in range(e2e_steps)]
    with contextlib.redirect_stdout(io.StringIO()):
        eft.train_one_epoch(None, model, crit, loader[:3], opt, device, 0, scaler, 1.0)
        barrier()
        e0.record()
        out_stats = eft.train_one_epoch(None, model, crit, loader, opt, device, 0, scaler, 1.0)
        e1.record()
        barrier()
    ms_e2e = maxreduce(e0.elapsed_time(e1)) / e2e_steps
    clocks = sampler.stop() if rank == 0 else None
    roof = bp.gemm_roofline(torch, model, B, dict(dim=768), measured_peaks) if rank == 0 else None
    if world > 1:
        dist.barrier()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    _, _, tf_sust, _ = measured_peaks()
    fl = flops_per_sample()
    line = {"metric": METRIC, "value": round(B * world / (ms_step * 1e-3), 1), "unit": "samples/s", "n_gpus": world, "steps": steps,
            "warmup": warm, "ms_per_step": round(ms_step, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "ft_vit ViT-B/16 classification step (BASELINE config 5): per-block rel-pos bias, mean-pool head, "
                                   "2 classes, CE, AdamW, clip 1.0, drop_path 0.1; synthetic 3x224x224 event histograms",
                       "batch_per_gpu": B, "global_batch": B * world, "parallelism": f"dp{world}",
                       "l2_policy": "inputs rotate over 4 resident batches; per-step activations (>3 GB) exceed the 126 MB L2"},
            "e2e": {"value": round(B * world / (ms_e2e * 1e-3), 1), "unit": "samples/s",
                    "h2d_bytes_per_step": int(sum(t.numel() * t.element_size() for t in host[0])), "d2h_bytes_per_step": 4,
                    "ms_per_step": round(ms_e2e, 3), "api": "engine_for_finetuning.train_one_epoch"},
            "gpu_launches": int(launches),
            "step_tensor_util": {"vit_bf16_gflop_per_sample": round(fl / 1e9, 2), "achieved_tflops": round(fl * B / (ms_step * 1e-3) / 1e12, 1),
                                 "peak_tflops_sustained": tf_sust, "frac": round(fl * B / (ms_step * 1e-3) / 1e12 / tf_sust, 4)},
            "loss_last": round(float(loss.item()), 4), "e2e_stats": {k: round(float(v), 5) for k, v in out_stats.items()},
            "clocks": clocks, "roofline": roof}
    if not args.no_cpu_baseline and world == 1:
        line["cpu_baseline"] = _cpu_line(2, 1)[2]
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
