"""ORACLE (test infrastructure, not product): fp32 restatement of one MEM pretraining step.

Follows ``mem/engine_for_pretraining.py:108-287`` (``train_one_epoch``): per-step lr / weight-decay write
into the param groups (:124-130), tokens = dVAE ``get_codebook_indices`` (:144), labels = tokens[mask]
(:145), ``CrossEntropyLoss`` of the masked-token logits (:151-152), backward, global-norm clip
(``clip_grad_norm_``, mem/utils.py:362-364), AdamW step, ``mlm_acc`` (:233); the optimizer is the one
``mem/optim_factory.py`` builds for ``opt=adamw``: no weight decay on 1-D / ``.bias`` / ``no_weight_decay()``
parameters (:56-95), betas forced to (0.9, 0.95) (:121), eps 1e-8.

Pinned by ``tests/golden/engine_tiny.npz`` (three steps of the UNMODIFIED reference ``train_one_epoch``).
"""
from __future__ import annotations

import random

import torch

from . import dvae_ref, vit_ref
from .masking_ref import blockwise_mask_ref

TINY_VAE = dict(input_H=112, input_W=112, num_tokens=512, codebook_dim=16, num_layers=4, num_resnet_blocks=1,
                hidden_dim=32, channels=2)
LR = [1e-3, 2e-3, 1.5e-3]
WD = [0.05, 0.04, 0.03]
MAX_NORM = 1.0


def synth_batches(steps=3, B=4, seed=77):
    """[(samples, images, masks int64 [B,7,7])]: what the reference DataLoader yields (CreateTwoPic gives the
    same tensor twice, datasets.py:34-38)."""
    random.seed(seed)
    gen = lambda: blockwise_mask_ref((7, 7), 20, min_num_patches=4)  # noqa: E731
    out = []
    for s in range(steps):
        img = dvae_ref.synth_images(B, 2, 112, 112, seed + s)
        masks = torch.stack([torch.from_numpy(gen()).long() for _ in range(B)])
        out.append((img, img.clone(), masks))
    return out


def param_groups(named_params, weight_decay, skip=("pos_embed", "cls_token")):
    decay, no_decay = [], []
    for n, p in named_params:
        (no_decay if (p.dim() == 1 or n.endswith(".bias") or n in skip) else decay).append(p)
    return [{"params": no_decay, "weight_decay": 0.0, "lr_scale": 1.0}, {"params": decay, "weight_decay": weight_decay, "lr_scale": 1.0}]


def run_steps(vit_sd, vae_sd, batches, heads=2, patch=16, vae_layers=4, vae_res=1):
    """Returns per-step dicts {loss, mlm_acc, grad_norm} and the final ViT weights."""
    names = [k for k, v in vit_sd.items() if v.is_floating_point()]
    sd = {k: (v.clone().requires_grad_(True) if v.is_floating_point() else v) for k, v in vit_sd.items()}
    groups = [g for g in param_groups([(n, sd[n]) for n in names], WD[0]) if g["params"]]
    opt = torch.optim.AdamW(groups, lr=LR[0], betas=(0.9, 0.95), eps=1e-8)
    out = []
    for it, (samples, images, masks) in enumerate(batches):
        for g in opt.param_groups:
            g["lr"] = LR[it] * g["lr_scale"]
            if g["weight_decay"] > 0:
                g["weight_decay"] = WD[it]
        with torch.no_grad():
            tokens = dvae_ref.codebook_indices(images, vae_sd, vae_layers, vae_res)
        mask = masks.flatten(1).bool()
        loss, acc, _ = vit_ref.mem_loss(samples, mask, tokens, sd, heads, patch)
        opt.zero_grad()
        loss.backward()
        norm = torch.nn.utils.clip_grad_norm_([sd[n] for n in names], MAX_NORM)
        opt.step()
        out.append({"loss": loss.item(), "mlm_acc": acc.item(), "grad_norm": norm.item()})
    return out, {k: v.detach() for k, v in sd.items()}


# ---------------------------------------------------------------------------------------------
# Finetuning loop (mem/engine_for_finetuning.py:42-205): seeded case shared by make_golden.py and the tests
# ---------------------------------------------------------------------------------------------
FT_LR = [1e-3, 2e-3]
FT_WD = [0.05, 0.04]
FT_UPDATE_FREQ = 2


def synth_class_batches(n=4, B=6, seed=91):
    """[(samples float32 [B,3,112,112], targets int64 [B])]: ``n`` micro-batches (``n / FT_UPDATE_FREQ`` optimizer steps)."""
    out = []
    for s in range(n):
        img = dvae_ref.synth_images(B, 3, 112, 112, seed + s)
        g = torch.Generator().manual_seed(seed + 100 + s)
        out.append((img, torch.randint(0, 2, (B,), generator=g)))
    return out


def run_finetune(vit_sd, batches, heads=2, patch=16, update_freq=FT_UPDATE_FREQ):
    """The finetuning loop restated: CE on ft_vit logits, loss / update_freq, accumulate, clip + AdamW every
    ``update_freq`` micro-steps with the schedule value of the optimizer step (engine_for_finetuning.py:77-129).
    Returns (global-average stats like the reference's MetricLogger, final weights)."""
    names = [k for k, v in vit_sd.items() if v.is_floating_point()]
    sd = {k: (v.clone().requires_grad_(True) if v.is_floating_point() else v) for k, v in vit_sd.items()}
    groups = [g for g in param_groups([(n, sd[n]) for n in names], FT_WD[0]) if g["params"]]
    opt = torch.optim.AdamW(groups, lr=FT_LR[0], betas=(0.9, 0.95), eps=1e-8)
    opt.zero_grad()
    losses, accs, norms = [], [], []
    for i, (samples, targets) in enumerate(batches):
        it = i // update_freq
        for g in opt.param_groups:
            g["lr"] = FT_LR[it] * g["lr_scale"]
            if g["weight_decay"] > 0:
                g["weight_decay"] = FT_WD[it]
        logits = vit_ref.classify_logits(samples, sd, heads, patch)
        loss = torch.nn.functional.cross_entropy(logits, targets)
        losses.append(loss.item())
        accs.append((logits.max(-1)[1] == targets).float().mean().item())
        (loss / update_freq).backward()
        if (i + 1) % update_freq == 0:
            norms.append(torch.nn.utils.clip_grad_norm_([sd[n] for n in names], MAX_NORM).item())
            opt.step()
            opt.zero_grad()
    mean = lambda v: sum(v) / len(v)  # noqa: E731
    return {"loss": mean(losses), "class_acc": mean(accs), "grad_norm": mean(norms)}, {k: v.detach() for k, v in sd.items()}


# ---- timm 0.4.12 criteria the reference's finetuning runner picks (run_class_finetuning.py:551-559); timm is not installed,
# these are its published formulas (timm/loss/cross_entropy.py)
def label_smoothing_ce(x, target, smoothing=0.1):
    import torch.nn.functional as F
    logprobs = F.log_softmax(x, dim=-1)
    nll = -logprobs.gather(dim=-1, index=target.unsqueeze(1)).squeeze(1)
    return ((1.0 - smoothing) * nll + smoothing * (-logprobs.mean(dim=-1))).mean()


def soft_target_ce(x, target):
    import torch
    import torch.nn.functional as F
    return torch.sum(-target * F.log_softmax(x, dim=-1), dim=-1).mean()
