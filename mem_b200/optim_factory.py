"""Optimizer factory of the MEM pretraining step.

Drop-in for ``mem/optim_factory.py``: ``get_parameter_groups`` (:56-95: no weight decay for 1-D tensors,
``.bias`` and the model's ``no_weight_decay()`` names; every group carries ``lr_scale``) and
``create_optimizer`` (:98-181), which -- whatever ``args.opt_betas`` says -- forces
``betas = (0.9, 0.95)`` (:121).  The MEM configs use ``opt = adamw``; that is the optimizer built here,
as ``FlatAdamW``: torch.optim.AdamW arithmetic executed as ONE libmemb pass over the model's flat fp32
parameter / gradient / moment buffers (global-norm clip folded in, bf16 weight shadow refreshed in the
same pass).  Other timm optimizers are outside the hot path.
"""
from __future__ import annotations

import ctypes
import json
import math

import torch

from . import _lib
from .vit_engine import CHUNK, engine_of


def get_num_layer_for_vit(var_name, num_max_layer):
    """Layer id of a parameter for layer-wise lr decay (mem/optim_factory.py:31-43): embedding 0, block i -> i + 1,
    everything else (shared rel-pos table, final norm, head) the last id."""
    if var_name in ("cls_token", "mask_token", "pos_embed") or var_name.startswith("patch_embed"):
        return 0
    if var_name.startswith("rel_pos_bias"):
        return num_max_layer - 1
    if var_name.startswith("blocks"):
        return int(var_name.split(".")[1]) + 1
    return num_max_layer - 1


class LayerDecayValueAssigner:
    """``values[layer_id]`` = lr scale of that layer (mem/optim_factory.py:46-53; built as
    ``layer_decay ** (num_layers + 1 - i)`` at run_class_finetuning.py:527-529)."""

    def __init__(self, values):
        self.values = values

    def get_scale(self, layer_id):
        return self.values[layer_id]

    def get_layer_id(self, var_name):
        return get_num_layer_for_vit(var_name, len(self.values))


def get_parameter_groups(model, weight_decay=1e-5, skip_list=(), get_num_layer=None, get_layer_scale=None):
    names, groups = {}, {}
    for name, param in model.named_parameters():
        if not param.requires_grad:
            continue
        no_decay = param.dim() == 1 or name.endswith(".bias") or name in skip_list
        gname = "no_decay" if no_decay else "decay"
        layer_id = None
        if get_num_layer is not None:
            layer_id = get_num_layer(name)
            gname = f"layer_{layer_id}_{gname}"
        if gname not in groups:
            scale = get_layer_scale(layer_id) if get_layer_scale is not None else 1.0
            groups[gname] = {"weight_decay": 0.0 if no_decay else weight_decay, "params": [], "lr_scale": scale}
            names[gname] = {"weight_decay": 0.0 if no_decay else weight_decay, "params": [], "lr_scale": scale}
        groups[gname]["params"].append(param)
        names[gname]["params"].append(name)
    print("Param groups = %s" % json.dumps(names, indent=2))
    return list(groups.values())


class FlatAdamW(torch.optim.Optimizer):
    """AdamW over a model's flat buffers (see ``vit_engine.FlatParams``).

    ``param_groups`` behave like torch's (the engine rewrites ``lr`` / ``weight_decay`` every step,
    engine_for_pretraining.py:124-130); ``step(max_norm)`` computes the global gradient norm, clips
    (``clip_grad_norm_`` semantics) and updates in two kernels, and returns the pre-clip norm as a device
    scalar.  ``grad_divisor`` divides the gradients first (sum-all-reduced gradients -> mean)."""

    def __init__(self, params, model, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2):
        super().__init__(params, dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay))
        assert len(self.param_groups) <= 64, "FlatAdamW supports at most 64 parameter groups"
        self.flat = engine_of(model).flat()
        flat = self.flat
        by_ptr = {flat.params[n].data_ptr(): n for n in flat.names}
        nchunks = flat.numel // CHUNK
        groups = torch.full((nchunks,), 255, dtype=torch.uint8)
        for gi, g in enumerate(self.param_groups):
            for p in g["params"]:
                n = by_ptr.get(p.data_ptr())
                assert n is not None, "FlatAdamW: parameter does not live in the model's flat buffer"
                c0 = flat.offsets[n] // CHUNK
                groups[c0: c0 + (flat.sizes[n] + CHUNK - 1) // CHUNK] = gi
        groups[nchunks - 1] = 254          # the status chunk (FlatParams.poison_grad): in the norm, never updated
        self.chunk_group = groups.to(flat.device)
        self.exp_avg = torch.zeros_like(flat.data)
        self.exp_avg_sq = torch.zeros_like(flat.data)
        self.sqnorm = torch.zeros(1, dtype=torch.float32, device=flat.device)
        self.norm = torch.zeros(1, dtype=torch.float32, device=flat.device)
        self.step_count = 0
        self.grad_divisor = 1.0

    @torch.no_grad()
    def step(self, closure=None, max_norm=0.0):
        assert closure is None
        lib, flat = _lib.load(), self.flat
        sp = _lib.stream_ptr(torch, flat.device)
        self.step_count += 1
        gs = 1.0 / self.grad_divisor
        self.sqnorm.zero_()
        # chunks of requires_grad = False tensors (group 255) stay out of the norm, like clip_grad_norm_(parameters)
        _lib.check(lib.memb_sqnorm_groups(flat.grad.data_ptr(), flat.numel, gs, self.chunk_group.data_ptr(),
                                          self.sqnorm.data_ptr(), sp))
        ng = len(self.param_groups)
        lr = (ctypes.c_float * ng)(*[float(g["lr"]) for g in self.param_groups])
        wd = (ctypes.c_float * ng)(*[float(g["weight_decay"]) for g in self.param_groups])
        b1, b2 = self.param_groups[0]["betas"]
        _lib.check(lib.memb_adamw(flat.data.data_ptr(), flat.grad.data_ptr(), self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr(),
                                  flat.shadow.data_ptr(), flat.numel, self.chunk_group.data_ptr(), lr, wd, ng, float(b1), float(b2),
                                  float(self.param_groups[0]["eps"]), self.step_count, gs, float(max_norm or 0.0),
                                  self.sqnorm.data_ptr(), sp))
        torch.sqrt(self.sqnorm, out=self.norm)
        return self.norm[0]

    def step_skipped(self):
        """The kernel leaves parameters and moments untouched when the gradient norm is not finite (the reference's
        GradScaler skips such steps, utils.py:357-371); the caller reports it once it has read the norm."""
        self.step_count = max(0, self.step_count - 1)

    def zero_grad(self, set_to_none: bool = True):
        # gradients live in the flat buffer: one fill kernel; p.grad stay bound views
        self.flat.zero_grad()

    # torch.optim.AdamW-shaped state for checkpoints (utils.save_model / auto_load_model)
    def state_dict(self):
        flat = self.flat
        ptr2name = {flat.params[n].data_ptr(): n for n in flat.names}
        state, groups, idx = {}, [], 0
        for g in self.param_groups:
            ids = []
            for p in g["params"]:
                n = ptr2name[p.data_ptr()]
                o, s = flat.offsets[n], flat.sizes[n]
                state[idx] = {"step": torch.tensor(float(self.step_count)),
                              "exp_avg": self.exp_avg[o:o + s].view_as(p).clone(),
                              "exp_avg_sq": self.exp_avg_sq[o:o + s].view_as(p).clone()}
                ids.append(idx)
                idx += 1
            groups.append({**{k: v for k, v in g.items() if k != "params"}, "params": ids})
        return {"state": state, "param_groups": groups}

    def load_state_dict(self, sd):
        flat = self.flat
        ptr2name = {flat.params[n].data_ptr(): n for n in flat.names}
        for g, saved in zip(self.param_groups, sd["param_groups"]):
            for k, v in saved.items():
                if k != "params":
                    g[k] = v
            for p, pid in zip(g["params"], saved["params"]):
                st = sd["state"].get(pid)
                if st is None:
                    continue
                n = ptr2name[p.data_ptr()]
                o, s = flat.offsets[n], flat.sizes[n]
                self.exp_avg[o:o + s].copy_(st["exp_avg"].reshape(-1))
                self.exp_avg_sq[o:o + s].copy_(st["exp_avg_sq"].reshape(-1))
                self.step_count = int(float(st["step"]))


def create_optimizer(args, model, get_num_layer=None, get_layer_scale=None, filter_bias_and_bn=True, skip_list=None):
    opt = args.opt.lower().split("_")[-1]
    if opt != "adamw":
        raise ValueError(f"mem_b200 implements the MEM pretraining optimizer (adamw); got opt={args.opt!r}")
    weight_decay = args.weight_decay
    if weight_decay and filter_bias_and_bn:
        skip = skip_list if skip_list is not None else (model.no_weight_decay() if hasattr(model, "no_weight_decay") else {})
        parameters = get_parameter_groups(model, weight_decay, skip, get_num_layer, get_layer_scale)
        weight_decay = 0.0
    else:
        parameters = [p for p in model.parameters() if p.requires_grad]
    kw = dict(lr=args.lr, weight_decay=weight_decay, betas=(0.9, 0.95))   # betas forced, optim_factory.py:121
    if getattr(args, "opt_eps", None) is not None:
        kw["eps"] = args.opt_eps
    core = model.module if hasattr(model, "module") else model
    opt_obj = FlatAdamW(parameters, core, **kw)
    for g in opt_obj.param_groups:
        g.setdefault("lr_scale", 1.0)
    return opt_obj
