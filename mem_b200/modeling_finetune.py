"""ViT building blocks with the reference's parameter names, run by ``vit_engine.VitEngine``.

Drop-in for ``mem/modeling_finetune.py`` (Block / Attention / Mlp / PatchEmbed /
RelativePositionBias / DropPath :42-247, VisionTransformer + ``ft_vit`` :250-385).  The sub-modules
here are *parameter containers*: they own tensors whose ``state_dict`` keys equal the reference's
(checkpoints interchange), while the arithmetic of a forward/backward pass is the kernel sequence in
``vit_engine`` (tcgen05 GEMMs, fused attention, LayerNorm ... from ``libmemb.so``).  There is no
PyTorch-operator path: calling a sub-module on its own raises.
"""
from __future__ import annotations

import math
from functools import partial

import torch
import torch.nn as nn

from .registry import register_model


def _pair(v):
    return tuple(v) if isinstance(v, (tuple, list)) else (v, v)


def trunc_normal_(tensor, mean=0.0, std=1.0, a=-2.0, b=2.0):
    """timm.models.layers.trunc_normal_ (absolute cut-offs a, b) -- same as torch's initialiser."""
    return nn.init.trunc_normal_(tensor, mean=mean, std=std, a=a, b=b)


def _cfg_event(url="", **kwargs):
    cfg = dict(url=url, num_classes=2, input_size=(3, 128, 128), pool_size=None, crop_pct=1,
               interpolation="bicubic", mean=(0.5, 0, 0.5), std=(0.5, 0, 0.5))
    cfg.update(kwargs)
    return cfg


def _cfg(url="", **kwargs):
    cfg = dict(url=url, num_classes=1000, input_size=(3, 224, 224), pool_size=None, crop_pct=0.9,
               interpolation="bicubic", mean=(0.5, 0.5, 0.5), std=(0.5, 0.5, 0.5))
    cfg.update(kwargs)
    return cfg


class _Container(nn.Module):
    def forward(self, *a, **k):
        raise RuntimeError(
            f"{type(self).__name__} is a parameter container: run the whole model (its forward goes through "
            "mem_b200.vit_engine on the CUDA library); there is no per-module PyTorch path.")


class DropPath(_Container):
    """Per-sample stochastic depth; the engine draws the keep mask (timm ``drop_path`` rule)."""

    def __init__(self, drop_prob=None):
        super().__init__()
        self.drop_prob = drop_prob

    def extra_repr(self):
        return f"p={self.drop_prob}"


class _NoDrop(_Container):
    drop_prob = 0.0


class Mlp(_Container):
    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU, drop=0.0):
        super().__init__()
        assert act_layer is nn.GELU, "the fc1 epilogue implements exact-erf GELU (nn.GELU) only"
        assert drop == 0.0, "dropout inside the MLP is not on the MEM hot path (all configs use 0)"
        self.fc1 = nn.Linear(in_features, hidden_features or in_features)
        self.fc2 = nn.Linear(hidden_features or in_features, out_features or in_features)


def relative_position_index(window_size):
    """[Wh*Ww+1, Wh*Ww+1] int64 index into the (2Wh-1)(2Ww-1)+3 row bias table; row/column 0 is the
    cls token and uses the three extra rows (cls->tok, tok->cls, cls->cls)."""
    wh, ww = window_size
    n_rel = (2 * wh - 1) * (2 * ww - 1) + 3
    ys, xs = torch.meshgrid(torch.arange(wh), torch.arange(ww), indexing="ij")
    ys, xs = ys.reshape(-1), xs.reshape(-1)
    dy = ys[:, None] - ys[None, :] + (wh - 1)
    dx = xs[:, None] - xs[None, :] + (ww - 1)
    idx = torch.empty(wh * ww + 1, wh * ww + 1, dtype=torch.int64)
    idx[1:, 1:] = dy * (2 * ww - 1) + dx
    idx[0, :] = n_rel - 3
    idx[:, 0] = n_rel - 2
    idx[0, 0] = n_rel - 1
    return idx, n_rel


class Attention(_Container):
    def __init__(self, dim, num_heads=8, qkv_bias=False, qk_scale=None, attn_drop=0.0, proj_drop=0.0,
                 window_size=None, attn_head_dim=None):
        super().__init__()
        assert attn_drop == 0.0 and proj_drop == 0.0, "attention dropout is not on the MEM hot path"
        self.num_heads = num_heads
        head_dim = attn_head_dim if attn_head_dim is not None else dim // num_heads
        inner = head_dim * num_heads
        self.scale = qk_scale or head_dim ** -0.5
        self.qkv = nn.Linear(dim, 3 * inner, bias=False)
        if qkv_bias:  # the key bias is structurally zero (modeling_finetune.py:131-133)
            self.q_bias = nn.Parameter(torch.zeros(inner))
            self.v_bias = nn.Parameter(torch.zeros(inner))
        else:
            self.q_bias = self.v_bias = None
        if window_size:
            self.window_size = window_size
            idx, self.num_relative_distance = relative_position_index(window_size)
            self.relative_position_bias_table = nn.Parameter(torch.zeros(self.num_relative_distance, num_heads))
            self.register_buffer("relative_position_index", idx)
        else:
            self.window_size = None
            self.relative_position_bias_table = None
            self.relative_position_index = None
        self.proj = nn.Linear(inner, dim)


class Block(_Container):
    def __init__(self, dim, num_heads, mlp_ratio=4.0, qkv_bias=False, qk_scale=None, drop=0.0, attn_drop=0.0,
                 drop_path=0.0, init_values=None, act_layer=nn.GELU, norm_layer=nn.LayerNorm, window_size=None,
                 attn_head_dim=None):
        super().__init__()
        self.norm1 = norm_layer(dim)
        self.attn = Attention(dim, num_heads=num_heads, qkv_bias=qkv_bias, qk_scale=qk_scale, attn_drop=attn_drop,
                              proj_drop=drop, window_size=window_size, attn_head_dim=attn_head_dim)
        self.drop_path = DropPath(drop_path) if drop_path > 0.0 else _NoDrop()
        self.norm2 = norm_layer(dim)
        self.mlp = Mlp(in_features=dim, hidden_features=int(dim * mlp_ratio), act_layer=act_layer, drop=drop)
        if init_values is not None and init_values > 0:
            self.gamma_1 = nn.Parameter(init_values * torch.ones(dim))
            self.gamma_2 = nn.Parameter(init_values * torch.ones(dim))
        else:
            self.gamma_1 = self.gamma_2 = None


class PatchEmbed(_Container):
    def __init__(self, img_size=(224, 224), patch_size=(16, 16), in_chans=3, embed_dim=768):
        super().__init__()
        img_size, patch_size = _pair(img_size), _pair(patch_size)
        self.patch_shape = (img_size[0] // patch_size[0], img_size[1] // patch_size[1])
        self.num_patches = self.patch_shape[0] * self.patch_shape[1]
        self.img_size, self.patch_size = img_size, patch_size
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size)


class RelativePositionBias(_Container):
    def __init__(self, window_size, num_heads):
        super().__init__()
        self.window_size = window_size
        idx, self.num_relative_distance = relative_position_index(window_size)
        self.relative_position_bias_table = nn.Parameter(torch.zeros(self.num_relative_distance, num_heads))
        self.register_buffer("relative_position_index", idx)


def _rescale_residual_projections(blocks):
    # proj / fc2 weights divided by sqrt(2 * layer_id), layer_id from 1 (fix_init_weight)
    for i, blk in enumerate(blocks):
        s = math.sqrt(2.0 * (i + 1))
        blk.attn.proj.weight.data.div_(s)
        blk.mlp.fc2.weight.data.div_(s)


class VisionTransformer(nn.Module):
    """Classification ViT (``ft_vit``): mean-pooled patch tokens -> fc_norm -> head."""

    def __init__(self, img_size=(224, 224), patch_size=(16, 16), in_chans=3, num_classes=1000, embed_dim=768, depth=12,
                 num_heads=12, mlp_ratio=4.0, qkv_bias=False, qk_scale=None, drop_rate=0.0, attn_drop_rate=0.0,
                 drop_path_rate=0.0, norm_layer=nn.LayerNorm, init_values=None, use_abs_pos_emb=True,
                 use_rel_pos_bias=False, use_shared_rel_pos_bias=False, use_mean_pooling=True, init_scale=0.001,
                 use_batch_norm=False):
        super().__init__()
        assert drop_rate == 0.0 and not use_batch_norm, "dropout / linear-probe batch norm are off the MEM hot path"
        self.num_classes = num_classes
        self.num_features = self.embed_dim = embed_dim
        self.patch_embed = PatchEmbed(img_size, patch_size, in_chans, embed_dim)
        n = self.patch_embed.num_patches
        self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        self.pos_embed = nn.Parameter(torch.zeros(1, n + 1, embed_dim)) if use_abs_pos_emb else None
        self.rel_pos_bias = RelativePositionBias(self.patch_embed.patch_shape, num_heads) if use_shared_rel_pos_bias else None
        self.use_rel_pos_bias = use_rel_pos_bias
        rates = torch.linspace(0, drop_path_rate, depth).tolist()
        self.blocks = nn.ModuleList([
            Block(embed_dim, num_heads, mlp_ratio, qkv_bias, qk_scale, drop_rate, attn_drop_rate, rates[i],
                  init_values=init_values, norm_layer=norm_layer,
                  window_size=self.patch_embed.patch_shape if use_rel_pos_bias else None)
            for i in range(depth)])
        self.norm = nn.Identity() if use_mean_pooling else norm_layer(embed_dim)
        self.fc_norm = norm_layer(embed_dim) if use_mean_pooling else None
        self.head = nn.Linear(embed_dim, num_classes) if num_classes > 0 else nn.Identity()

        if self.pos_embed is not None:
            trunc_normal_(self.pos_embed, std=0.02)
        trunc_normal_(self.cls_token, std=0.02)
        for m in self.modules():
            if isinstance(m, nn.Linear):
                trunc_normal_(m.weight, std=0.02)
                if m.bias is not None:
                    nn.init.zeros_(m.bias)
            elif isinstance(m, nn.LayerNorm):
                nn.init.ones_(m.weight)
                nn.init.zeros_(m.bias)
        _rescale_residual_projections(self.blocks)
        if isinstance(self.head, nn.Linear):
            self.head.weight.data.mul_(init_scale)
            self.head.bias.data.mul_(init_scale)

    def get_num_layers(self):
        return len(self.blocks)

    def no_weight_decay(self):
        return {"pos_embed", "cls_token"}

    def get_classifier(self):
        return self.head

    def forward(self, x):
        from .vit_engine import classify_forward
        return classify_forward(self, x)


@register_model
def ft_vit(pretrained=False, **kwargs):
    model = VisionTransformer(qkv_bias=True, norm_layer=partial(nn.LayerNorm, eps=1e-6), **kwargs)
    model.default_cfg = _cfg_event()
    return model
