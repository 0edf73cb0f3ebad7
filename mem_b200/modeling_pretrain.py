"""Masked-image-modelling ViT of the MEM pretraining step.

Drop-in for ``mem/modeling_pretrain.py`` (``VisionTransformerForMaskedImageModeling`` :22-126,
``pt_vit`` :128-140).  The reference registers only ``pt_vit``; ``beit_base_patch16_224_8k_vocab``
and ``beit_large_patch16_224_8k_vocab`` (the upstream BEiT names BASELINE.json uses) are
registered as presets of the same class.  ``state_dict`` keys equal the reference's.

``model(x, bool_masked_pos, return_all_tokens=False)`` returns the ``[sum(mask), vocab]`` logits in
row-major (batch, patch) order like boolean indexing does (modeling_pretrain.py:126); the work is
done by ``vit_engine`` on libmemb kernels and is differentiable through a custom autograd node.
``engine_for_pretraining.train_one_epoch`` uses the fused forward + cross-entropy + backward entry
instead (no logits round trip, no host sync for the masked-row count).
"""
from __future__ import annotations

from functools import partial

import torch
import torch.nn as nn

from .modeling_finetune import (Block, PatchEmbed, RelativePositionBias, _cfg_event, _rescale_residual_projections,
                                trunc_normal_ as _tn)
from .registry import register_model


def trunc_normal_(tensor, mean=0.0, std=1.0):
    # cut at +-1 std (modeling_pretrain.py:19-20)
    _tn(tensor, mean=mean, std=std, a=-std, b=std)


class VisionTransformerForMaskedImageModeling(nn.Module):
    def __init__(self, img_size=(224, 224), patch_size=(16, 16), in_chans=3, vocab_size=8192, embed_dim=768, depth=12,
                 num_heads=12, mlp_ratio=4.0, qkv_bias=True, qk_scale=None, drop_rate=0.0, attn_drop_rate=0.0,
                 drop_path_rate=0.0, norm_layer=None, init_values=None, attn_head_dim=None, use_abs_pos_emb=True,
                 use_rel_pos_bias=False, use_shared_rel_pos_bias=False, init_std=0.02, **kwargs):
        super().__init__()
        assert drop_rate == 0.0, "pos_drop / MLP dropout are off the MEM hot path (all configs use 0)"
        norm_layer = norm_layer or partial(nn.LayerNorm, eps=1e-6)
        self.num_features = self.embed_dim = embed_dim
        self.patch_embed = PatchEmbed(img_size, patch_size, in_chans, embed_dim)
        n = self.patch_embed.num_patches
        self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        self.mask_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        self.pos_embed = nn.Parameter(torch.zeros(1, n + 1, embed_dim)) if use_abs_pos_emb else None
        self.rel_pos_bias = RelativePositionBias(self.patch_embed.patch_shape, num_heads) if use_shared_rel_pos_bias else None
        rates = torch.linspace(0, drop_path_rate, depth).tolist()
        self.blocks = nn.ModuleList([
            Block(embed_dim, num_heads, mlp_ratio, qkv_bias, qk_scale, drop_rate, attn_drop_rate, rates[i],
                  init_values=init_values, norm_layer=norm_layer,
                  window_size=self.patch_embed.patch_shape if use_rel_pos_bias else None, attn_head_dim=attn_head_dim)
            for i in range(depth)])
        self.norm = norm_layer(embed_dim)
        self.init_std = init_std
        self.lm_head = nn.Linear(embed_dim, vocab_size)

        for p in (self.pos_embed, self.cls_token, self.mask_token):
            if p is not None:
                trunc_normal_(p, std=init_std)
        for m in self.modules():
            if isinstance(m, (nn.Linear, nn.Conv2d)):
                trunc_normal_(m.weight, std=init_std)
                if m.bias is not None:
                    nn.init.zeros_(m.bias)
            elif isinstance(m, nn.LayerNorm):
                nn.init.ones_(m.weight)
                nn.init.zeros_(m.bias)
        _rescale_residual_projections(self.blocks)

    def no_weight_decay(self):
        return {"pos_embed", "cls_token"}

    def get_num_layers(self):
        return len(self.blocks)

    def forward(self, x, bool_masked_pos, return_all_tokens=False):
        from .vit_engine import masked_forward
        return masked_forward(self, x, bool_masked_pos, return_all_tokens)


@register_model
def pt_vit(pretrained=False, **kwargs):
    init_ckpt = kwargs.pop("init_ckpt", None)
    model = VisionTransformerForMaskedImageModeling(qkv_bias=True, norm_layer=partial(nn.LayerNorm, eps=1e-6), **kwargs)
    model.default_cfg = _cfg_event()
    if pretrained:
        model.load_state_dict(torch.load(init_ckpt, map_location="cpu")["model"])
    return model


def _beit_preset(embed_dim, depth, num_heads, kwargs):
    cfg = dict(img_size=(224, 224), patch_size=(16, 16), embed_dim=embed_dim, depth=depth, num_heads=num_heads,
               mlp_ratio=4, vocab_size=8192)
    cfg.update(kwargs)
    return pt_vit(**cfg)


@register_model
def beit_base_patch16_224_8k_vocab(pretrained=False, **kwargs):
    return _beit_preset(768, 12, 12, dict(kwargs, pretrained=pretrained))


@register_model
def beit_large_patch16_224_8k_vocab(pretrained=False, **kwargs):
    return _beit_preset(1024, 24, 16, dict(kwargs, pretrained=pretrained))
