"""Execution engine of the masked ViT: flat parameter storage + forward / backward over libmemb.

The ``nn.Module`` classes in ``modeling_finetune`` / ``modeling_pretrain`` are parameter containers
with the reference's names (so ``state_dict`` keys match ``mem/modeling_pretrain.py`` /
``mem/modeling_finetune.py``); this module is what runs them:

* parameters live in ONE flat fp32 buffer (every tensor on a 1024-element boundary, ordered by when
  backward finishes with them), gradients in a parallel flat buffer, weights have a bf16 shadow --
  one fused AdamW pass, one (bucketed) all-reduce stream, no per-tensor loops;
* forward / backward are explicit kernel sequences (tcgen05 GEMMs with fused epilogues, fused
  attention, LayerNorm, cross entropy ...); the residual stream stays fp32 like the reference under
  autocast, GEMM operands are bf16, LayerNorm / softmax / CE math is fp32.

Reference call stack restated here: ``forward_features`` (modeling_pretrain.py:97-117), ``Block.forward``
(modeling_finetune.py:182-189), ``Attention.forward`` (:128-157), ``Mlp.forward`` (:66-71),
``forward`` + ``CrossEntropyLoss`` (modeling_pretrain.py:119-126, engine_for_pretraining.py:152).
"""
from __future__ import annotations

import math

import torch

from . import _lib, ops
from ._lib import (DT_BF16, DT_F32, EPI_ATOMIC_ADD, EPI_BIAS_GELU, EPI_DGELU, EPI_RESIDUAL, EPI_STORE, EPI_STORE_ROWDOT)

CHUNK = 1024  # flat-buffer alignment (elements) == AdamW chunk size


def _sp(torch_mod=torch, device=None):
    return _lib.stream_ptr(torch_mod, device)


class FlatParams:
    """One fp32 buffer for all parameters of a model (+ grads + bf16 shadow)."""

    def __init__(self, model, order):
        params = dict(model.named_parameters())
        names = [n for n in order if n in params] + [n for n in params if n not in set(order)]
        dev = next(iter(params.values())).device
        self.names, self.offsets, self.sizes = names, {}, {}
        off = 0
        for n in names:
            self.offsets[n] = off
            self.sizes[n] = params[n].numel()
            off += (params[n].numel() + CHUNK - 1) // CHUNK * CHUNK
        off += CHUNK       # one status chunk at the end: no tensor lives there (see poison_grad)
        self.numel = off
        self.device = dev
        self.data = torch.zeros(off, dtype=torch.float32, device=dev)
        self.grad = torch.zeros(off, dtype=torch.float32, device=dev)
        self.shadow = torch.zeros(off, dtype=torch.bfloat16, device=dev)
        self.params = params
        with torch.no_grad():
            for n in names:
                p = params[n]
                view = self.data[self.offsets[n]: self.offsets[n] + p.numel()].view_as(p)
                view.copy_(p.data)
                p.data = view
        self.shadow_version = -1
        self._plist = [params[n] for n in names]
        self.grad_views = {n: self.grad[self.offsets[n]: self.offsets[n] + self.sizes[n]].view_as(params[n])
                           for n in names}

    def valid(self):
        p0 = self.params[self.names[0]]
        return p0.data_ptr() == self.data.data_ptr() + 4 * self.offsets[self.names[0]] and \
            self.params[self.names[-1]].data_ptr() == self.data.data_ptr() + 4 * self.offsets[self.names[-1]]

    def refresh_shadow(self, force=False):
        # p.data are views of the flat buffer but carry their own version counters (load_state_dict, manual
        # in-place edits); libmemb's AdamW writes through raw pointers and refreshes the shadow itself.
        v = sum(p._version for p in self._plist)
        if force or v != self.shadow_version:
            lib = _lib.load()
            _lib.check(lib.memb_cast_bf16(self.data.data_ptr(), self.shadow.data_ptr(), self.numel, _sp(device=self.device)))
            self.shadow_version = v

    def w16(self, name):
        p = self.params[name]
        return self.shadow[self.offsets[name]: self.offsets[name] + p.numel()].view_as(p)

    def g(self, name):
        return self.grad_views[name]

    def zero_grad(self):
        lib = _lib.load()
        _lib.check(lib.memb_fill_f32(self.grad.data_ptr(), self.numel, 0.0, _sp(device=self.device)))

    def poison_grad(self, value):
        """Store a device scalar (0 or NaN) in the status chunk at the end of the gradient buffer: no kernel writes
        there, the DP all-reduce and the global norm read it (see engine_for_pretraining.train_one_epoch)."""
        self.grad[self.numel - 1: self.numel].copy_(value.reshape(1))


def get_flat(model, order_fn):
    flat = getattr(model, "_memb_flat", None)
    if flat is None or not flat.valid() or flat.device != next(model.parameters()).device:
        flat = FlatParams(model, order_fn())
        model._memb_flat = flat
    return flat


class _Bufs:
    """Named scratch / activation tensors, allocated once per shape."""

    def __init__(self):
        self.t = {}

    def get(self, name, shape, dtype, device, zero=False):
        key = name
        cur = self.t.get(key)
        shape = tuple(int(s) for s in shape)
        if cur is None or tuple(cur.shape) != shape or cur.dtype != dtype or cur.device != device:
            cur = (torch.zeros if zero else torch.empty)(shape, dtype=dtype, device=device)
            self.t[key] = cur
        return cur


def _ln_fwd(lib, x, w, b, eps, y, mean, rstd, rows, D, row_index=None, count=None):
    _lib.check(lib.memb_layernorm_fwd(x.data_ptr(), x.stride(0), w.data_ptr(), b.data_ptr(), eps, rows, D,
                                      y.data_ptr(), y.stride(0), mean.data_ptr(), rstd.data_ptr(),
                                      ops._ptr(row_index), ops._ptr(count), _sp(device=x.device)))


def _ln_bwd(lib, dy, x, w, mean, rstd, rows, D, dx, dw, db, row_index=None, count=None):
    _lib.check(lib.memb_layernorm_bwd(dy.data_ptr(), DT_BF16 if dy.dtype == torch.bfloat16 else DT_F32, dy.stride(0),
                                      x.data_ptr(), x.stride(0), w.data_ptr(), mean.data_ptr(), rstd.data_ptr(), rows, D,
                                      dx.data_ptr(), dx.stride(0), ops._ptr(dw), ops._ptr(db), ops._ptr(row_index),
                                      ops._ptr(count), _sp(device=x.device)))


def _index_t(mod):
    """Transposed relative_position_index (cached on the module): attention_bwd emits dS^T[b, h, key, query], so
    the scatter into the table walks index[q, k] at position [k, q]."""
    idx = mod.relative_position_index
    cached = getattr(mod, "_index_t_cache", None)
    if cached is None or cached.device != idx.device or cached.shape != idx.shape:
        cached = idx.t().contiguous()
        mod._index_t_cache = cached
    return cached


class VitEngine:
    """Runs forward/backward of a (masked) ViT whose parameters follow the reference naming."""

    def __init__(self, model):
        self.model = model
        self.bufs = _Bufs()
        self.saved = None
        self.generation = 0

    # Activations saved for backward live in engine-owned buffers that the next forward of the same model reuses
    # (with or without grad: the no-grad path ping-pongs through the same residual buffers): ONE outstanding
    # autograd graph per model.  The stamp turns a violation (forward, another forward, then backward through the
    # first) into an error instead of silently wrong gradients.
    def stamp(self, need_grad):
        self.generation += 1
        return self.generation

    def check_stamp(self, generation):
        if generation != self.generation:
            raise RuntimeError("mem_b200: backward through a graph whose saved activations were overwritten by a later "
                               "forward of the same model (one outstanding graph per model: call backward before the "
                               "next forward of this model)")

    # ---- static description of the model ---------------------------------------------------
    def cfg(self):
        m = self.model
        pe = m.patch_embed
        D = m.embed_dim
        depth = len(m.blocks)
        H = m.blocks[0].attn.num_heads
        hidden = m.blocks[0].mlp.fc1.out_features
        gh, gw = pe.patch_shape
        return dict(D=D, depth=depth, H=H, hidden=hidden, P=gh * gw, N=gh * gw + 1, patch=pe.patch_size[0],
                    C=pe.proj.in_channels, img=pe.img_size, eps=m.blocks[0].norm1.eps)

    def param_order(self):
        """Backward-completion order: head first, then blocks last-to-first, then the embedding."""
        m = self.model
        order = []
        for n in ("lm_head.weight", "lm_head.bias", "head.weight", "head.bias", "fc_norm.weight", "fc_norm.bias",
                  "norm.weight", "norm.bias"):
            order.append(n)
        per_block = ["gamma_2", "mlp.fc2.weight", "mlp.fc2.bias", "mlp.fc1.weight", "mlp.fc1.bias", "norm2.weight",
                     "norm2.bias", "gamma_1", "attn.proj.weight", "attn.proj.bias", "attn.qkv.weight", "attn.q_bias",
                     "attn.v_bias", "attn.relative_position_bias_table", "norm1.weight", "norm1.bias"]
        for i in reversed(range(len(m.blocks))):
            order += [f"blocks.{i}.{n}" for n in per_block]
        order += ["patch_embed.proj.weight", "patch_embed.proj.bias", "cls_token", "mask_token", "pos_embed",
                  "rel_pos_bias.relative_position_bias_table"]
        return order

    def flat(self):
        return get_flat(self.model, self.param_order)

    # ---- forward ----------------------------------------------------------------------------
    def forward_features(self, x, mask_u8, need_grad, droppath_scales=None):
        """x fp32 [B,C,H,W] -> residual stream fp32 [B*N, D] after the last block (pre final norm).

        mask_u8: uint8 [B*P] or None.  droppath_scales: list over blocks of (s1, s2) fp32 [B] tensors or
        None (per-sample DropPath keep/(1-p) factors, timm drop_path semantics)."""
        lib = _lib.load()
        c = self.cfg()
        m, flat = self.model, self.flat()
        flat.refresh_shadow()
        dev = x.device
        B = x.shape[0]
        D, N, P, H, depth, hidden = c["D"], c["N"], c["P"], c["H"], c["depth"], c["hidden"]
        M = B * N
        assert x.shape[1] == c["C"] and tuple(x.shape[2:]) == tuple(c["img"]), \
            f"Input image size ({x.shape[2]}*{x.shape[3]}) doesn't match model ({c['img'][0]}*{c['img'][1]})."
        x = x.contiguous().float()
        bf, f32 = torch.bfloat16, torch.float32
        g = self.bufs.get
        sp = _sp(device=dev)

        # patch embedding: patchify -> GEMM (+bias, mask-token rows, write at row offset 1) -> cls / pos
        kdim = c["C"] * c["patch"] ** 2
        a0 = g("a0", (B * P, kdim), bf, dev)
        _lib.check(lib.memb_patchify(x.data_ptr(), B, c["C"], c["img"][0], c["img"][1], c["patch"], a0.data_ptr(), sp))
        xs = [g(f"x{i}", (M, D), f32, dev) for i in range(depth + 1)] if need_grad else \
            [g("x_a", (M, D), f32, dev), g("x_b", (M, D), f32, dev)]
        x0 = xs[0]
        w_pe = flat.w16("patch_embed.proj.weight").view(D, kdim)
        mask_tok = m.mask_token.view(-1) if (mask_u8 is not None and getattr(m, "mask_token", None) is not None) else None
        ops.gemm(a0, w_pe, out=x0, bias=m.patch_embed.proj.bias, row_remap=(P, N, 1, M),
                 rowmask=mask_u8 if mask_tok is not None else None, maskvec=mask_tok)
        pos = m.pos_embed.view(N, D) if getattr(m, "pos_embed", None) is not None else None
        _lib.check(lib.memb_cls_pos(x0.data_ptr(), m.cls_token.data_ptr(), ops._ptr(pos), B, N, D, sp))

        # relative position bias (shared table: once per step; per-block tables: once per block)
        ldk = (N + 7) // 8 * 8
        shared_bias = None
        if getattr(m, "rel_pos_bias", None) is not None:
            rp = m.rel_pos_bias
            shared_bias = self._packed_bias(lib, rp.relative_position_bias_table, rp.relative_position_index, "", N, H, ldk, dev, sp)

        # q/v biases -> one [depth, 3D] vector table (k part stays zero), modeling_finetune.py:131-133
        qkvb = g("qkvb", (depth, 3 * D), f32, dev, zero=True)
        has_qkv_bias = m.blocks[0].attn.q_bias is not None
        if has_qkv_bias:  # plumbing only: two strided copies per step
            with torch.no_grad():
                torch.stack([blk.attn.q_bias for blk in m.blocks], out=qkvb[:, :D])
                torch.stack([blk.attn.v_bias for blk in m.blocks], out=qkvb[:, 2 * D:])

        saved = []
        scale = m.blocks[0].attn.scale
        for i, blk in enumerate(m.blocks):
            tag = f"L{i}_" if need_grad else "L_"
            xin = xs[i] if need_grad else xs[i % 2]
            xout = xs[i + 1] if need_grad else xs[(i + 1) % 2]
            pre = f"blocks.{i}."
            s1, s2 = droppath_scales[i] if droppath_scales is not None else (None, None)
            ln1 = g(tag + "ln1", (M, D), bf, dev); mu1 = g(tag + "mu1", (M,), f32, dev); rs1 = g(tag + "rs1", (M,), f32, dev)
            _ln_fwd(lib, xin, blk.norm1.weight, blk.norm1.bias, blk.norm1.eps, ln1, mu1, rs1, M, D)
            qkv = g(tag + "qkv", (M, 3 * D), bf, dev)
            ops.gemm(ln1, flat.w16(pre + "attn.qkv.weight"), out=qkv, bias=qkvb[i] if has_qkv_bias else None)
            bias_pair = shared_bias
            if blk.attn.relative_position_bias_table is not None:
                own = self._packed_bias(lib, blk.attn.relative_position_bias_table, blk.attn.relative_position_index,
                                        tag, N, H, ldk, dev, sp)
                assert shared_bias is None, "shared and per-block relative position bias together are not supported"
                bias_pair = own
            ao = g(tag + "ao", (M, D), bf, dev); lse = g(tag + "lse", (B, H, N), f32, dev)
            _lib.check(lib.memb_attention_fwd(qkv.data_ptr(), ops._ptr(bias_pair[0]) if bias_pair else None, ldk, B, N, H,
                                              D // H, scale, ao.data_ptr(), lse.data_ptr(), sp))
            xmid = g(tag + "xmid", (M, D), f32, dev)
            br1 = g(tag + "br1", (M, D), bf, dev) if need_grad else None
            ops.gemm(ao, flat.w16(pre + "attn.proj.weight"), out=xmid, epilogue=EPI_RESIDUAL, bias=blk.attn.proj.bias,
                     aux=xin, d2=br1, colscale=blk.gamma_1, rowscale=s1, rows_per_group=N)
            ln2 = g(tag + "ln2", (M, D), bf, dev); mu2 = g(tag + "mu2", (M,), f32, dev); rs2 = g(tag + "rs2", (M,), f32, dev)
            _ln_fwd(lib, xmid, blk.norm2.weight, blk.norm2.bias, blk.norm2.eps, ln2, mu2, rs2, M, D)
            act = g(tag + "act", (M, hidden), bf, dev)
            fpre = g(tag + "fpre", (M, hidden), bf, dev) if need_grad else None
            ops.gemm(ln2, flat.w16(pre + "mlp.fc1.weight"), out=act, epilogue=EPI_BIAS_GELU, bias=blk.mlp.fc1.bias, d2=fpre)
            br2 = g(tag + "br2", (M, D), bf, dev) if need_grad else None
            ops.gemm(act, flat.w16(pre + "mlp.fc2.weight"), out=xout, epilogue=EPI_RESIDUAL, bias=blk.mlp.fc2.bias,
                     aux=xmid, d2=br2, colscale=blk.gamma_2, rowscale=s2, rows_per_group=N)
            if need_grad:
                saved.append(dict(xin=xin, ln1=ln1, mu1=mu1, rs1=rs1, qkv=qkv, ao=ao, lse=lse, xmid=xmid, br1=br1,
                                  ln2=ln2, mu2=mu2, rs2=rs2, act=act, fpre=fpre, br2=br2, s1=s1, s2=s2, bias=bias_pair))
        xlast = xs[depth] if need_grad else xs[depth % 2]
        ctx = dict(B=B, M=M, a0=a0, mask=mask_u8, blocks=saved, ldk=ldk, shared_bias=shared_bias, cfg=c, xlast=xlast)
        return xlast, ctx

    def _packed_bias(self, lib, table, index, tag, N, H, ldk, dev, sp):
        """RelativePositionBias.forward (modeling_finetune.py:242-247) -> (bias, bias^T) in the attention kernels'
        packed layout (include/memb.h: memb_attention_pack_bias)."""
        g, f32 = self.bufs.get, torch.float32
        dense, denseT = g("relb_dense", (H, N, ldk), f32, dev), g("relbT_dense", (H, N, ldk), f32, dev)
        _lib.check(lib.memb_relpos_gather(table.data_ptr(), index.data_ptr(), N, H, ldk, dense.data_ptr(), denseT.data_ptr(), sp))
        packed = (g(tag + "relb", (H, _lib.ATTN_BIAS_FLOATS_PER_HEAD), f32, dev),
                  g(tag + "relbT", (H, _lib.ATTN_BIAS_FLOATS_PER_HEAD), f32, dev))
        _lib.check(lib.memb_attention_pack_bias(dense.data_ptr(), ldk, N, H, packed[0].data_ptr(), sp))
        _lib.check(lib.memb_attention_pack_bias(denseT.data_ptr(), ldk, N, H, packed[1].data_ptr(), sp))
        return packed

    # ---- masked-token head + cross entropy (pretraining) ----------------------------------------
    def pretrain_head(self, xlast, ctx, mask_u8, tokens, need_grad, cap=None, want_logits=False):
        """final LN on the masked rows -> lm_head -> (optional) CE.  Returns dict with logits [cap,V],
        count (device int32), stats (device fp32 [4]: loss_sum, hits, count, -)."""
        lib = _lib.load()
        m, flat, c = self.model, self.flat(), ctx["cfg"]
        dev = xlast.device
        B, D, N, P = ctx["B"], c["D"], c["N"], c["P"]
        V = m.lm_head.out_features
        cap = min(int(cap), B * P) if cap else B * P
        cap = max(cap, 1)
        # buffers are sized for the largest cap seen (rounded to 1024 rows) and sliced, so a varying masked
        # count does not reallocate the big logits / dlogits tensors every step
        self._cap_alloc = max(getattr(self, "_cap_alloc", 0), (cap + 1023) // 1024 * 1024)
        ca = self._cap_alloc
        g = self.bufs.get
        sp = _sp(device=dev)
        row_index = g("row_index", (ca,), torch.int32, dev)[:cap]; patch_index = g("patch_index", (ca,), torch.int32, dev)[:cap]
        count = g("count", (1,), torch.int32, dev)
        _lib.check(lib.memb_mask_compact(mask_u8.data_ptr(), B, P, row_index.data_ptr(), patch_index.data_ptr(),
                                         count.data_ptr(), cap, sp))
        xm = g("xm", (ca, D), torch.bfloat16, dev)[:cap]; mu = g("mu_f", (ca,), torch.float32, dev)[:cap]; rs = g("rs_f", (ca,), torch.float32, dev)[:cap]
        _ln_fwd(lib, xlast, m.norm.weight, m.norm.bias, m.norm.eps, xm, mu, rs, cap, D, row_index, count)
        logits = g("logits", (ca, V), torch.float32, dev)[:cap]
        ops.gemm(xm, flat.w16("lm_head.weight"), out=logits, bias=m.lm_head.bias)
        out = dict(logits=logits, count=count, row_index=row_index, patch_index=patch_index, xm=xm, mu=mu, rs=rs, cap=cap)
        if tokens is not None:
            stats = g("stats", (4,), torch.float32, dev)
            stats.zero_()
            dlogits = g("dlogits", (ca, V), torch.bfloat16, dev)[:cap] if need_grad else None
            _lib.check(lib.memb_cross_entropy(logits.data_ptr(), logits.stride(0), tokens.data_ptr(), patch_index.data_ptr(),
                                              count.data_ptr(), cap, V, ops._ptr(dlogits), V, stats.data_ptr(), 1.0, sp))
            out.update(stats=stats, dlogits=dlogits)
        return out

    # ---- classification head (ft_vit): mean pool -> fc_norm -> head, or norm -> cls token -> head ------
    def classify_head(self, xlast, ctx):
        lib = _lib.load()
        m, c = self.model, ctx["cfg"]
        dev = xlast.device
        B, D, N = ctx["B"], c["D"], c["N"]
        C = m.head.out_features
        g, sp = self.bufs.get, _sp(device=dev)
        if getattr(m, "fc_norm", None) is None:
            # use_mean_pooling=False (modeling_finetune.py:286-287, :349-352): final LayerNorm, then the cls token.
            # Only row b*N of every sample reaches the head, so the LayerNorm runs on those B rows (row gather).
            rows = g("cls_rows", (B,), torch.int32, dev)
            if getattr(self, "_cls_rows_for", None) != (B, N):
                rows.copy_(torch.arange(B, dtype=torch.int32, device=dev) * N)
                self._cls_rows_for = (B, N)
            cnt = g("cls_count", (1,), torch.int32, dev)
            cnt.fill_(B)
            z = g("z_cls", (B, D), torch.bfloat16, dev); mu = g("mu_cls", (B,), torch.float32, dev); rs = g("rs_cls", (B,), torch.float32, dev)
            _ln_fwd(lib, xlast, m.norm.weight, m.norm.bias, m.norm.eps, z, mu, rs, B, D, rows, cnt)
            logits = g("cls_logits", (B, C), torch.float32, dev)
            _lib.check(lib.memb_linear_small_fwd(z.data_ptr(), m.head.weight.data_ptr(), ops._ptr(m.head.bias), B, D, C,
                                                 logits.data_ptr(), sp))
            return dict(logits=logits, z=z, mu=mu, rs=rs, rows=rows, count=cnt, cls_token=True)
        pooled = g("pooled", (B, D), torch.float32, dev)
        _lib.check(lib.memb_meanpool_fwd(xlast.data_ptr(), B, N, D, pooled.data_ptr(), sp))
        z = g("z_cls", (B, D), torch.bfloat16, dev); mu = g("mu_cls", (B,), torch.float32, dev); rs = g("rs_cls", (B,), torch.float32, dev)
        _ln_fwd(lib, pooled, m.fc_norm.weight, m.fc_norm.bias, m.fc_norm.eps, z, mu, rs, B, D)
        logits = g("cls_logits", (B, C), torch.float32, dev)
        _lib.check(lib.memb_linear_small_fwd(z.data_ptr(), m.head.weight.data_ptr(), ops._ptr(m.head.bias), B, D, C,
                                             logits.data_ptr(), sp))
        return dict(logits=logits, pooled=pooled, z=z, mu=mu, rs=rs)

    def backward_classify(self, ctx, head, dlogits, bucket_hook=None):
        """dlogits fp32 [B, num_classes] -> every parameter gradient (accumulated into flat.grad)."""
        lib = _lib.load()
        m, flat, c = self.model, self.flat(), ctx["cfg"]
        B, M, D, N = ctx["B"], ctx["M"], c["D"], c["N"]
        C = m.head.out_features
        dev = flat.device
        g, sp = self.bufs.get, _sp(device=dev)
        dz = g("dz_cls", (B, D), torch.float32, dev)
        has_bias = m.head.bias is not None
        _lib.check(lib.memb_linear_small_bwd(dlogits.data_ptr(), head["z"].data_ptr(), m.head.weight.data_ptr(), B, D, C,
                                             flat.g("head.weight").data_ptr(), flat.g("head.bias").data_ptr() if has_bias else None,
                                             dz.data_ptr(), sp))
        gres = g("gres", (M, D), torch.float32, dev)
        if head.get("cls_token"):
            # gradient reaches the residual stream only through the B cls rows of the final LayerNorm
            _lib.check(lib.memb_fill_f32(gres.data_ptr(), gres.numel(), 0.0, sp))
            _ln_bwd(lib, dz, ctx["xlast"], m.norm.weight, head["mu"], head["rs"], B, D, gres, flat.g("norm.weight"),
                    flat.g("norm.bias"), head["rows"], head["count"])
        else:
            dpool = g("dpool", (B, D), torch.float32, dev)
            _lib.check(lib.memb_fill_f32(dpool.data_ptr(), dpool.numel(), 0.0, sp))
            _ln_bwd(lib, dz, head["pooled"], m.fc_norm.weight, head["mu"], head["rs"], B, D, dpool, flat.g("fc_norm.weight"),
                    flat.g("fc_norm.bias"))
            _lib.check(lib.memb_meanpool_bwd(dpool.data_ptr(), B, N, D, gres.data_ptr(), sp))
        if bucket_hook:
            bucket_hook("head")
        self._backward_blocks(ctx, gres, bucket_hook)
        self._backward_embed(ctx, gres)
        if bucket_hook:
            bucket_hook("embed")

    # ---- backward -----------------------------------------------------------------------------
    def backward_pretrain(self, ctx, head, grad_scale_dev, dlogits=None, bucket_hook=None):
        """Accumulates all parameter gradients into flat.grad.  dlogits (bf16 [cap,V]) defaults to the one
        the fused CE produced; grad_scale_dev is an optional device scalar multiplied into it."""
        lib = _lib.load()
        m, flat, c = self.model, self.flat(), ctx["cfg"]
        B, M, D, N, P, H, hidden = ctx["B"], ctx["M"], c["D"], c["N"], c["P"], c["H"], c["hidden"]
        dev = flat.device
        g = self.bufs.get
        sp = _sp(device=dev)
        V = m.lm_head.out_features
        cap = head["cap"]
        dl = dlogits if dlogits is not None else head["dlogits"]
        # lm_head: wgrad (both operands MN-major), bias grad, dgrad (B = W as stored)
        ops.gemm(dl, head["xm"], out=flat.g("lm_head.weight"), a_layout=1, b_layout=1, epilogue=EPI_ATOMIC_ADD,
                 alpha_dev=grad_scale_dev)
        dxm = g("dxm", (self._cap_alloc, D), torch.bfloat16, dev)[:cap]
        ops.gemm(dl, flat.w16("lm_head.weight"), out=dxm, b_layout=1, alpha_dev=grad_scale_dev)
        if grad_scale_dev is None:
            _lib.check(lib.memb_colsum_bf16(dl.data_ptr(), V, cap, V, flat.g("lm_head.bias").data_ptr(), sp))
        else:  # rare path (scaled loss): bias grad through a scaled temporary
            tmp = torch.zeros(V, dtype=torch.float32, device=dev)
            _lib.check(lib.memb_colsum_bf16(dl.data_ptr(), V, cap, V, tmp.data_ptr(), sp))
            flat.g("lm_head.bias").add_(tmp * grad_scale_dev)
        gres = g("gres", (M, D), torch.float32, dev)
        _lib.check(lib.memb_fill_f32(gres.data_ptr(), gres.numel(), 0.0, sp))
        xlast = ctx["xlast"]
        _ln_bwd(lib, dxm, xlast, m.norm.weight, head["mu"], head["rs"], cap, D, gres, flat.g("norm.weight"),
                flat.g("norm.bias"), head["row_index"], head["count"])
        if bucket_hook:
            bucket_hook("head")
        self._backward_blocks(ctx, gres, bucket_hook)
        self._backward_embed(ctx, gres)
        if bucket_hook:
            bucket_hook("embed")

    def _backward_blocks(self, ctx, gres, bucket_hook=None):
        lib = _lib.load()
        m, flat, c = self.model, self.flat(), ctx["cfg"]
        B, M, D, N, H, hidden, ldk = ctx["B"], ctx["M"], c["D"], c["N"], c["H"], c["hidden"], ctx["ldk"]
        dev = flat.device
        g = self.bufs.get
        sp = _sp(device=dev)
        bf = torch.bfloat16
        dz = g("dz", (M, D), bf, dev); dh = g("dh", (M, hidden), bf, dev); dln = g("dln", (M, D), bf, dev)
        dao = g("dao", (M, D), bf, dev); dqkv = g("dqkv", (M, 3 * D), bf, dev)
        ds = g("ds", (B, H, N, ldk), bf, dev)
        attn_ws = g("attn_ws", (int(lib.memb_attention_bwd_workspace_bytes(B, N, H)),), torch.uint8, dev)
        shared = ctx["shared_bias"] is not None
        dbias_acc = g("dbias_acc", (H, N, ldk), torch.float32, dev)
        if shared:
            _lib.check(lib.memb_fill_f32(dbias_acc.data_ptr(), dbias_acc.numel(), 0.0, sp))
        scale = m.blocks[0].attn.scale
        # The branch backward (LayerScale x DropPath, bias gradient) of a sub-block runs inside the LayerNorm backward that
        # precedes it in the backward pass (memb_layernorm_bwd_branch) whenever four column accumulators fit the block's
        # shared memory; only the first one of the pass (last block, MLP branch) is a launch of its own.
        fuse = 4 * 8 * D * 4 <= 113 * 1024

        # attention branch: the proj bias gradient of THIS pass goes to a scratch row first; memb_vbias_chain adds it to the
        # parameter's gradient and turns it into the v_bias gradient (colsum(dV) = colsum(dZ) W_proj: no pass over dV)
        pw = m.blocks[0].attn.proj.weight
        chain = (m.blocks[0].attn.q_bias is not None and m.blocks[0].attn.proj.bias is not None and tuple(pw.shape) == (D, D)
                 and pw.dtype == torch.float32)
        pb_tmp = g("proj_bias_pass", (len(m.blocks), D), torch.float32, dev) if chain else None
        if chain:
            _lib.check(lib.memb_fill_f32(pb_tmp.data_ptr(), pb_tmp.numel(), 0.0, sp))

        def branch_args(j, which):
            """(branch, colscale, rowscale, dcolscale, dbias) of block j's MLP (2) / attention (1) branch."""
            b, sj, pj = m.blocks[j], ctx["blocks"][j], f"blocks.{j}."
            gamma = b.gamma_2 if which == 2 else b.gamma_1
            bias = flat.g(pj + "mlp.fc2.bias") if which == 2 else (pb_tmp[j] if chain else flat.g(pj + "attn.proj.bias"))
            return (sj[f"br{which}"], gamma, sj[f"s{which}"], flat.g(pj + f"gamma_{which}") if gamma is not None else None, bias)

        def branch_bwd(j, which):
            br, cs, rs, dcs, db = branch_args(j, which)
            _lib.check(lib.memb_branch_bwd(gres.data_ptr(), D, ops._ptr(br), D, ops._ptr(cs), ops._ptr(rs), N, M, D, dz.data_ptr(), D,
                                           ops._ptr(dcs), db.data_ptr(), sp))

        def ln_bwd_then_branch(dy, x, norm, mean, rstd, dw, db_, j, which):
            if not fuse or j < 0:
                _ln_bwd(lib, dy, x, norm.weight, mean, rstd, M, D, gres, dw, db_)
                if j >= 0:
                    branch_bwd(j, which)
                return
            br, cs, rs, dcs, db = branch_args(j, which)
            _lib.check(lib.memb_layernorm_bwd_branch(dy.data_ptr(), DT_BF16 if dy.dtype == torch.bfloat16 else DT_F32, dy.stride(0),
                                                     x.data_ptr(), x.stride(0), norm.weight.data_ptr(), mean.data_ptr(), rstd.data_ptr(),
                                                     M, D, gres.data_ptr(), gres.stride(0), ops._ptr(dw), ops._ptr(db_), ops._ptr(br), D,
                                                     ops._ptr(cs), ops._ptr(rs), N, dz.data_ptr(), D, ops._ptr(dcs), db.data_ptr(), sp))

        branch_bwd(len(m.blocks) - 1, 2)
        for i in reversed(range(len(m.blocks))):
            blk, s, pre = m.blocks[i], ctx["blocks"][i], f"blocks.{i}."
            G = lambda n: flat.g(pre + n)  # noqa: E731
            # ---- MLP branch (dz = its branch backward, produced by the previous LayerNorm backward)
            ops.gemm(dz, s["act"], out=G("mlp.fc2.weight"), a_layout=1, b_layout=1, epilogue=EPI_ATOMIC_ADD)
            # fc1's bias gradient (column sums of dh) leaves the same epilogue
            ops.gemm(dz, flat.w16(pre + "mlp.fc2.weight"), out=dh, b_layout=1, epilogue=EPI_DGELU, aux=s["fpre"],
                     colsum=G("mlp.fc1.bias"))
            ops.gemm(dh, s["ln2"], out=G("mlp.fc1.weight"), a_layout=1, b_layout=1, epilogue=EPI_ATOMIC_ADD)
            ops.gemm(dh, flat.w16(pre + "mlp.fc1.weight"), out=dln, b_layout=1)
            # ---- attention branch (its branch backward rides on the LayerNorm backward of norm2)
            ln_bwd_then_branch(dln, s["xmid"], blk.norm2, s["mu2"], s["rs2"], G("norm2.weight"), G("norm2.bias"), i, 1)
            ops.gemm(dz, s["ao"], out=G("attn.proj.weight"), a_layout=1, b_layout=1, epilogue=EPI_ATOMIC_ADD)
            # dO = dz W_proj; the same epilogue leaves rowsum(dO * O) per head in the attention backward's workspace
            ops.gemm(dz, flat.w16(pre + "attn.proj.weight"), out=dao, b_layout=1, epilogue=EPI_STORE_ROWDOT, aux=s["ao"],
                     rowdot=attn_ws, rows_per_group=N)
            bias_pair = s["bias"]
            _lib.check(lib.memb_attention_bwd(s["qkv"].data_ptr(), None, dao.data_ptr(), s["lse"].data_ptr(),
                                              ops._ptr(bias_pair[0]) if bias_pair else None,
                                              ops._ptr(bias_pair[1]) if bias_pair else None, ldk, B, N, H, D // H, scale,
                                              dqkv.data_ptr(), ds.data_ptr() if bias_pair else None,
                                              attn_ws.data_ptr(), attn_ws.numel(), sp))
            if bias_pair is not None:
                if not shared:
                    _lib.check(lib.memb_fill_f32(dbias_acc.data_ptr(), dbias_acc.numel(), 0.0, sp))
                _lib.check(lib.memb_batch_reduce_bf16(ds.data_ptr(), B, H * N * ldk, dbias_acc.data_ptr(), sp))
                if not shared:
                    _lib.check(lib.memb_relpos_scatter(dbias_acc.data_ptr(), ldk, _index_t(blk.attn).data_ptr(),
                                                       N, H, G("attn.relative_position_bias_table").data_ptr(), sp))
            if blk.attn.q_bias is not None:
                _lib.check(lib.memb_colsum_bf16(dqkv.data_ptr(), 3 * D, M, D, G("attn.q_bias").data_ptr(), sp))
                if chain:
                    _lib.check(lib.memb_vbias_chain(pb_tmp[i].data_ptr(), blk.attn.proj.weight.data_ptr(), D,
                                                    G("attn.proj.bias").data_ptr(), G("attn.v_bias").data_ptr(), sp))
                else:
                    _lib.check(lib.memb_colsum_bf16(dqkv.data_ptr() + 2 * D * 2, 3 * D, M, D, G("attn.v_bias").data_ptr(), sp))
            ops.gemm(dqkv, s["ln1"], out=G("attn.qkv.weight"), a_layout=1, b_layout=1, epilogue=EPI_ATOMIC_ADD)
            ops.gemm(dqkv, flat.w16(pre + "attn.qkv.weight"), out=dln, b_layout=1)
            # block i - 1's MLP branch backward rides on the LayerNorm backward of norm1 (its gradients belong to the next bucket)
            ln_bwd_then_branch(dln, s["xin"], blk.norm1, s["mu1"], s["rs1"], G("norm1.weight"), G("norm1.bias"), i - 1, 2)
            if bucket_hook:
                bucket_hook(i)
        if shared:
            rp = m.rel_pos_bias
            _lib.check(lib.memb_relpos_scatter(dbias_acc.data_ptr(), ldk, _index_t(rp).data_ptr(), N, H,
                                               flat.g("rel_pos_bias.relative_position_bias_table").data_ptr(), sp))

    def _backward_embed(self, ctx, gres):
        lib = _lib.load()
        m, flat, c = self.model, self.flat(), ctx["cfg"]
        B, D, N, P = ctx["B"], c["D"], c["N"], c["P"]
        dev = flat.device
        sp = _sp(device=dev)
        dpatch = self.bufs.get("dpatch", (B * P, D), torch.bfloat16, dev)
        mask = ctx["mask"]
        if mask is None:
            mask = self.bufs.get("zero_mask", (B * P,), torch.uint8, dev, zero=True)
        has_mt = getattr(m, "mask_token", None) is not None
        has_pos = getattr(m, "pos_embed", None) is not None
        _lib.check(lib.memb_embed_bwd(gres.data_ptr(), mask.data_ptr(), B, P, D, dpatch.data_ptr(),
                                      flat.g("mask_token").data_ptr() if has_mt else None, flat.g("cls_token").data_ptr(),
                                      flat.g("patch_embed.proj.bias").data_ptr(),
                                      flat.g("pos_embed").data_ptr() if has_pos else None, sp))
        kdim = ctx["a0"].shape[1]
        ops.gemm(dpatch, ctx["a0"], out=flat.g("patch_embed.proj.weight").view(D, kdim), a_layout=1, b_layout=1,
                 epilogue=EPI_ATOMIC_ADD)


def droppath_scales(model, B, device, training):
    """Per-block, per-sample keep/(1-p) factors with timm 0.4.12 ``drop_path`` semantics
    (``floor(keep + U[0,1)) / keep``, call site mem/modeling_finetune.py:49-50), drawn from torch's
    generator in the order the reference would draw them (attention branch, then MLP branch)."""
    if not training:
        return None
    probs = [float(getattr(blk.drop_path, "drop_prob", 0.0) or 0.0) for blk in model.blocks]
    idx = [i for i, p in enumerate(probs) if p > 0.0]
    if not idx:
        return None
    # one batched draw for all blocks (3 launches instead of 8 per block); the per-block keep probabilities live on the
    # device once per (model, device)
    cache = getattr(model, "_memb_dp_keep", None)
    if cache is None or cache[0] != (tuple(probs), str(device)):
        keep = torch.tensor([1.0 - probs[i] for i in idx], dtype=torch.float32).view(-1, 1, 1).to(device)
        cache = ((tuple(probs), str(device)), keep)
        object.__setattr__(model, "_memb_dp_keep", cache)
    keep = cache[1]
    scales = (torch.rand(len(idx), 2, B, device=device) + keep).floor_().div_(keep)
    out = [(None, None)] * len(probs)
    for k, i in enumerate(idx):
        out[i] = (scales[k, 0], scales[k, 1])
    return out


# ------------------------------------------------------------------------------------------------
# Model-level entry points (what modeling_pretrain / engine_for_pretraining call)
# ------------------------------------------------------------------------------------------------
def engine_of(model) -> VitEngine:
    eng = getattr(model, "_memb_engine", None)
    if eng is None or eng.model is not model:
        eng = VitEngine(model)
        object.__setattr__(model, "_memb_engine", eng)  # not a sub-module / not in the state_dict
    return eng


def _mask_u8(bool_masked_pos, B, P):
    m = bool_masked_pos.reshape(B, -1)
    assert m.shape[1] == P, f"bool_masked_pos has {m.shape[1]} entries per sample, the model has {P} patches"
    return m.to(torch.uint8).contiguous().view(-1)


def bind_param_grads(flat: FlatParams, params):
    """Make every ``p.grad`` a view into the flat gradient buffer (zeroing segments whose grad was None),
    so torch optimizers / clip_grad_norm_ see the gradients the kernels accumulate."""
    todo = [n for n in flat.names if params[n].requires_grad and params[n].grad is not flat.grad_views[n]]
    if len(todo) == len(flat.names):
        flat.zero_grad()
    for n in todo:
        p = params[n]
        if len(todo) != len(flat.names):
            if p.grad is not None:
                flat.grad_views[n].copy_(p.grad)
            else:
                flat.grad_views[n].zero_()
        p.grad = flat.grad_views[n]


class _MaskedVitFn(torch.autograd.Function):
    """logits[sum(mask), V] = lm_head(norm(blocks(embed(x, mask)))[:, 1:][mask]) with a kernel backward
    that accumulates straight into the flat gradient buffer (``p.grad`` are views of it)."""

    @staticmethod
    def forward(ctx, model, x, mask_u8, head_mask_u8, droppath, *params):
        eng = engine_of(model)
        need_grad = bool(ctx.needs_input_grad and any(ctx.needs_input_grad[5:]))
        xlast, fctx = eng.forward_features(x, mask_u8, need_grad, droppath)
        head = eng.pretrain_head(xlast, fctx, head_mask_u8, None, need_grad)
        n = int(head["count"].item())  # the output shape is data dependent: one host sync on this path
        ctx.eng, ctx.fctx, ctx.head, ctx.n = eng, fctx, head, n
        ctx.generation = eng.stamp(need_grad)
        return head["logits"][:n].clone()

    @staticmethod
    def backward(ctx, grad_logits):
        eng, head = ctx.eng, ctx.head
        eng.check_stamp(ctx.generation)
        flat = eng.flat()
        bind_param_grads(flat, flat.params)
        V = grad_logits.shape[1]
        dl = eng.bufs.get("dlogits", (eng._cap_alloc, V), torch.bfloat16, grad_logits.device)[:head["cap"]]
        dl.zero_()
        dl[:ctx.n].copy_(grad_logits)
        eng.backward_pretrain(ctx.fctx, head, None, dlogits=dl)
        return (None,) * (5 + len(flat.params))


def masked_forward(model, x, bool_masked_pos, return_all_tokens=False):
    """``VisionTransformerForMaskedImageModeling.forward`` (mem/modeling_pretrain.py:119-126)."""
    _lib.require_cuda()
    if not x.is_cuda:
        raise RuntimeError("mem_b200 models run on CUDA tensors only (no CPU path)")
    eng = engine_of(model)
    c = eng.cfg()
    B, P = x.shape[0], c["P"]
    mask = _mask_u8(bool_masked_pos, B, P)
    head_mask = torch.ones_like(mask) if return_all_tokens else mask
    dp = droppath_scales(model, B, x.device, model.training)
    flat = eng.flat()
    params = [flat.params[n] for n in flat.names]
    out = _MaskedVitFn.apply(model, x, mask, head_mask, dp, *params)
    if return_all_tokens:
        out = out.view(B, P, -1)
    return out


def pretrain_step(model, samples, bool_masked_pos, tokens, grad_scale_dev=None, bucket_hook=None, backward=True, cap=None):
    """Fused MEM step body: forward + masked cross entropy (+ accuracy) + backward into the flat gradient
    buffer.  Replaces engine_for_pretraining.py:147-161 (autocast forward, CrossEntropyLoss, scaled backward)
    and :233 (mlm_acc).  ``tokens``: int64 [B, P] codebook indices (get_codebook_indices output); labels are
    tokens[mask].  ``cap``: upper bound of the number of masked patches in the batch (sizes the lm_head GEMM; masked
    patches beyond it would be dropped) -- default B*P; train_one_epoch passes the exact count taken from the host
    copy of the mask.  Returns the device stats tensor [sum of per-token losses, top-1 hits, masked count, 0]."""
    eng = engine_of(model)
    c = eng.cfg()
    B = samples.shape[0]
    mask = _mask_u8(bool_masked_pos, B, c["P"])
    tokens = tokens.reshape(-1).contiguous()
    assert tokens.dtype == torch.int64 and tokens.numel() == B * c["P"]
    flat = eng.flat()
    if backward:
        bind_param_grads(flat, flat.params)
    dp = droppath_scales(model, B, samples.device, model.training)
    xlast, fctx = eng.forward_features(samples, mask, backward, dp)
    head = eng.pretrain_head(xlast, fctx, mask, tokens, backward, cap=cap)
    if backward:
        eng.backward_pretrain(fctx, head, grad_scale_dev, bucket_hook=bucket_hook)
    return head["stats"]


class _ClassifyVitFn(torch.autograd.Function):
    """logits[B, num_classes] = head(fc_norm(mean over patch tokens of blocks(embed(x)))) with a kernel backward
    that accumulates into the flat gradient buffer (``VisionTransformer.forward``, modeling_finetune.py:343-357)."""

    @staticmethod
    def forward(ctx, model, x, droppath, *params):
        eng = engine_of(model)
        need_grad = bool(ctx.needs_input_grad and any(ctx.needs_input_grad[3:]))
        xlast, fctx = eng.forward_features(x, None, need_grad, droppath)
        head = eng.classify_head(xlast, fctx)
        ctx.eng, ctx.fctx, ctx.head = eng, fctx, head
        ctx.generation = eng.stamp(need_grad)
        return head["logits"].clone()

    @staticmethod
    def backward(ctx, grad_logits):
        eng = ctx.eng
        eng.check_stamp(ctx.generation)
        flat = eng.flat()
        bind_param_grads(flat, flat.params)
        eng.backward_classify(ctx.fctx, ctx.head, grad_logits.contiguous().float())
        return (None,) * (3 + len(flat.params))


def classify_forward(model, x):
    """``VisionTransformer.forward`` of ft_vit (mem/modeling_finetune.py:354-357): mean-pooling head
    (``use_mean_pooling=True``, the configs' choice) or final norm + cls token (``False``)."""
    _lib.require_cuda()
    if not x.is_cuda:
        raise RuntimeError("mem_b200 models run on CUDA tensors only (no CPU path)")
    eng = engine_of(model)
    dp = droppath_scales(model, x.shape[0], x.device, model.training)
    flat = eng.flat()
    params = [flat.params[n] for n in flat.names]
    return _ClassifyVitFn.apply(model, x, dp, *params)
