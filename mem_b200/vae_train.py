"""dVAE training step and decoder on libmemb.

Kernel schedule behind ``DiscreteVAE.forward(img, return_loss=..., return_recons=...)`` and ``DiscreteVAE.decode``
(``eventvae/vae/vae_model.py:160-213``), the calls ``eventvae/train_vae.py:304-392`` makes::

    loss, recons = vae(images, return_loss=True, return_recons=True, temp=temp)
    opt.zero_grad(); loss.backward(); clip_grad_norm_(vae.parameters(), clip); opt.step()

Activations are NHWC bf16 matrices ``[B*H*W, C]`` (channels padded to a multiple of 8), every convolution is an explicit
im2col / col2im around the tcgen05 GEMM (``ops.gemm``: bf16 operands, fp32 accumulation), weights are repacked from the
``nn.Parameter`` tensors (reference layouts, so ``state_dict`` is unchanged) at every call, and ``loss.backward()`` runs
the kernel backward through a ``torch.autograd.Function`` whose inputs are the module's parameters -- any torch
optimizer the reference's loop uses (``Adam`` + ``ExponentialLR``, train_vae.py:219-236) works on the result.

Training is not index-exact work: unlike the tokenizer path (``vae_model._Tokenizer``, fp32-faithful) this path computes
in bf16 like the ViT; tests bound it by the reference's own bf16-autocast-vs-fp32 error (tests/test_dvae_train_gpu.py).
"""
from __future__ import annotations

import torch

from . import _lib, ops
from ._lib import EPI_ATOMIC_ADD

_IM2COL_BUDGET = 1 << 30      # bytes of im2col / col2im scratch per convolution call (batch is chunked to fit)


def _pad8(n):
    return (n + 7) // 8 * 8


def _sp(t):
    return _lib.stream_ptr(torch, t.device)


class _Act:
    """NHWC bf16 activation: ``t`` is ``[B*H*W, C]`` (C = padded channel count)."""
    __slots__ = ("t", "B", "H", "W", "C")

    def __init__(self, t, B, H, W, C):
        self.t, self.B, self.H, self.W, self.C = t, B, H, W, C


def _ew(op, a, b=None, out=None):
    lib = _lib.load()
    out = torch.empty_like(a) if out is None else out
    _lib.check(lib.memb_vae_ew_bf16(op, a.data_ptr(), ops._ptr(b), out.data_ptr(), a.numel(), _sp(a)))
    return out


def _colsum(dy, cols, out):
    """out[cols] += column sums of the bf16 matrix dy."""
    lib = _lib.load()
    _lib.check(lib.memb_colsum_bf16(dy.data_ptr(), dy.stride(0), dy.shape[0], cols, out.data_ptr(), _sp(dy)))


class _Conv:
    """``nn.Conv2d`` (k x k, stride s, padding p) [+ ReLU] as im2col + GEMM."""

    def __init__(self, mod, relu):
        self.mod, self.relu = mod, relu
        self.k, self.s, self.p = mod.kernel_size[0], mod.stride[0], mod.padding[0]
        self.cin, self.cout = mod.in_channels, mod.out_channels
        self.cinp, self.coutp = _pad8(self.cin), _pad8(self.cout)

    def pack(self, dev):
        w = self.mod.weight.detach()
        wm = torch.zeros(self.coutp, self.k, self.k, self.cinp, dtype=torch.bfloat16, device=dev)
        wm[:self.cout, :, :, :self.cin] = w.permute(0, 2, 3, 1)
        self.wm = wm.view(self.coutp, self.k * self.k * self.cinp)
        self.bias = torch.zeros(self.coutp, dtype=torch.float32, device=dev)
        self.bias[:self.cout] = self.mod.bias.detach()

    def _chunks(self, a, OH, OW):
        per_img = OH * OW * self.k * self.k * self.cinp * 4          # fp32 dcol is the larger of the two scratch matrices
        step = max(1, min(a.B, _IM2COL_BUDGET // max(per_img, 1)))
        return [(b0, min(a.B, b0 + step)) for b0 in range(0, a.B, step)]

    def _col(self, a, b0, b1, OH, OW):
        if self.k == 1 and self.s == 1 and self.p == 0:
            return a.t[b0 * a.H * a.W: b1 * a.H * a.W]
        lib = _lib.load()
        col = torch.empty((b1 - b0) * OH * OW, self.k * self.k * a.C, dtype=torch.bfloat16, device=a.t.device)
        x = a.t[b0 * a.H * a.W: b1 * a.H * a.W]
        _lib.check(lib.memb_vae_im2col(x.data_ptr(), b1 - b0, a.H, a.W, a.C, self.k, self.k, self.s, self.p, col.data_ptr(), _sp(x)))
        return col

    def out_hw(self, a):
        return (a.H + 2 * self.p - self.k) // self.s + 1, (a.W + 2 * self.p - self.k) // self.s + 1

    def fwd(self, a, save, out_f32=False):
        assert a.C == self.cinp
        OH, OW = self.out_hw(a)
        y = torch.empty(a.B * OH * OW, self.coutp, dtype=torch.float32 if out_f32 else torch.bfloat16, device=a.t.device)
        for b0, b1 in self._chunks(a, OH, OW):
            ops.gemm(self._col(a, b0, b1, OH, OW), self.wm, out=y[b0 * OH * OW: b1 * OH * OW], bias=self.bias)
        if self.relu:
            _ew(0, y, out=y)
        if save is not None:
            save.append((self, a, y))
        return y if out_f32 else _Act(y, a.B, OH, OW, self.coutp)

    def bwd(self, dy, saved, grads, need_dx=True):
        """dy bf16 [M, coutp] -> dx bf16 [B*H*W, cinp]; accumulates the packed weight / bias gradients into ``grads``."""
        _, a, y = saved
        lib = _lib.load()
        dev = dy.device
        if self.relu:
            dy = _ew(1, dy, y)
        gw, gb = grads.setdefault(self, (torch.zeros(self.wm.shape, dtype=torch.float32, device=dev),
                                         torch.zeros(self.coutp, dtype=torch.float32, device=dev)))
        _colsum(dy, self.coutp, gb)
        OH, OW = self.out_hw(a)
        dx = torch.empty_like(a.t) if need_dx else None
        for b0, b1 in self._chunks(a, OH, OW):
            d = dy[b0 * OH * OW: b1 * OH * OW]
            col = self._col(a, b0, b1, OH, OW)
            ops.gemm(d, col, out=gw, a_layout=1, b_layout=1, epilogue=EPI_ATOMIC_ADD)       # dW += dy^T col
            if not need_dx:
                continue
            xs = dx[b0 * a.H * a.W: b1 * a.H * a.W]
            if self.k == 1 and self.s == 1 and self.p == 0:
                ops.gemm(d, self.wm, out=xs, b_layout=1)                                      # dx = dy W
            else:
                dcol = ops.gemm(d, self.wm, b_layout=1, out_dtype=torch.float32)            # [M, k*k*cinp]
                _lib.check(lib.memb_vae_col2im(dcol.data_ptr(), dcol.stride(0), b1 - b0, a.H, a.W, a.C, self.k, self.k, self.s,
                                               self.p, None, 0, None, xs.data_ptr(), None, a.C, _sp(xs)))
        return dx

    def param_grads(self, grads):
        gw, gb = grads[self]
        w = gw.view(self.coutp, self.k, self.k, self.cinp)[:self.cout, :, :, :self.cin].permute(0, 3, 1, 2)
        return [(self.mod.weight, w), (self.mod.bias, gb[:self.cout])]


class _ConvT:
    """``nn.ConvTranspose2d`` (k x k, stride s, padding p) + ReLU as GEMM + col2im."""

    def __init__(self, mod, relu):
        self.mod, self.relu = mod, relu
        self.k, self.s, self.p = mod.kernel_size[0], mod.stride[0], mod.padding[0]
        self.cin, self.cout = mod.in_channels, mod.out_channels
        self.cinp, self.coutp = _pad8(self.cin), _pad8(self.cout)

    def pack(self, dev):
        w = self.mod.weight.detach()                                  # [cin, cout, k, k]
        wm = torch.zeros(self.cinp, self.k, self.k, self.coutp, dtype=torch.bfloat16, device=dev)
        wm[:self.cin, :, :, :self.cout] = w.permute(0, 2, 3, 1)
        self.wm = wm.view(self.cinp, self.k * self.k * self.coutp)
        self.bias = torch.zeros(self.coutp, dtype=torch.float32, device=dev)
        self.bias[:self.cout] = self.mod.bias.detach()

    def out_hw(self, a):
        return (a.H - 1) * self.s - 2 * self.p + self.k, (a.W - 1) * self.s - 2 * self.p + self.k

    def _chunks(self, a):
        per_img = a.H * a.W * self.k * self.k * self.coutp * 4
        step = max(1, min(a.B, _IM2COL_BUDGET // max(per_img, 1)))
        return [(b0, min(a.B, b0 + step)) for b0 in range(0, a.B, step)]

    def fwd(self, a, save):
        assert a.C == self.cinp
        lib = _lib.load()
        OH, OW = self.out_hw(a)
        y = torch.empty(a.B * OH * OW, self.coutp, dtype=torch.bfloat16, device=a.t.device)
        for b0, b1 in self._chunks(a):
            x = a.t[b0 * a.H * a.W: b1 * a.H * a.W]
            col = ops.gemm(x, self.wm, b_layout=1, out_dtype=torch.float32)                   # [M_in, k*k*coutp]
            ys = y[b0 * OH * OW: b1 * OH * OW]
            _lib.check(lib.memb_vae_col2im(col.data_ptr(), col.stride(0), b1 - b0, OH, OW, self.coutp, self.k, self.k, self.s, self.p,
                                           self.bias.data_ptr(), int(self.relu), None, ys.data_ptr(), None, self.coutp, _sp(ys)))
        if save is not None:
            save.append((self, a, y))
        return _Act(y, a.B, OH, OW, self.coutp)

    def bwd(self, dy, saved, grads, need_dx=True):
        _, a, y = saved
        lib = _lib.load()
        dev = dy.device
        if self.relu:
            dy = _ew(1, dy, y)
        gw, gb = grads.setdefault(self, (torch.zeros(self.wm.shape, dtype=torch.float32, device=dev),
                                         torch.zeros(self.coutp, dtype=torch.float32, device=dev)))
        _colsum(dy, self.coutp, gb)
        OH, OW = self.out_hw(a)
        dx = torch.empty_like(a.t) if need_dx else None
        for b0, b1 in self._chunks(a):
            d = dy[b0 * OH * OW: b1 * OH * OW]
            dcol = torch.empty((b1 - b0) * a.H * a.W, self.k * self.k * self.coutp, dtype=torch.bfloat16, device=dev)
            _lib.check(lib.memb_vae_im2col(d.data_ptr(), b1 - b0, OH, OW, self.coutp, self.k, self.k, self.s, self.p, dcol.data_ptr(), _sp(d)))
            x = a.t[b0 * a.H * a.W: b1 * a.H * a.W]
            ops.gemm(x, dcol, out=gw, a_layout=1, b_layout=1, epilogue=EPI_ATOMIC_ADD)       # dWt += x^T im2col(dy)
            if need_dx:
                ops.gemm(dcol, self.wm, out=dx[b0 * a.H * a.W: b1 * a.H * a.W])             # dx = im2col(dy) Wt^T
        return dx

    def param_grads(self, grads):
        gw, gb = grads[self]
        w = gw.view(self.cinp, self.k, self.k, self.coutp)[:self.cin, :, :, :self.cout].permute(0, 3, 1, 2)
        return [(self.mod.weight, w), (self.mod.bias, gb[:self.cout])]


class _Res:
    """``ResBlock``: x + conv1x1(relu(conv3x3(relu(conv3x3(x)))))  (vae_model.py:29-41)."""

    def __init__(self, mod):
        self.convs = [_Conv(mod.net[0], True), _Conv(mod.net[2], True), _Conv(mod.net[4], False)]

    def pack(self, dev):
        for c in self.convs:
            c.pack(dev)

    def fwd(self, a, save):
        local = [] if save is not None else None
        h = a
        for c in self.convs:
            h = c.fwd(h, local)
        y = _ew(2, h.t, a.t)
        if save is not None:
            save.append((self, local, None))
        return _Act(y, a.B, a.H, a.W, a.C)

    def bwd(self, dy, saved, grads, need_dx=True):
        _, local, _ = saved
        d = dy
        for c, sv in zip(reversed(self.convs), reversed(local)):
            d = c.bwd(d, sv, grads)
        return _ew(2, d, dy)

    def param_grads(self, grads):
        return [pg for c in self.convs for pg in c.param_grads(grads)]


def _build(seq):
    """nn.Sequential of the reference's encoder / decoder -> kernel-schedule layers."""
    from .vae_model import ResBlock
    layers = []
    for m in seq:
        if isinstance(m, ResBlock):
            layers.append(_Res(m))
        elif isinstance(m, torch.nn.Sequential):             # Conv2d / ConvTranspose2d + ReLU
            layers.append((_ConvT if isinstance(m[0], torch.nn.ConvTranspose2d) else _Conv)(m[0], True))
        elif isinstance(m, torch.nn.Conv2d):
            layers.append(_Conv(m, False))
        else:
            raise TypeError(f"unexpected dVAE layer {type(m).__name__}")
    return layers


class VaeTrainer:
    """Forward / backward schedule of one ``DiscreteVAE`` (built lazily by the module, one per model)."""

    def __init__(self, vae):
        self.vae = vae
        self.enc, self.dec = _build(vae.encoder), _build(vae.decoder)
        self.kind = {"mse": 0, "smooth_l1": 1}.get(vae.loss_name)

    def params(self):
        v = self.vae
        return [v.codebook.weight] + list(v.encoder.parameters()) + list(v.decoder.parameters())

    def _pack(self, dev):
        for l in self.enc + self.dec:
            l.pack(dev)
        v = self.vae
        D = v.codebook.weight.shape[1]
        self.dp = _pad8(D)
        cb = torch.zeros(v.num_tokens, self.dp, dtype=torch.bfloat16, device=dev)
        cb[:, :D] = v.codebook.weight.detach()
        self.cb = cb

    # ------------------------------------------------------------------ decoder only
    def decode(self, img_seq):
        lib = _lib.load()
        v = self.vae
        dev = img_seq.device
        self._pack(dev)
        B, n = img_seq.shape
        h, w = v.input_H >> v.num_layers, v.input_W >> v.num_layers
        assert n == h * w, f"decode expects {h * w} tokens per image, got {n}"
        idx = img_seq.reshape(-1).contiguous().long()
        z = torch.empty(B * n, self.dp, dtype=torch.bfloat16, device=dev)
        err = ops._err_flag(torch, dev)
        err.zero_()
        _lib.check(lib.memb_vae_gather_rows(v.codebook.weight.detach().float().contiguous().data_ptr(), v.num_tokens,
                                            v.codebook.weight.shape[1], idx.data_ptr(), B * n, self.dp, z.data_ptr(), err.data_ptr(),
                                            _sp(z)))
        if int(err.item()):
            raise IndexError("index out of range in self")          # nn.Embedding's error
        return self._decoder_forward(_Act(z, B, h, w, self.dp), None)[0]

    def _decoder_forward(self, a, save):
        lib = _lib.load()
        v = self.vae
        for l in self.dec[:-1]:
            a = l.fwd(a, save)
        out = self.dec[-1].fwd(a, save, out_f32=True)                  # [B*H*W, cpad] fp32
        img = torch.empty(a.B, v.channels, a.H, a.W, dtype=torch.float32, device=out.device)
        _lib.check(lib.memb_vae_nhwc_to_nchw(out.data_ptr(), a.B, v.channels, a.H, a.W, out.shape[1], img.data_ptr(), _sp(out)))
        return img, out

    # ------------------------------------------------------------------ training forward
    def forward(self, img, temp, noise, need_grad):
        """Returns (loss 0-d fp32, recons fp32 [B,C,H,W], ctx for backward)."""
        lib = _lib.load()
        v = self.vae
        dev = img.device
        self._pack(dev)
        B, C, H, W = img.shape
        cpad = _pad8(C)
        save = [] if need_grad else None
        x = torch.empty(B * H * W, cpad, dtype=torch.bfloat16, device=dev)
        target = torch.empty(B * H * W, C, dtype=torch.float32, device=dev)
        mean = std = None
        if v.normalization is not None:
            mean, std = (torch.as_tensor(t, dtype=torch.float32, device=dev).contiguous() for t in v.normalization)
        img = img.contiguous().float()
        _lib.check(lib.memb_vae_nchw_to_nhwc(img.data_ptr(), B, C, H, W, cpad, ops._ptr(mean), ops._ptr(std), x.data_ptr(),
                                             target.data_ptr(), _sp(x)))
        a = _Act(x, B, H, W, cpad)
        for l in self.enc[:-1]:
            a = l.fwd(a, save)
        logits = self.enc[-1].fwd(a, save, out_f32=True)               # [B*h*w, num_tokens] fp32 (num_tokens % 8 == 0)
        h, w, N = a.H, a.W, v.num_tokens
        rows = B * h * w
        # Gumbel noise in the reference's memory order ([B, N, h, w], F.gumbel_softmax on the NCHW logits) so that the
        # same torch seed gives the same sample; the kernels read it token-major
        if noise is None:
            noise = -torch.empty(B, N, h, w, dtype=torch.float32, device=dev).exponential_().log()
        g = noise.permute(0, 2, 3, 1).reshape(rows, N).contiguous()
        y = torch.empty(rows, N, dtype=torch.bfloat16, device=dev)
        y_fwd = torch.empty_like(y) if v.straight_through else None
        lse = torch.empty(rows, 2, dtype=torch.float32, device=dev)
        scal = torch.zeros(2, dtype=torch.float32, device=dev)          # [kl, recon loss]
        _lib.check(lib.memb_vae_gumbel_fwd(logits.data_ptr(), g.data_ptr(), rows, N, float(temp), int(v.straight_through),
                                           y.data_ptr(), ops._ptr(y_fwd), lse.data_ptr(), scal.data_ptr(), _sp(y)))
        y_used = y_fwd if v.straight_through else y
        z = ops.gemm(y_used, self.cb, b_layout=1, out_dtype=torch.bfloat16)           # einsum('b n h w, n d -> b d h w')
        recons, out = self._decoder_forward(_Act(z, B, h, w, self.dp), save)
        dout = torch.empty(out.shape, dtype=torch.bfloat16, device=dev) if need_grad else None
        _lib.check(lib.memb_vae_recon_loss(target.data_ptr(), out.data_ptr(), B * H * W, C, out.shape[1], self.kind,
                                           scal[1:].data_ptr(), ops._ptr(dout), _sp(out)))
        loss = scal[1] + scal[0] * float(v.kl_div_loss_weight)
        ctx = dict(save=save, logits=logits, g=g, lse=lse, y_used=y_used, z=z, dout=dout, temp=float(temp), rows=rows) if need_grad else None
        return loss, recons, ctx

    def backward(self, ctx):
        """Parameter gradients of the loss (upstream gradient 1), in ``self.params()`` order."""
        lib = _lib.load()
        v = self.vae
        save, grads = ctx["save"], {}
        n_enc = len(self.enc)
        enc_saved, dec_saved = save[:n_enc], save[n_enc:]
        d = ctx["dout"]
        for l, sv in zip(reversed(self.dec), reversed(dec_saved)):
            d = l.bwd(d, sv, grads)
        dz = d                                                             # [rows, dp] bf16
        dev = dz.device
        N = v.num_tokens
        gcb = ops.gemm(ctx["y_used"], dz, a_layout=1, b_layout=1, epilogue=EPI_ATOMIC_ADD)      # d codebook = y^T dz  [N, dp]
        dy = ops.gemm(dz, self.cb, out_dtype=torch.float32)                                     # [rows, N]
        dlogits = torch.empty(ctx["rows"], N, dtype=torch.bfloat16, device=dev)
        _lib.check(lib.memb_vae_gumbel_bwd(ctx["logits"].data_ptr(), ctx["g"].data_ptr(), ctx["lse"].data_ptr(), dy.data_ptr(),
                                           ctx["rows"], N, ctx["temp"], None, float(v.kl_div_loss_weight), dlogits.data_ptr(), _sp(dz)))
        d = dlogits
        for i, (l, sv) in enumerate(zip(reversed(self.enc), reversed(enc_saved))):
            d = l.bwd(d, sv, grads, need_dx=i < n_enc - 1)
        by_param = {id(v.codebook.weight): gcb[:, :v.codebook.weight.shape[1]]}
        for l in self.enc + self.dec:
            for p, gval in l.param_grads(grads):
                by_param[id(p)] = gval
        return [by_param[id(p)] for p in self.params()]


class _VaeLossFn(torch.autograd.Function):
    """(loss, recons) = DiscreteVAE.forward(img, return_loss=True, return_recons=True) with a kernel backward."""

    @staticmethod
    def forward(ctx, trainer, img, temp, noise, *params):
        need_grad = any(ctx.needs_input_grad[4:])
        loss, recons, saved = trainer.forward(img, temp, noise, need_grad)
        ctx.trainer, ctx.saved = trainer, saved
        ctx.mark_non_differentiable(recons)
        return loss, recons

    @staticmethod
    def backward(ctx, grad_loss, _grad_recons):
        if ctx.saved is None:
            raise RuntimeError("mem_b200 dVAE: backward called twice (activations are released after the first pass)")
        grads = ctx.trainer.backward(ctx.saved)
        ctx.saved = None
        out = []
        for p, gval in zip(ctx.trainer.params(), grads):
            out.append((gval * grad_loss).reshape(p.shape).to(p.dtype).contiguous() if p.requires_grad else None)
        return (None, None, None, None) + tuple(out)


def train_forward(vae, img, temp, noise=None):
    _lib.require_cuda()
    if not img.is_cuda:
        raise RuntimeError("mem_b200.DiscreteVAE runs on CUDA tensors only (no CPU path)")
    tr = vae._trainer()
    if tr.kind is None:
        raise NotImplementedError("mem_b200.DiscreteVAE trains with loss='mse' or 'smooth_l1' ('cosine' is not implemented)")
    if not torch.is_grad_enabled():
        loss, recons, _ = tr.forward(img, temp, noise, False)
        return loss, recons
    return _VaeLossFn.apply(tr, img, temp, noise, *tr.params())
