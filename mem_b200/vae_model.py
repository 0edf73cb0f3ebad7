"""Discrete VAE tokenizer of the MEM pretraining step.

Drop-in for ``eventvae/vae/vae_model.py`` (``DiscreteVAE`` :45-213, ``ResBlock`` :29-41) on the path
the pretraining engine uses: ``get_codebook_indices(images)`` (:153-158) and
``forward(img, return_logits=True)`` (:182-189).  Same constructor signature and ``state_dict`` keys
(``codebook.weight``, ``encoder.{i}.0.*``, ``encoder.{j}.net.{0,2,4}.*``, ``decoder.*``) so that
checkpoints written by the reference's ``train_vae.py`` load unchanged.

The encoder runs on libmemb's fp32-faithful convolution kernel (3xTF32 on tcgen05, see
``csrc/conv.cu``) with the codebook argmax fused into the last convolution's epilogue: logits are never
written to memory on the token path.  dVAE *training* (gumbel-softmax + decoder, stage 1 of the
reference pipeline) and ``decode`` are outside the pretraining hot path.
"""
from __future__ import annotations

import ctypes

import torch
from torch import nn

from . import _lib, ops
from ._lib import ConvDesc


class _Container(nn.Module):
    def forward(self, *a, **k):
        raise RuntimeError("dVAE sub-modules are parameter containers; use DiscreteVAE.get_codebook_indices")


class ResBlock(_Container):
    def __init__(self, chan):
        super().__init__()
        self.net = nn.Sequential(nn.Conv2d(chan, chan, 3, padding=1), nn.ReLU(), nn.Conv2d(chan, chan, 3, padding=1),
                                 nn.ReLU(), nn.Conv2d(chan, chan, 1))


class DiscreteVAE(nn.Module):
    def __init__(self, input_H=256, input_W=256, num_tokens=512, codebook_dim=512, num_layers=3, num_resnet_blocks=0,
                 hidden_dim=64, channels=3, loss="mse", temperature=0.9, straight_through=False, kl_div_loss_weight=0.0,
                 normalization=None):
        super().__init__()
        assert input_H % (2 ** num_layers) == 0 and input_W % (2 ** num_layers) == 0, \
            "input size has to be divisible by num_layers"
        assert num_layers >= 1, "number of layers must be greater than or equal to 1"
        assert loss in ("mse", "smooth_l1", "cosine")
        self.input_H, self.input_W, self.input_size = input_H, input_W, (input_H, input_W)
        self.num_tokens, self.num_layers = num_tokens, num_layers
        self.num_resnet_blocks, self.hidden_dim, self.channels = num_resnet_blocks, hidden_dim, channels
        self.temperature, self.straight_through = temperature, straight_through
        self.kl_div_loss_weight, self.normalization = kl_div_loss_weight, normalization
        # Module construction order follows the reference (codebook; encoder/decoder stages interleaved; decoder
        # ResBlock before encoder ResBlock; decoder stem; encoder head; decoder head) so that the same
        # torch.manual_seed gives the same random-init tokenizer.
        self.codebook = nn.Embedding(num_tokens, codebook_dim)
        has_res = num_resnet_blocks > 0
        enc_io = [channels] + [hidden_dim] * num_layers
        dec_io = [hidden_dim if has_res else codebook_dim] + [hidden_dim] * num_layers
        enc, dec = [], []
        for i in range(num_layers):
            enc.append(nn.Sequential(nn.Conv2d(enc_io[i], enc_io[i + 1], 4, stride=2, padding=1), nn.ReLU()))
            dec.append(nn.Sequential(nn.ConvTranspose2d(dec_io[i], dec_io[i + 1], 4, stride=2, padding=1), nn.ReLU()))
        for _ in range(num_resnet_blocks):
            dec.insert(0, ResBlock(dec_io[1]))
            enc.append(ResBlock(enc_io[-1]))
        if has_res:
            dec.insert(0, nn.Conv2d(codebook_dim, dec_io[1], 1))
        enc.append(nn.Conv2d(enc_io[-1], num_tokens, 1))
        dec.append(nn.Conv2d(dec_io[-1], channels, 1))
        self.encoder = nn.Sequential(*enc)
        self.decoder = nn.Sequential(*dec)
        self._tok = None

    # ------------------------------------------------------------------ reference API
    @torch.no_grad()
    def get_codebook_indices(self, images):
        """``logits.argmax(dim=1).flatten(1)`` of the encoder (vae_model.py:153-158): int64 [B, h*w]."""
        return self._tokenizer().run(images, want_logits=False)

    def forward(self, img, return_loss=False, return_recons=False, return_logits=False, temp=None):
        if not return_logits:
            raise NotImplementedError(
                "mem_b200.DiscreteVAE implements the tokenizer path (return_logits=True / get_codebook_indices); "
                "dVAE training is outside the MEM pretraining hot path")
        with torch.no_grad():
            return self._tokenizer().run(img, want_logits=True)

    def decode(self, img_seq):
        raise NotImplementedError("DiscreteVAE.decode (visualisation branch) is outside the MEM pretraining hot path")

    def norm(self, images):
        if self.normalization is None:
            return images
        means, stds = (torch.as_tensor(t).to(images).view(1, -1, 1, 1) for t in self.normalization)
        return (images - means) / stds

    def _tokenizer(self):
        if self._tok is None:
            object.__setattr__(self, "_tok", _Tokenizer(self))
        return self._tok


def _out_layout(kind, C, OH, OW):
    """(sB, sy_major, sy_minor, sx_major, sx_minor, pad, shift, slots shape) of a conv output written for its consumer."""
    if kind == "s2d":      # next: 4x4/s2/p1 conv -> space-to-depth(2) of the zero-padded map, slot inner = (dy, dx, c)
        h2, w2 = OH // 2 + 1, OW // 2 + 1
        return dict(sB=h2 * w2 * 4 * C, sy_major=w2 * 4 * C, sy_minor=2 * C, sx_major=4 * C, sx_minor=C, pad=1, shift=1,
                    slots=(h2, w2, 4 * C))
    if kind == "pad":      # next: 3x3/p1 conv (or a 1x1 conv reading at tap offset (1,1))
        return dict(sB=(OH + 2) * (OW + 2) * C, sy_major=(OW + 2) * C, sy_minor=0, sx_major=C, sx_minor=0, pad=1, shift=0,
                    slots=(OH + 2, OW + 2, C))
    raise ValueError(kind)


class _Tokenizer:
    """Kernel schedule of the encoder: im2col(l1) -> L conv stages -> R residual blocks -> head GEMM + argmax."""

    def __init__(self, vae: DiscreteVAE, chunk: int = 64):
        self.vae, self.chunk = vae, chunk
        self.seg_kblocks = 2   # K blocks per tensor-core accumulation segment (csrc/conv.cu; profiles/r01_dvae_probe_seg_sweep.json)
        self.packed_version = None
        self.bufs = {}

    # ---- weights: [Cout][2K] hi | lo, K ordered to match the activation slot layouts
    def _pack(self, w2d, device):
        lib = _lib.load()
        w2d = w2d.detach().to(device=device, dtype=torch.float32).contiguous()
        out = torch.empty(w2d.shape[0], 2 * w2d.shape[1], dtype=torch.float32, device=device)
        hi, lo = torch.empty_like(w2d), torch.empty_like(w2d)
        _lib.check(lib.memb_split_tf32(w2d.data_ptr(), hi.data_ptr(), lo.data_ptr(), w2d.numel(), _lib.stream_ptr(torch, device)))
        out[:, :w2d.shape[1]] = hi
        out[:, w2d.shape[1]:] = lo
        return out

    def _prepare(self, device):
        v = self.vae
        version = tuple(p._version for p in v.encoder.parameters()) + (str(device),)
        if version == self.packed_version:
            return
        L, R, Hd = v.num_layers, v.num_resnet_blocks, v.hidden_dim
        assert Hd % 32 == 0 and v.num_tokens % 32 == 0, "hidden_dim / num_tokens must be multiples of 32 for the tcgen05 conv kernel"
        f32 = lambda t: t.detach().to(device=device, dtype=torch.float32).contiguous()  # noqa: E731
        self.w, self.b = [], []
        conv = v.encoder[0][0]
        C = conv.in_channels
        self.kpad = (16 * C + 31) // 32 * 32
        w0 = torch.zeros(Hd, self.kpad, device=device)
        w0[:, :16 * C] = f32(conv.weight).reshape(Hd, 16 * C)
        self.w.append(self._pack(w0, device)); self.b.append(f32(conv.bias))
        for i in range(1, L):  # ky = 2a+dy, kx = 2b+dx  ->  K order (a, b, dy, dx, c)
            conv = v.encoder[i][0]
            w = f32(conv.weight).view(Hd, Hd, 2, 2, 2, 2).permute(0, 2, 4, 3, 5, 1).reshape(Hd, 16 * Hd)
            self.w.append(self._pack(w, device)); self.b.append(f32(conv.bias))
        for j in range(R):
            net = v.encoder[L + j].net
            for idx in (0, 2):
                self.w.append(self._pack(f32(net[idx].weight).permute(0, 2, 3, 1).reshape(Hd, 9 * Hd), device))
                self.b.append(f32(net[idx].bias))
            self.w.append(self._pack(f32(net[4].weight).reshape(Hd, Hd), device)); self.b.append(f32(net[4].bias))
        head = v.encoder[L + R]
        self.w_head = self._pack(f32(head.weight).reshape(v.num_tokens, Hd), device)
        self.b_head = f32(head.bias)
        if v.normalization is not None:
            self.mean, self.std = (torch.as_tensor(t, dtype=torch.float32, device=device).contiguous() for t in v.normalization)
        else:
            self.mean = self.std = None
        self.packed_version = version

    def _buf(self, name, shape, device, dtype=torch.float32):
        key = (name, str(device))
        t = self.bufs.get(key)
        if t is None or tuple(t.shape) != tuple(shape) or t.dtype != dtype:
            t = torch.zeros(shape, dtype=dtype, device=device)   # zero borders are part of the layout
            self.bufs[key] = t
        return t

    def _conv(self, lib, a, geom, w, bias, B, OH, OW, relu, out_kind, out_name, device, aux=None, full=None, keys=None):
        """a = (hi, lo) tensors [B*rows_per_img, x_slots, inner]; geom = (taps_y, taps_x, tap_y0, tap_x0).
        out_kind: slot layout of the hi/lo result for its consumer, or None (only ``full`` / ``keys`` outputs)."""
        Cout = w.shape[0]
        d = ConvDesc()
        out = None
        if out_kind is not None:
            lay = _out_layout(out_kind, Cout, OH, OW)
            sh = lay["slots"]
            hi = self._buf(out_name + "_hi", (B * sh[0], sh[1], sh[2]), device)
            lo = self._buf(out_name + "_lo", (B * sh[0], sh[1], sh[2]), device)
            d.d_hi, d.d_lo, out = hi.data_ptr(), lo.data_ptr(), (hi, lo)
            d.sB, d.sy_major, d.sy_minor, d.sx_major, d.sx_minor = lay["sB"], lay["sy_major"], lay["sy_minor"], lay["sx_major"], lay["sx_minor"]
            d.pad, d.shift = lay["pad"], lay["shift"]
        d.a_hi, d.a_lo = a[0].data_ptr(), a[1].data_ptr()
        d.r_slots, d.x_slots, d.inner = a[0].shape
        d.rows_per_img = a[0].shape[0] // B
        d.taps_y, d.taps_x, d.tap_y0, d.tap_x0 = geom
        d.w, d.bias = w.data_ptr(), bias.data_ptr()
        assert w.shape[1] == 2 * geom[0] * geom[1] * d.inner, "weight K does not match the activation layout"
        d.B, d.OH, d.OW, d.Cout, d.relu = B, OH, OW, Cout, int(relu)
        d.aux, d.d_full, d.keys = ops._ptr(aux), ops._ptr(full), ops._ptr(keys)
        d.seg_kblocks = self.seg_kblocks
        d.err_flag = ops._err_flag(torch, device).data_ptr()
        _lib.check(lib.memb_conv_tf32x3(ctypes.byref(d), _lib.stream_ptr(torch, device)))
        return out

    def run(self, images, want_logits):
        _lib.require_cuda()
        if not images.is_cuda:
            raise RuntimeError("mem_b200.DiscreteVAE runs on CUDA tensors only (no CPU path)")
        v = self.vae
        assert images.shape[-1] == v.input_W and images.shape[-2] == v.input_H, \
            f"input must have the correct image size {v.input_H}x{v.input_W}, but is ({images.shape[-2]},{images.shape[-1]})"
        assert images.shape[1] == v.encoder[0][0].in_channels, "channel count does not match the tokenizer"
        device = images.device
        self._prepare(device)
        images = images.contiguous().float()
        Btot = images.shape[0]
        h, w = v.input_H >> v.num_layers, v.input_W >> v.num_layers
        tokens = torch.empty(Btot, h * w, dtype=torch.int64, device=device)
        logits = torch.empty(Btot, h * w, v.num_tokens, dtype=torch.float32, device=device) if want_logits else None
        for b0 in range(0, Btot, self.chunk):
            b1 = min(Btot, b0 + self.chunk)
            self._run_chunk(images[b0:b1], tokens[b0:b1], logits[b0:b1] if want_logits else None)
        if want_logits:
            return logits.view(Btot, h, w, v.num_tokens).permute(0, 3, 1, 2)
        return tokens

    def _run_chunk(self, img, tokens, logits):
        lib = _lib.load()
        v = self.vae
        device = img.device
        sp = _lib.stream_ptr(torch, device)
        B, C, H, W = img.shape
        L, R, Hd = v.num_layers, v.num_resnet_blocks, v.hidden_dim
        tag = f"B{B}_"
        # ---- layer 1: explicit im2col (K = 16*C is tiny), then a one-tap "conv" over the plain matrix
        OH, OW = H // 2, W // 2
        a_hi = self._buf(tag + "a1_hi", (B * OH, OW, self.kpad), device)
        a_lo = self._buf(tag + "a1_lo", (B * OH, OW, self.kpad), device)
        _lib.check(lib.memb_dvae_im2col_l1(img.data_ptr(), B, C, H, W, self.kpad, ops._ptr(self.mean), ops._ptr(self.std),
                                           a_hi.data_ptr(), a_lo.data_ptr(), sp))

        def consumer(stage):  # layout wanted by whatever reads the output of conv stage `stage` (0-based)
            return "s2d" if stage + 1 < L else "pad"

        x_full = self._buf(tag + "x_full", (B * (H >> L) * (W >> L), Hd), device) if R > 0 else None
        cur = self._conv(lib, (a_hi, a_lo), (1, 1, 0, 0), self.w[0], self.b[0], B, OH, OW, True, consumer(0), tag + "act0", device,
                         full=x_full if (L == 1 and R > 0) else None)
        for i in range(1, L):
            OH, OW = OH // 2, OW // 2
            cur = self._conv(lib, cur, (2, 2, 0, 0), self.w[i], self.b[i], B, OH, OW, True, consumer(i), tag + f"act{i}", device,
                             full=x_full if (i == L - 1 and R > 0) else None)
        # ---- residual blocks: x + conv1x1(relu(conv3x3(relu(conv3x3(x)))))
        wi = L
        for j in range(R):
            t1 = self._conv(lib, cur, (3, 3, 0, 0), self.w[wi], self.b[wi], B, OH, OW, True, "pad", tag + "res_t1", device)
            t2 = self._conv(lib, t1, (3, 3, 0, 0), self.w[wi + 1], self.b[wi + 1], B, OH, OW, True, "pad", tag + "res_t2", device)
            cur = self._conv(lib, t2, (1, 1, 1, 1), self.w[wi + 2], self.b[wi + 2], B, OH, OW, False, "pad", tag + f"act{L - 1}",
                             device, aux=x_full, full=x_full)
            wi += 3
        # ---- head: 1x1 conv to num_tokens with the codebook argmax in the epilogue (logits only on request)
        rows = B * OH * OW
        if logits is not None:
            self._conv(lib, cur, (1, 1, 1, 1), self.w_head, self.b_head, B, OH, OW, False, None, None, device,
                       full=logits.view(rows, v.num_tokens))
        keys = self._buf(tag + "keys", (rows,), device, torch.int64)
        keys.zero_()
        self._conv(lib, cur, (1, 1, 1, 1), self.w_head, self.b_head, B, OH, OW, False, None, None, device, keys=keys)
        _lib.check(lib.memb_argmax_decode(keys.data_ptr(), tokens.data_ptr(), rows, sp))
