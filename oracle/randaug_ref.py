"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): CPU restatement of the reference's ``EventRandAugment``
(mem/transforms.py:351-471, ``_apply_op`` :291-331) on uint8 ``[3, H, W]`` event images.

The reference applies, per sample, ``num_ops`` operations drawn from the global torch generator (three ``torch.randint``
calls per operation: operation index, magnitude bin, sign) through torchvision's functional API.  torchvision is a
third-party dependency that is not vendored in /root/reference (the image has 0.26.0); its tensor code path
(``torchvision/transforms/_functional_tensor.py``) is restated here operation by operation, float32 operation by float32
operation, and pinned by ``tests/golden/randaug.npz`` -- outputs of the UNMODIFIED reference class running on torchvision
in the build container (``oracle/make_golden.py::golden_randaug``).

Exactness: every photometric operation is reproduced bit for bit (each float32 product / sum of the reference is one
correctly rounded operation here; the blur of ``Sharpness`` can never sit on a rounding boundary: S/13 is never k + 1/2).
The geometric operations (ShearX/Y, Rotate; TranslateX/Y are exact) go through ``affine_grid`` (a float32 GEMM whose summation
order is the BLAS's) and ``grid_sample``; the restatement fixes one order (documented below), so a handful of pixels whose
interpolated value lands within float32 noise of k + 1/2 can differ by one count from the reference.
"""
from __future__ import annotations

import math

import numpy as np

OPS = ["Identity", "ShearX", "ShearY", "TranslateX", "TranslateY", "Rotate", "Brightness", "Color", "Contrast", "Sharpness",
       "Posterize", "Solarize", "AutoContrast", "Equalize"]
OP_ID = {n: i for i, n in enumerate(OPS)}
SMALL = OPS[:11]
f32 = np.float32


def augmentation_space(names, num_bins, height, width):
    """``EventRandAugment._augmentation_space`` (transforms.py:411-430): name -> (float32 magnitudes | None, signed).
    The tables are ``torch.linspace`` values (its vectorised float32 evaluation is not ``start + i * step``), so they
    are taken from torch itself."""
    import torch
    lin = lambda a, b: torch.linspace(a, b, num_bins).numpy()  # noqa: E731
    post = (8 - (torch.arange(num_bins) / ((num_bins - 1) / 4)).round().int()).numpy()
    d = {
        "Identity": (None, False),
        "ShearX": (lin(0.0, 0.3), True), "ShearY": (lin(0.0, 0.3), True),
        "TranslateX": (lin(0.0, 150.0 / 331.0 * width), True), "TranslateY": (lin(0.0, 150.0 / 331.0 * height), True),
        "Rotate": (lin(0.0, 30.0), True),
        "Brightness": (lin(0.0, 0.9), True), "Color": (lin(0.0, 0.9), True), "Contrast": (lin(0.0, 0.9), True),
        "Sharpness": (lin(0.0, 0.9), True),
        "Posterize": (post, False), "Solarize": (lin(255.0, 0.0), False),
        "AutoContrast": (None, False), "Equalize": (None, False),
    }
    return {k: d[k] for k in d if k in names}


def draw_ops(gen_randint, names, num_ops, magnitude, num_bins, height, width):
    """The reference's draws, in its order (transforms.py:447-462): per operation ``randint(len(space))``,
    ``randint(magnitude + 1)``, ``randint(2)``.  ``gen_randint(n)`` returns one draw in ``[0, n)``.
    Returns ``[(name, magnitude as a Python float)]``."""
    space = augmentation_space(names, num_bins, height, width)
    keys = list(space.keys())
    out = []
    for _ in range(num_ops):
        name = keys[gen_randint(len(keys))]
        mags, signed = space[name]
        i0 = gen_randint(magnitude + 1)
        i1 = gen_randint(2)
        mag = float(mags[i0]) if mags is not None else 0.0
        if signed and i1:
            mag *= -1.0
        out.append((name, mag))
    return out


def inverse_affine_matrix(angle, translate, shear):
    """torchvision ``_get_inverse_affine_matrix(center=[0, 0], angle, translate, scale=1, shear)`` (functional.py:1006-1063)."""
    rot, sx, sy = math.radians(angle), math.radians(shear[0]), math.radians(shear[1])
    tx, ty = translate
    a = math.cos(rot - sy) / math.cos(sy)
    b = -math.cos(rot - sy) * math.tan(sx) / math.cos(sy) - math.sin(rot)
    c = math.sin(rot - sy) / math.cos(sy)
    d = -math.sin(rot - sy) * math.tan(sx) / math.cos(sy) + math.cos(rot)
    m = [d, -b, 0.0, -c, a, 0.0]
    m = [x / 1.0 for x in m]
    m[2] += m[0] * (-0.0 - tx) + m[1] * (-0.0 - ty)
    m[5] += m[3] * (-0.0 - tx) + m[4] * (-0.0 - ty)
    m[2] += 0.0
    m[5] += 0.0
    return m


def op_matrix(name, mag):
    """The 2x3 inverse matrix ``_apply_op`` hands to the tensor backend for a geometric operation (transforms.py:293-308;
    ``F.affine`` / ``F.rotate`` with a tensor input use center (0, 0); rotate negates the angle)."""
    if name == "ShearX":
        return inverse_affine_matrix(0.0, [0.0, 0.0], [math.degrees(mag), 0.0])
    if name == "ShearY":
        return inverse_affine_matrix(0.0, [0.0, 0.0], [0.0, math.degrees(mag)])
    if name == "TranslateX":
        return inverse_affine_matrix(0.0, [1.0 * int(mag), 0.0], [0.0, 0.0])
    if name == "TranslateY":
        return inverse_affine_matrix(0.0, [0.0, 1.0 * int(mag)], [0.0, 0.0])
    if name == "Rotate":
        return inverse_affine_matrix(-mag, [0.0, 0.0], [0.0, 0.0])
    raise ValueError(name)


def _affine_bilinear(img, matrix):
    """``_gen_affine_grid`` + ``grid_sample(bilinear, zeros, align_corners=False)`` + round (functional_tensor.py:545-618),
    float32.  Fixed evaluation order: g = (xb * t0 + yb * t1) + t2 per axis, corner sum nw + ne + sw + se."""
    C, H, W = img.shape
    th = np.asarray(matrix, dtype=f32).reshape(2, 3)
    rt = np.empty((3, 2), dtype=f32)                       # theta^T / [0.5 w, 0.5 h]
    rt[:, 0] = th[0] / f32(0.5 * W)
    rt[:, 1] = th[1] / f32(0.5 * H)
    xb = (np.arange(W, dtype=f32) + f32(-W * 0.5 + 0.5)).astype(f32)[None, :]      # linspace with step exactly 1
    yb = (np.arange(H, dtype=f32) + f32(-H * 0.5 + 0.5)).astype(f32)[:, None]
    gx = ((xb * rt[0, 0]).astype(f32) + (yb * rt[1, 0]).astype(f32)).astype(f32) + rt[2, 0]
    gy = ((xb * rt[0, 1]).astype(f32) + (yb * rt[1, 1]).astype(f32)).astype(f32) + rt[2, 1]
    ix = (((gx + f32(1)) * f32(W)).astype(f32) - f32(1)).astype(f32) / f32(2)
    iy = (((gy + f32(1)) * f32(H)).astype(f32) - f32(1)).astype(f32) / f32(2)
    x0, y0 = np.floor(ix), np.floor(iy)
    x1, y1 = x0 + f32(1), y0 + f32(1)
    wnw = ((x1 - ix) * (y1 - iy)).astype(f32)
    wne = ((ix - x0) * (y1 - iy)).astype(f32)
    wsw = ((x1 - ix) * (iy - y0)).astype(f32)
    wse = ((ix - x0) * (iy - y0)).astype(f32)
    src = img.astype(f32)
    out = np.zeros((C, H, W), dtype=f32)

    def corner(xi, yi, w):
        ok = (xi >= 0) & (xi < W) & (yi >= 0) & (yi < H)
        xs = np.clip(xi, 0, W - 1).astype(np.int64)
        ys = np.clip(yi, 0, H - 1).astype(np.int64)
        return np.where(ok[None], (src[:, ys, xs] * w[None]).astype(f32), f32(0))

    for xi, yi, w in ((x0, y0, wnw), (x1, y0, wne), (x0, y1, wsw), (x1, y1, wse)):
        out = (out + corner(xi, yi, w)).astype(f32)
    return np.rint(out).astype(np.uint8)                    # torch.round: half to even


def _gray(img):
    r, g, b = (img[c].astype(f32) for c in range(3))
    v = ((f32(0.2989) * r).astype(f32) + (f32(0.587) * g).astype(f32)).astype(f32)
    v = (v + (f32(0.114) * b).astype(f32)).astype(f32)
    return v.astype(np.uint8)                                # .to(uint8): truncation


def _blend(img, other, ratio):
    """``(ratio * img1 + (1.0 - ratio) * img2).clamp(0, 255).to(uint8)``; ``other``: uint8 array or float32 scalar."""
    r0, r1 = f32(float(ratio)), f32(1.0 - float(ratio))
    a = (r0 * img.astype(f32)).astype(f32)
    b = (r1 * np.asarray(other, dtype=f32)).astype(f32)
    return np.clip((a + b).astype(f32), 0, 255).astype(np.uint8)


def _blur(img):
    C, H, W = img.shape
    k = np.ones((3, 3), dtype=f32)
    k[1, 1] = 5.0
    k = (k / k.sum(dtype=f32)).astype(f32)
    out = img.copy()
    src = img.astype(f32)
    acc = np.zeros((C, H - 2, W - 2), dtype=f32)
    for dy in range(3):
        for dx in range(3):
            acc = (acc + (src[:, dy:dy + H - 2, dx:dx + W - 2] * k[dy, dx]).astype(f32)).astype(f32)
    out[:, 1:-1, 1:-1] = np.rint(acc).astype(np.uint8)
    return out


def _equalize_channel(ch):
    hist = np.bincount(ch.reshape(-1), minlength=256).astype(np.int64)
    nz = hist[hist != 0]
    step = int(nz[:-1].sum()) // 255
    if step == 0:
        return ch
    lut = (np.cumsum(hist) + step // 2) // step
    lut = np.clip(np.concatenate([[0], lut[:-1]]), 0, 255)
    return lut[ch].astype(np.uint8)


def apply_op(img, name, mag):
    """``_apply_op(img, op_name, magnitude, BILINEAR, fill=None)`` (transforms.py:291-331) on uint8 [3, H, W]."""
    assert img.dtype == np.uint8 and img.ndim == 3 and img.shape[0] == 3
    if name == "Identity":
        return img
    if name in ("ShearX", "ShearY", "TranslateX", "TranslateY", "Rotate"):
        return _affine_bilinear(img, op_matrix(name, mag))
    if name == "Brightness":
        return _blend(img, np.zeros_like(img), 1.0 + mag)
    if name == "Color":
        return _blend(img, _gray(img)[None], 1.0 + mag)
    if name == "Contrast":
        g = _gray(img)
        mean = f32(f32(int(g.astype(np.int64).sum())) / f32(g.size))
        return _blend(img, mean, 1.0 + mag)
    if name == "Sharpness":
        if img.shape[1] <= 2 or img.shape[2] <= 2:
            return img
        return _blend(img, _blur(img), 1.0 + mag)
    if name == "Posterize":
        return img & np.uint8((-int(2 ** (8 - int(mag)))) & 0xff)
    if name == "Solarize":
        return np.where(img.astype(f32) >= f32(mag), 255 - img, img).astype(np.uint8)
    if name == "AutoContrast":
        lo = img.reshape(3, -1).min(1).astype(f32)
        hi = img.reshape(3, -1).max(1).astype(f32)
        with np.errstate(divide="ignore"):
            scale = (f32(255.0) / (hi - lo)).astype(f32)
        bad = ~np.isfinite(scale)
        lo[bad], scale[bad] = 0, 1
        v = ((img.astype(f32) - lo[:, None, None]).astype(f32) * scale[:, None, None]).astype(f32)
        return np.clip(v, 0, 255).astype(np.uint8)
    if name == "Equalize":
        return np.stack([_equalize_channel(img[c]) for c in range(3)])
    raise ValueError(name)


def rand_augment(img, ops):
    """img uint8 [3, H, W]; ops: [(name, magnitude)] as returned by ``draw_ops``."""
    for name, mag in ops:
        img = apply_op(img, name, mag)
    return img


def to_uint8(x):
    """``ToUnit8`` (transforms.py:343-349): ``(255 * x).to(torch.uint8)`` on a float32 tensor."""
    return (f32(255) * np.asarray(x, dtype=f32)).astype(f32).astype(np.uint8)


def to_float32(x):
    """``ToFloat32`` (transforms.py:333-340): ``x.to(torch.float32) / 255``."""
    return (np.asarray(x).astype(f32) / f32(255)).astype(f32)
