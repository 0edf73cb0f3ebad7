"""Image-space augmentation at the end of the event pipeline.

Drop-in for ``mem/transforms.py``: ``EventRandAugment`` (:351-471), ``ToUnit8`` (:343-349), ``ToFloat32`` (:333-340) -- the
three transforms ``build_transformNPY`` appends when ``args.rand_aug`` is set (mem/datasets.py:655-658; the reference's run
scripts default to ``--rand_aug 1``).  Same constructor arguments, same augmentation space, and the same three
``torch.randint`` draws per operation from the same generator, so a run with the same torch seed picks the same operations
and magnitudes as the reference.

The reference applies the operations per sample on the CPU through torchvision; here the HOST only draws them
(``draw``), and one kernel (``csrc/randaug.cu``: one CTA per sample, the image in shared memory) applies all operations of a
whole batch, with ToUnit8 / ToFloat32 folded into its load and store (``augment_batch``).  CUDA tensors only.
"""
from __future__ import annotations

import copy
import ctypes
import math
from typing import List, Optional

import numpy as np
import torch

from . import _lib

ALL = ["Identity", "ShearX", "ShearY", "TranslateX", "TranslateY", "Rotate", "Brightness", "Color", "Contrast", "Sharpness",
       "Posterize", "Solarize", "AutoContrast", "Equalize"]
SMALL = ALL[:11]
RA_IDENTITY, RA_AFFINE, RA_BRIGHTNESS, RA_COLOR, RA_CONTRAST, RA_SHARPNESS, RA_POSTERIZE, RA_SOLARIZE, RA_AUTOCONTRAST, RA_EQUALIZE = range(10)
_BLEND = {"Brightness": RA_BRIGHTNESS, "Color": RA_COLOR, "Contrast": RA_CONTRAST, "Sharpness": RA_SHARPNESS}
OP_DTYPE = np.dtype([("op", "<i4"), ("ival", "<i4"), ("f0", "<f4"), ("f1", "<f4"), ("theta", "<f4", (6,))])   # memb_randaug_op
assert OP_DTYPE.itemsize == 40


class ToFloat32:
    """uint8 in (0, 255) -> float32 in (0, 1) (transforms.py:333-340)."""

    def __call__(self, x):
        return x.to(torch.float32) / 255


class ToUnit8:
    """float32 in (0, 1) -> uint8 (transforms.py:343-349; the reference's spelling)."""

    def __call__(self, x):
        return (255 * x).to(torch.uint8)


def _inverse_affine_matrix(angle, translate, shear):
    """torchvision's ``_get_inverse_affine_matrix`` for center (0, 0) and scale 1 (what ``F.affine`` / ``F.rotate`` use for
    tensors), in Python doubles like the original."""
    rot, sx, sy = math.radians(angle), math.radians(shear[0]), math.radians(shear[1])
    tx, ty = translate
    a = math.cos(rot - sy) / math.cos(sy)
    b = -math.cos(rot - sy) * math.tan(sx) / math.cos(sy) - math.sin(rot)
    c = math.sin(rot - sy) / math.cos(sy)
    d = -math.sin(rot - sy) * math.tan(sx) / math.cos(sy) + math.cos(rot)
    m = [d, -b, 0.0, -c, a, 0.0]
    m[2] += m[0] * (-tx) + m[1] * (-ty)
    m[5] += m[3] * (-tx) + m[4] * (-ty)
    return m


def encode_op(name: str, magnitude: float) -> np.void:
    """One ``memb_randaug_op`` record for ``_apply_op(img, name, magnitude, BILINEAR, fill=None)`` (transforms.py:291-331)."""
    rec = np.zeros((), dtype=OP_DTYPE)
    if name == "Identity":
        rec["op"] = RA_IDENTITY
    elif name in ("ShearX", "ShearY", "TranslateX", "TranslateY", "Rotate"):
        rec["op"] = RA_AFFINE
        if name == "ShearX":
            m = _inverse_affine_matrix(0.0, [0.0, 0.0], [math.degrees(magnitude), 0.0])
        elif name == "ShearY":
            m = _inverse_affine_matrix(0.0, [0.0, 0.0], [0.0, math.degrees(magnitude)])
        elif name == "TranslateX":
            m = _inverse_affine_matrix(0.0, [1.0 * int(magnitude), 0.0], [0.0, 0.0])
        elif name == "TranslateY":
            m = _inverse_affine_matrix(0.0, [0.0, 1.0 * int(magnitude)], [0.0, 0.0])
        else:
            m = _inverse_affine_matrix(-magnitude, [0.0, 0.0], [0.0, 0.0])
        rec["theta"] = np.asarray(m, dtype=np.float32)
    elif name in _BLEND:
        ratio = 1.0 + magnitude
        if ratio < 0:
            raise ValueError(f"{name.lower()} factor ({ratio}) is not non-negative.")
        rec["op"], rec["f0"], rec["f1"] = _BLEND[name], np.float32(ratio), np.float32(1.0 - ratio)
    elif name == "Posterize":
        rec["op"], rec["ival"] = RA_POSTERIZE, int(magnitude)
    elif name == "Solarize":
        if magnitude > 255:
            raise TypeError("Threshold should be less than bound of img.")
        rec["op"], rec["f0"] = RA_SOLARIZE, np.float32(magnitude)
    elif name == "AutoContrast":
        rec["op"] = RA_AUTOCONTRAST
    elif name == "Equalize":
        rec["op"] = RA_EQUALIZE
    else:
        raise ValueError("The provided operator {} is not recognized.".format(name))
    return rec


def ops_to_device(ops: np.ndarray, device) -> torch.Tensor:
    """``OP_DTYPE`` records ``[B, num_ops]`` -> the uint8 CUDA tensor ``[B, num_ops, 40]`` ``apply_ops`` also accepts."""
    ops = np.ascontiguousarray(ops, dtype=OP_DTYPE)
    return torch.from_numpy(ops.view(np.uint8).reshape(ops.shape + (OP_DTYPE.itemsize,)).copy()).to(device, non_blocking=True)


def apply_ops(x: torch.Tensor, ops, out_float: Optional[bool] = None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """x: CUDA uint8 or float32 ``[B, 3, H, W]`` (float32 is converted like ``ToUnit8``); ops: ``OP_DTYPE`` array
    ``[B, num_ops]`` (or its device copy from ``ops_to_device``).  Returns uint8, or float32 converted like ``ToFloat32``
    (default: the dtype class of ``x``)."""
    _lib.require_cuda()
    if not x.is_cuda:
        raise RuntimeError("mem_b200.transforms runs on CUDA tensors only (no CPU path)")
    assert x.ndim == 4 and x.shape[1] == 3 and x.dtype in (torch.uint8, torch.float32), "expected uint8 / float32 [B, 3, H, W]"
    B, C, H, W = x.shape
    if isinstance(ops, torch.Tensor):
        assert ops.is_cuda and ops.dtype == torch.uint8 and ops.is_contiguous() and ops.shape[0] == B and ops.shape[-1] == OP_DTYPE.itemsize
        dev_ops, num_ops = ops, (ops.shape[1] if ops.ndim == 3 else 1)
    else:
        ops = np.ascontiguousarray(ops, dtype=OP_DTYPE).reshape(B, -1)
        num_ops = ops.shape[1]
        dev_ops = ops_to_device(ops, x.device) if ops.size else None
    x = x.contiguous()
    out_float = (x.dtype == torch.float32) if out_float is None else out_float
    if out is None:
        out = torch.empty(B, C, H, W, dtype=torch.float32 if out_float else torch.uint8, device=x.device)
    lib = _lib.load()
    _lib.check(lib.memb_event_randaug(x.data_ptr(), int(x.dtype == torch.float32), B, C, H, W,
                                      dev_ops.data_ptr() if dev_ops is not None and num_ops else None, num_ops,
                                      out.data_ptr(), int(out_float), _lib.stream_ptr(torch, x.device)))
    return out


class EventRandAugment(torch.nn.Module):
    """``EventRandAugment`` of the reference (transforms.py:351-471), applied on the device.

    ``forward(img)``: uint8 CUDA ``[3, H, W]`` (the reference's call) or a batch ``[B, 3, H, W]`` -- samples draw their
    operations one after the other in batch order, exactly as consecutive calls of the reference would."""

    def __init__(self, num_ops: int = 2, magnitude: int = 9, num_magnitude_bins: int = 31, interpolation=None,
                 fill: Optional[List[float]] = None, verbose: bool = False, gen: Optional[torch.Generator] = None, small=False) -> None:
        super().__init__()
        if fill is not None:
            raise NotImplementedError("EventRandAugment: only fill=None (what build_transformNPY uses) is implemented")
        if interpolation is not None and getattr(interpolation, "value", interpolation) != "bilinear":
            raise NotImplementedError("EventRandAugment: only BILINEAR interpolation (the reference's default) is implemented")
        self.num_ops, self.magnitude, self.num_magnitude_bins = num_ops, magnitude, num_magnitude_bins
        self.interpolation, self.fill, self.verbose, self.gen = interpolation, fill, verbose, gen
        self.names = list(SMALL) if small else list(ALL)
        print(f"Created RandAug with {self.names}")

    def set_names(self, x):
        """A list of operation names, or "all" / "small" (same behaviour as the reference's method)."""
        presets = {"all": ALL, "small": SMALL}
        if isinstance(x, list):
            self.names = copy.copy(x)
        elif isinstance(x, str) and x in presets:
            self.names = list(presets[x])
        else:
            raise RuntimeError(f"Not implemented for {x}")

    def _augmentation_space(self, num_bins, image_size):
        """name -> (magnitude table, signed), in the reference's key order (the operation index is drawn against it).
        Ranges as in transforms.py:411-430; the tables are ``torch.linspace`` values like the reference's."""
        height, width = image_size
        ranged = {"ShearX": (0.0, 0.3), "ShearY": (0.0, 0.3), "TranslateX": (0.0, 150.0 / 331.0 * width),
                  "TranslateY": (0.0, 150.0 / 331.0 * height), "Rotate": (0.0, 30.0), "Brightness": (0.0, 0.9),
                  "Color": (0.0, 0.9), "Contrast": (0.0, 0.9), "Sharpness": (0.0, 0.9)}
        space = {}
        for name in ALL:
            if name not in self.names:
                continue
            if name in ranged:
                space[name] = (torch.linspace(ranged[name][0], ranged[name][1], num_bins), True)
            elif name == "Posterize":
                space[name] = (8 - (torch.arange(num_bins) / ((num_bins - 1) / 4)).round().int(), False)
            elif name == "Solarize":
                space[name] = (torch.linspace(255.0, 0.0, num_bins), False)
            else:                                   # Identity, AutoContrast, Equalize: no magnitude
                space[name] = (torch.tensor(0.0), False)
        return space

    def draw(self, height: int, width: int):
        """One sample's ``[(name, magnitude)]``: the reference's draws in its order (transforms.py:447-462)."""
        op_meta = self._augmentation_space(self.num_magnitude_bins, (height, width))
        keys = list(op_meta.keys())
        picked = []
        for _ in range(self.num_ops):
            op_name = keys[int(torch.randint(len(op_meta), (1,), generator=self.gen).item())]
            magnitudes, signed = op_meta[op_name]
            randi0 = torch.randint(self.magnitude + 1, (1,), generator=self.gen).item()
            randi1 = torch.randint(2, (1,), generator=self.gen)
            magnitude = float(magnitudes[randi0].item()) if magnitudes.ndim > 0 else 0.0
            if signed and randi1:
                magnitude *= -1.0
            if self.verbose:
                print(f"EventRandAug: {op_name:20s} {magnitude:7.2f}")
            picked.append((op_name, magnitude))
        return picked

    def draw_batch(self, B: int, height: int, width: int) -> np.ndarray:
        ops = np.zeros((B, self.num_ops), dtype=OP_DTYPE)
        for b in range(B):
            for k, (name, mag) in enumerate(self.draw(height, width)):
                ops[b, k] = encode_op(name, mag)
        return ops

    def forward(self, img: torch.Tensor) -> torch.Tensor:
        assert img.dtype == torch.uint8
        single = img.ndim == 3
        x = img.unsqueeze(0) if single else img
        out = apply_ops(x, self.draw_batch(x.shape[0], x.shape[-2], x.shape[-1]))
        return out[0] if single else out

    def augment_batch(self, x: torch.Tensor) -> torch.Tensor:
        """``ToUnit8 -> EventRandAugment -> ToFloat32`` (mem/datasets.py:655-658) on a float32 batch ``[B, 3, H, W]`` in one
        launch: float32 in, float32 out."""
        assert x.dtype == torch.float32
        return apply_ops(x, self.draw_batch(x.shape[0], x.shape[-2], x.shape[-1]), out_float=True)

    def __repr__(self) -> str:
        return (f"{self.__class__.__name__}(num_ops={self.num_ops}, magnitude={self.magnitude}, "
                f"num_magnitude_bins={self.num_magnitude_bins}, interpolation={self.interpolation}, fill={self.fill})")
