"""Oracle (test infrastructure, not product): event-space augmentations + post-raster transforms.

numpy / torch restatement of the per-sample chain ``build_transformNPY`` composes around the rasteriser
(reference ``mem/datasets.py:611-660``), for the fixed-sensor N-ImageNet path:

* ``ReshapeScaleXandY``   (datasets.py:464-485)  ``x *= scale_x; y *= scale_y``
* ``SliceRandomMaxEvs``   (:488-498)  ``random.choice(range(L - max + 1))`` -> contiguous window, only when L > max
* ``RandomTimeFlip``      (:598-608)  ``np.random.random() < p`` -> rows reversed, ``t = t[0] - t``, ``p = -p``
* ``Aug_FlipEvsAlongX``   (:501-521)  ``np.random.random() < p`` -> ``x = W - 1 - x``
* ``Aug_RandomShiftEvs``  (:524-549)  ``np.random.randint(-s, s + 1, size=(2,))`` -> shift, drop rows outside the sensor
* ``EventArrToImg``       (:552-595)  -> ``oracle.histogram_ref``
* torchvision ``ToTensor`` (uint8 HWC -> float32 CHW / 255), ``RandomCrop(size, pad_if_needed=True)``
  (``torch.randint`` for top, then left, when the image is larger than the crop)
* ``RemoveTimesurface``   (mem/transforms.py:239-247), ``RemoveHotPixels(num_stds)`` (:249-275),
  ``NormalizeEvent`` (:225-237)

``draw_params`` consumes the three global generators exactly as the chain above does, in its order; the
``apply_*`` functions are pure.  Pinned against the reference's own classes by ``tests/golden/event_pipeline.npz``
(``oracle/make_golden.py::golden_event_pipeline``).
"""
from __future__ import annotations

import random
from dataclasses import dataclass

import numpy as np
import torch

from .histogram_ref import event_hist_ref


@dataclass
class PipelineCfg:
    """The ``args`` fields build_transformNPY reads, for a fixed sensor (datasets.py:615-621)."""
    is_train: bool = True
    sensor_H: int = 480
    sensor_W: int = 640
    input_H: int = 224
    input_W: int = 224
    slice_max_evs: int = 30000
    max_random_shift_evs: int = 15
    timesurface: bool = False
    hotpixfilter: bool = True
    hotpix_num_stds: float = 10
    normalize_events: bool = False
    rand_aug: bool = False                                     # args.rand_aug: datasets.py:655-658
    logtrafo: bool = False                                     # datasets.py:647-650
    gammatrafo: bool = False
    gamma: float = 0.5

    def scales(self):
        if self.is_train:                                      # datasets.py:472-478
            s = 256 / [self.sensor_H, self.sensor_W][int(np.argmin([self.sensor_H, self.sensor_W]))]
            return s, s
        return self.input_W / self.sensor_W, self.input_H / self.sensor_H

    def raster_hw(self):
        if self.is_train:                                      # datasets.py:618-621
            scale = 256 / 480
            return int(480 * scale), int(640 * scale)
        return self.input_H, self.input_W


def draw_params(n_events: int, cfg: PipelineCfg) -> dict:
    """Consume ``random`` / ``np.random`` / ``torch`` global generators like one pass of the reference chain."""
    H, W = cfg.raster_hw()
    sx, sy = cfg.scales()
    p = dict(scale_x=sx, scale_y=sy, start=0, count=n_events, time_flip=False, flip_x=False, flip_w=W, cull=False,
             shift_x=0, shift_y=0, cull_w=W, cull_h=H, top=0, left=0)
    if n_events > cfg.slice_max_evs:
        p["start"] = random.choice(range(n_events - cfg.slice_max_evs + 1))
        p["count"] = cfg.slice_max_evs
    if cfg.is_train:
        p["time_flip"] = bool(np.random.random() < 0.5)
        p["flip_x"] = bool(np.random.random() < 0.5)
        xs, ys = np.random.randint(-cfg.max_random_shift_evs, cfg.max_random_shift_evs + 1, size=(2,))
        p["shift_x"], p["shift_y"], p["cull"] = int(xs), int(ys), True
        # torchvision RandomCrop.get_params on the (padded) image
        ph = H + 2 * max(cfg.input_H - H, 0)
        pw = W + 2 * max(cfg.input_W - W, 0)
        if not (ph == cfg.input_H and pw == cfg.input_W):
            p["top"] = int(torch.randint(0, ph - cfg.input_H + 1, size=(1,)).item())
            p["left"] = int(torch.randint(0, pw - cfg.input_W + 1, size=(1,)).item())
        if cfg.rand_aug:
            p["randaug"] = _draw_randaug(cfg)
    return p


def _draw_randaug(cfg):
    """EventRandAugment(small=False, magnitude=20)'s draws (datasets.py:657), after the sample's other draws."""
    from . import randaug_ref
    return randaug_ref.draw_ops(lambda n: int(torch.randint(n, (1,)).item()), randaug_ref.OPS, 2, 20, 31, cfg.input_H, cfg.input_W)


def _apply_randaug(x: torch.Tensor, p: dict) -> torch.Tensor:
    """ToUnit8 -> EventRandAugment -> ToFloat32 (datasets.py:655-658) on the chain's float32 [3, h, w] output."""
    from . import randaug_ref
    u8 = randaug_ref.rand_augment(randaug_ref.to_uint8(x.numpy()), p["randaug"])
    return torch.from_numpy(randaug_ref.to_float32(u8))


def apply_event_aug(events: np.ndarray, p: dict) -> np.ndarray:
    x = np.array(events, dtype=np.float64, copy=True)
    x[:, 0] *= p["scale_x"]
    x[:, 1] *= p["scale_y"]
    x = x[p["start"]:p["start"] + p["count"], :]
    if p["time_flip"]:
        x = np.flip(x, axis=0)
        x[:, 2] = x[0, 2] - x[:, 2]
        x[:, 3] = -x[:, 3]
    if p["flip_x"]:
        x[:, 0] = p["flip_w"] - 1 - x[:, 0]
    if p["cull"]:
        x[:, 0] += np.int64(p["shift_x"])
        x[:, 1] += np.int64(p["shift_y"])
        valid = (x[:, 0] >= 0) & (x[:, 0] < p["cull_w"]) & (x[:, 1] >= 0) & (x[:, 1] < p["cull_h"])
        x = x[valid]
    return x


def apply_post_raster(hist: np.ndarray, p: dict, cfg: PipelineCfg) -> torch.Tensor:
    """uint8 (H,W,3) -> float32 (3,outH,outW), torch ops as the reference executes them."""
    x = torch.from_numpy(np.ascontiguousarray(hist)).permute(2, 0, 1).contiguous().to(torch.float32).div(255)
    if cfg.is_train:
        _, h, w = x.shape
        if w < cfg.input_W:
            x = torch.nn.functional.pad(x, (cfg.input_W - w, cfg.input_W - w, 0, 0))
        if h < cfg.input_H:
            x = torch.nn.functional.pad(x, (0, 0, cfg.input_H - h, cfg.input_H - h))
        x = x[:, p["top"]:p["top"] + cfg.input_H, p["left"]:p["left"] + cfg.input_W].clone()
    if not cfg.timesurface:
        x[1, :, :] = 0.0
    if cfg.hotpixfilter:
        pol = x[0::2, :, :]
        thr = torch.mean(pol) + cfg.hotpix_num_stds * torch.std(pol)
        hot = torch.atleast_1d(torch.squeeze(torch.argwhere(pol.flatten() > thr)))
        idx = np.asarray(np.unravel_index(hot, x.shape)).T        # the reference unravels against x.shape (transforms.py:272)
        x[0::2, idx[:, 1], idx[:, 2]] = 0
    if cfg.logtrafo:                                            # LogTransform (transforms.py:200-210)
        ones = torch.ones(x[0:1, :, :].shape)
        x[0:1, :, :] = torch.log(x[0:1, :, :] + ones)
        x[2:3, :, :] = torch.log(x[2:3, :, :] + ones)
        x = x.float()
    if cfg.gammatrafo:                                          # GammaTransform (transforms.py:212-222)
        x[0:1, :, :] = x[0:1, :, :] ** cfg.gamma
        x[2:3, :, :] = x[2:3, :, :] ** cfg.gamma
        x = x.float()
    if cfg.normalize_events:
        if x[0::2, :, :].max() != 0:
            x[0::2, :, :] = x[0::2, :, :] * (1.0 / x[0::2, :, :].max())
    return x.float()


def pipeline_ref(events: np.ndarray, cfg: PipelineCfg, params: dict | None = None) -> torch.Tensor:
    p = params if params is not None else draw_params(len(events), cfg)
    H, W = cfg.raster_hw()
    ev = apply_event_aug(events, p)
    hist = event_hist_ref(ev, H, W, cfg.timesurface) if len(ev) else np.zeros((H, W, 3), np.uint8)
    out = apply_post_raster(hist, p, cfg)
    return _apply_randaug(out, p) if (cfg.is_train and cfg.rand_aug) else out


# ---------------------------------------------------------------------------------------------
# Variable sensor size: the N-Caltech101 / N-Cars branch of build_transformNPY (H = W = None, datasets.py:611-660)
# ---------------------------------------------------------------------------------------------
@dataclass
class VarPipelineCfg:
    """The ``args`` fields build_transformNPY reads on the branch without a fixed sensor."""
    is_train: bool = True
    input_H: int = 224
    input_W: int = 224
    slice_max_evs: int = 30000
    max_random_shift_evs: int = 15
    hotpixfilter: bool = True
    hotpix_num_stds: float = 10
    normalize_events: bool = False
    rand_aug: bool = False
    logtrafo: bool = False
    gammatrafo: bool = False
    gamma: float = 0.5
    timesurface: bool = False


def draw_params_var(n_events: int, cfg: VarPipelineCfg) -> dict:
    """Generator consumption of one sample: ``random.choice`` (window, only for long streams), two ``np.random.random``
    (time flip, x flip), ``np.random.randint(size=(2,))`` (shift).  ``RandomCrop`` draws nothing: after ``Resize`` the
    image already has the crop size (torchvision ``get_params`` returns early)."""
    p = dict(start=0, count=n_events, time_flip=False, flip_x=False, cull=False, shift_x=0, shift_y=0)
    if n_events > cfg.slice_max_evs:
        p["start"] = random.choice(range(n_events - cfg.slice_max_evs + 1))
        p["count"] = cfg.slice_max_evs
    if cfg.is_train:
        p["time_flip"] = bool(np.random.random() < 0.5)
        p["flip_x"] = bool(np.random.random() < 0.5)
        xs, ys = np.random.randint(-cfg.max_random_shift_evs, cfg.max_random_shift_evs + 1, size=(2,))
        p["shift_x"], p["shift_y"], p["cull"] = int(xs), int(ys), True
        if cfg.rand_aug:
            p["randaug"] = _draw_randaug(cfg)
    return p


def apply_event_aug_var(events: np.ndarray, p: dict) -> np.ndarray:
    """Sizes are inferred where the reference infers them: the flip width from the window (datasets.py:513-515), the cull
    window from the flipped rows before the shift (:538-541)."""
    x = np.array(events, dtype=np.float64, copy=True)
    x = x[p["start"]:p["start"] + p["count"], :]
    if p["time_flip"]:
        x = np.flip(x, axis=0)
        x[:, 2] = x[0, 2] - x[:, 2]
        x[:, 3] = -x[:, 3]
    if p["flip_x"]:
        W = x[:, 0].max().astype(np.int64) + 1
        x[:, 0] = W - 1 - x[:, 0]
    if p["cull"]:
        W = x[:, 0].max().astype(np.int64) + 1
        H = x[:, 1].max().astype(np.int64) + 1
        x[:, 0] += np.int64(p["shift_x"])
        x[:, 1] += np.int64(p["shift_y"])
        x = x[(x[:, 0] >= 0) & (x[:, 0] < W) & (x[:, 1] >= 0) & (x[:, 1] < H)]
    return x


def pipeline_var_ref(events: np.ndarray, cfg: VarPipelineCfg, params: dict | None = None) -> torch.Tensor:
    p = params if params is not None else draw_params_var(len(events), cfg)
    ev = apply_event_aug_var(events, p)
    hist = event_hist_ref(ev, None, None, cfg.timesurface)      # raises ValueError on an empty stream, like the reference
    x = torch.from_numpy(np.ascontiguousarray(hist)).permute(2, 0, 1).contiguous().to(torch.float32).div(255)
    # transforms.Resize((H, W), BILINEAR, antialias=True) on a tensor image
    x = torch.nn.functional.interpolate(x[None], size=(cfg.input_H, cfg.input_W), mode="bilinear", align_corners=False,
                                        antialias=True)[0]
    if not cfg.timesurface:                                     # RemoveTimesurface, datasets.py:644-645
        x[1, :, :] = 0.0
    if cfg.hotpixfilter:
        pol = x[0::2, :, :]
        thr = torch.mean(pol) + cfg.hotpix_num_stds * torch.std(pol)
        hot = torch.atleast_1d(torch.squeeze(torch.argwhere(pol.flatten() > thr)))
        idx = np.asarray(np.unravel_index(hot, x.shape)).T
        x[0::2, idx[:, 1], idx[:, 2]] = 0
    if cfg.logtrafo:                                            # LogTransform, transforms.py:200-210 (float32 here: after Resize)
        x[0::2, :, :] = torch.log(x[0::2, :, :] + 1.0)
    if cfg.gammatrafo:                                          # GammaTransform, transforms.py:212-222
        x[0::2, :, :] = x[0::2, :, :] ** cfg.gamma
    if cfg.normalize_events:
        if x[0::2, :, :].max() != 0:
            x[0::2, :, :] = x[0::2, :, :] * (1.0 / x[0::2, :, :].max())
    x = x.float()
    return _apply_randaug(x, p) if (cfg.is_train and cfg.rand_aug) else x
