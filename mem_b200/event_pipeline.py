"""Events -> model input for a whole batch on the GPU: event-space augmentations fused into the rasteriser,
then the post-raster tensor transforms.

Drop-in for what ``build_transformNPY`` composes per sample on the CPU (reference ``mem/datasets.py:611-660``,
fixed-sensor N-ImageNet path): ``ReshapeScaleXandY`` -> ``SliceRandomMaxEvs`` -> ``RandomTimeFlip`` ->
``Aug_FlipEvsAlongX`` -> ``Aug_RandomShiftEvs`` -> ``EventArrToImg`` -> ``ToTensor`` -> ``RandomCrop`` ->
``RemoveTimesurface`` -> ``RemoveHotPixels`` -> ``NormalizeEvent``.

Division of labour: the random draws stay on the host and consume Python's ``random``, ``numpy.random`` and
``torch``'s global generators exactly like the reference chain does, sample by sample (``draw_params``; same idea as
``MaskingGenerator``); the arithmetic runs on the device for the whole batch: ``memb_hist_aug_u8`` (one pass over the
raw float64 rows, bit-exact counts) and ``memb_raster_post_f32`` (uint8 counts -> float32 [B,C,h,w]).

The branch WITHOUT a fixed sensor (``H = W = None``: N-Caltech101 / N-Cars, sizes inferred per sample, bilinear
anti-aliased ``Resize`` to the input size) is ``VarPipelineConfig`` / ``draw_params_var`` / ``pipeline_var_fused`` /
``EventBatchPipelineVar`` below (one kernel, ``memb_event_pipeline_var_f32``).

``rand_aug`` (the reference scripts' default) appends ``ToUnit8 -> EventRandAugment(magnitude=20) -> ToFloat32`` as one more
launch over the batch (``mem_b200/transforms.py``, ``csrc/randaug.cu``), with each sample's operations drawn right after its
other draws so the torch generator is consumed in the reference's order.
``logtrafo`` / ``gammatrafo`` (``LogTransform`` / ``GammaTransform``, off in the reference's scripts) are a 256-entry value
table evaluated on the host with the reference's own CPU routines (``value_table``) and applied inside the fused kernel
or the post-raster pass; on the variable-sensor branch they act on the resized float32 planes (evaluated on the device).
``timesurface`` (off in the reference's scripts) takes the two-kernel path on the fixed-sensor branch:
``memb_hist_aug_tss_u8`` normalises the timestamps over the rows that survive the augmentations and honours
RandomTimeFlip's reversed order, then ``post_raster``; the variable-sensor kernel makes a second pass over its window.
"""
from __future__ import annotations

import ctypes
import random
from dataclasses import dataclass

import numpy as np

from . import _lib

__all__ = ["PipelineConfig", "EventAug", "draw_params", "pack_params", "rasterise_augmented", "post_raster",
           "pipeline_fused", "EventBatchPipeline", "VarPipelineConfig", "draw_params_var", "pipeline_var_fused",
           "EventBatchPipelineVar"]

FUSED_MAX_PIXELS = 50 * 1024      # one shared-memory tile (csrc/hist.cu kTileMaxWords)


class EventAug(ctypes.Structure):
    """``memb_event_aug`` (include/memb.h), 64 bytes."""
    _fields_ = [("scale_x", ctypes.c_double), ("scale_y", ctypes.c_double), ("start", ctypes.c_int64),
                ("count", ctypes.c_int64), ("time_flip", ctypes.c_int32), ("flip_x", ctypes.c_int32),
                ("flip_w", ctypes.c_int32), ("cull", ctypes.c_int32), ("shift_x", ctypes.c_int32),
                ("shift_y", ctypes.c_int32), ("cull_w", ctypes.c_int32), ("cull_h", ctypes.c_int32)]


AUG_DTYPE = np.dtype([("scale_x", "<f8"), ("scale_y", "<f8"), ("start", "<i8"), ("count", "<i8"), ("time_flip", "<i4"),
                      ("flip_x", "<i4"), ("flip_w", "<i4"), ("cull", "<i4"), ("shift_x", "<i4"), ("shift_y", "<i4"),
                      ("cull_w", "<i4"), ("cull_h", "<i4")])
assert AUG_DTYPE.itemsize == ctypes.sizeof(EventAug) == 64


@dataclass
class PipelineConfig:
    """The ``args`` fields ``build_transformNPY`` reads (run_mem_pretraining.py:48-57,80-81), fixed sensor."""
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
    rand_aug: bool = False          # args.rand_aug (the reference's scripts default to 1): EventRandAugment(magnitude=20) when training
    logtrafo: bool = False          # LogTransform / GammaTransform (off in the reference's scripts)
    gammatrafo: bool = False
    gamma: float = 0.5

    def __post_init__(self):
        # the reference's own range checks (datasets.py:466-469, 491, 530)
        assert 100 <= self.input_H <= 640 and 100 <= self.input_W <= 640
        assert 100 <= self.sensor_H <= 640 and 100 <= self.sensor_W <= 640
        assert 5000 <= self.slice_max_evs < 200000
        assert 0 <= self.max_random_shift_evs <= 200
        if self.is_train:
            hw = [self.sensor_H, self.sensor_W]
            s = 256 / hw[int(np.argmin(hw))]
            scale = 256 / 480
            self._scales, self._raster = (s, s), (int(480 * scale), int(640 * scale))
        else:
            self._scales = (self.input_W / self.sensor_W, self.input_H / self.sensor_H)
            self._raster = (self.input_H, self.input_W)

    def scales(self):
        """(scale_x, scale_y) of ReshapeScaleXandY (datasets.py:472-480)."""
        return self._scales

    def raster_hw(self):
        """(H, W) the rasteriser runs at (datasets.py:616-621)."""
        return self._raster


_RANDAUG = None


def _draw_randaug(cfg) -> list:
    """The sample's EventRandAugment operations (mem/datasets.py:655-658: ``EventRandAugment(small=False, magnitude=20)``),
    drawn from the global torch generator right after the sample's other draws -- the order the reference's per-sample
    transform chain consumes it in."""
    global _RANDAUG
    if _RANDAUG is None:
        import contextlib
        import io
        from .transforms import EventRandAugment
        with contextlib.redirect_stdout(io.StringIO()):
            _RANDAUG = EventRandAugment(small=False, magnitude=20)
    return _RANDAUG.draw(cfg.input_H, cfg.input_W)


def _apply_randaug(out, params, cfg, channels):
    """ToUnit8 -> EventRandAugment -> ToFloat32 on the pipeline's float32 batch, one launch (mem/datasets.py:655-658)."""
    if not (cfg.is_train and cfg.rand_aug):
        return out
    if channels != 3:
        raise ValueError("EventRandAugment needs the 3-channel [pos, 0, neg] image (torchvision's colour operations)")
    from .transforms import OP_DTYPE, apply_ops, encode_op
    ops = np.zeros((len(params), max(len(p["randaug"]) for p in params)), dtype=OP_DTYPE)
    for b, p in enumerate(params):
        for k, (name, mag) in enumerate(p["randaug"]):
            ops[b, k] = encode_op(name, mag)
    return apply_ops(out, ops, out_float=True)


def draw_params(n_events: int, cfg: PipelineConfig) -> dict:
    """One sample's random draws, consuming the global generators in the reference's order:
    ``random.choice`` (slice start, only when the stream is longer than ``slice_max_evs``, datasets.py:495-496),
    ``np.random.random`` (time flip, :603), ``np.random.random`` (x flip, :518), ``np.random.randint(size=(2,))``
    (shift, :541), ``torch.randint`` twice (torchvision ``RandomCrop.get_params``: top, then left)."""
    import torch
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
        ph = H + 2 * max(cfg.input_H - H, 0)        # RandomCrop pads both sides when the image is smaller
        pw = W + 2 * max(cfg.input_W - W, 0)
        if not (ph == cfg.input_H and pw == cfg.input_W):
            p["top"] = int(torch.randint(0, ph - cfg.input_H + 1, size=(1,)).item())
            p["left"] = int(torch.randint(0, pw - cfg.input_W + 1, size=(1,)).item())
        if cfg.rand_aug:
            p["randaug"] = _draw_randaug(cfg)
    return p


def pack_params(params) -> tuple[np.ndarray, np.ndarray]:
    """List of ``draw_params`` dicts -> (``memb_event_aug`` records [B], crop (top, left) int32 [B,2])."""
    aug = np.zeros(len(params), dtype=AUG_DTYPE)
    crop = np.zeros((len(params), 2), dtype=np.int32)
    for i, p in enumerate(params):
        aug[i] = (p["scale_x"], p["scale_y"], p["start"], p["count"], int(p["time_flip"]), int(p["flip_x"]), p["flip_w"],
                  int(p["cull"]), p["shift_x"], p["shift_y"], p["cull_w"], p["cull_h"])
        crop[i] = (p["top"], p["left"])
    return aug, crop


def rasterise_augmented(events, offsets, aug, H, W, channels=3, *, max_stream_len=None, strategy=_lib.HIST_AUTO,
                        check=True, out=None, timesurface=False):
    """Ragged batch of raw streams + per-stream ``memb_event_aug`` records -> ``uint8 (B,H,W,channels)`` on the device.

    events ``float64 (sum N_b, 4)`` CUDA tensor (or numpy / CPU tensor: copied), offsets ``int64 (B+1,)``,
    aug: numpy structured array (``AUG_DTYPE``) or a ``uint8 (B*64,)`` / ``(B,64)`` CUDA tensor holding the records."""
    torch = _lib.require_cuda()
    from .process_data import _as_device_events
    device = torch.device(events.device if (isinstance(events, torch.Tensor) and events.is_cuda) else "cuda")
    with torch.cuda.device(device):
        ev, _ = _as_device_events(torch, events, device)
        off = offsets if isinstance(offsets, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(offsets, dtype=np.int64))
        off = off.to(device=device, dtype=torch.int64).contiguous()
        B = int(off.numel()) - 1
        if B < 1:
            raise ValueError("offsets must have B+1 >= 2 entries")
        aug_dev = _aug_to_device(torch, aug, B, device)
        if out is None:
            out = torch.empty((B, H, W, channels), dtype=torch.uint8, device=device)
        n = int(ev.shape[0])
        lib = _lib.load()
        if timesurface and channels != 3:
            raise ValueError("the time surface is the middle one of three channels")
        need = lib.memb_hist_workspace_bytes(B, n, H, W, int(bool(timesurface)), _lib.HIST_GLOBAL if timesurface else strategy)
        ws = _lib.workspace.get(torch, need, device, "hist")
        stream = _lib.stream_ptr(torch, device)
        if timesurface:     # EventArrToImg(timeSurface=True) after the augmentations: L2-RED strategy + last-writer pass
            _lib.check(lib.memb_hist_aug_tss_u8(ev.data_ptr() if n else None, n, off.data_ptr(), B,
                                                int(max_stream_len if max_stream_len is not None else n), aug_dev.data_ptr(),
                                                H, W, 1, out.data_ptr(), ws.data_ptr(), ws.numel(), stream))
        else:
            _lib.check(lib.memb_hist_aug_u8(ev.data_ptr() if n else None, n, off.data_ptr(), B,
                                            int(max_stream_len if max_stream_len is not None else n), aug_dev.data_ptr(),
                                            H, W, channels, strategy, out.data_ptr(), ws.data_ptr(), ws.numel(), stream))
        if check:
            _lib.check(lib.memb_hist_status(ws.data_ptr(), stream))
    return out


def _aug_to_device(torch, aug, B, device):
    if isinstance(aug, np.ndarray):
        if aug.dtype != AUG_DTYPE or aug.shape != (B,):
            raise ValueError(f"aug must be a ({B},) array of memb_event_aug records")
        return torch.from_numpy(aug.view(np.uint8).reshape(B, 64)).to(device)
    aug_dev = aug.to(device).contiguous()
    if aug_dev.dtype != torch.uint8 or aug_dev.numel() != B * 64:
        raise ValueError("aug tensor must hold B 64-byte memb_event_aug records (uint8)")
    return aug_dev


def value_table(logtrafo=False, gammatrafo=False, gamma=0.5):
    """The 256 values an image can hold after ``LogTransform`` / ``GammaTransform`` (mem/transforms.py:200-222; applied in
    this order between RemoveHotPixels and NormalizeEvent, mem/datasets.py:647-652), evaluated on ``c / 255`` with the
    reference's own CPU routines (``torch.log``, ``Tensor.__pow__``) so that the device path reproduces them bit for bit.
    None when neither transform is on."""
    if not (logtrafo or gammatrafo):
        return None
    import torch
    x = torch.arange(256, dtype=torch.uint8).to(torch.float32).div(255)        # ToTensor
    if logtrafo:
        x = torch.log(x + torch.ones(x.shape))
    if gammatrafo:
        x = x ** gamma
    return x.float().contiguous()


def pipeline_fused(events, offsets, aug, crop_tl, H, W, out_hw, channels=3, *, hot_num_stds=10.0, normalize=False,
                   check=True, out=None, value_lut=None):
    """The whole chain in one kernel (``memb_event_pipeline_f32``): raw ragged batch -> ``float32 (B,C,outH,outW)``.
    Needs ``outH * outW <= FUSED_MAX_PIXELS``; arguments as ``rasterise_augmented`` + ``post_raster``.
    Inputs that already are contiguous CUDA tensors of the right dtype are used as they are (no copies)."""
    torch = _lib.require_cuda()
    outH, outW = int(out_hw[0]), int(out_hw[1])
    ready = (isinstance(events, torch.Tensor) and events.is_cuda and events.dtype == torch.float64 and events.is_contiguous()
             and events.ndim == 2 and events.shape[1] == 4)
    device = events.device if ready else torch.device(
        events.device if (isinstance(events, torch.Tensor) and events.is_cuda) else "cuda")
    if device.index is not None and device.index != torch.cuda.current_device():
        with torch.cuda.device(device):
            return pipeline_fused(events, offsets, aug, crop_tl, H, W, out_hw, channels, hot_num_stds=hot_num_stds,
                                  normalize=normalize, check=check, out=out, value_lut=value_lut)
    if ready:
        ev = events
    else:
        from .process_data import _as_device_events
        ev, _ = _as_device_events(torch, events, device)
    if isinstance(offsets, torch.Tensor) and offsets.is_cuda and offsets.dtype == torch.int64 and offsets.is_contiguous():
        off = offsets
    else:
        off = offsets if isinstance(offsets, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(offsets, dtype=np.int64))
        off = off.to(device=device, dtype=torch.int64).contiguous()
    B = off.numel() - 1
    if B < 1:
        raise ValueError("offsets must have B+1 >= 2 entries")
    if isinstance(aug, torch.Tensor) and aug.is_cuda and aug.dtype == torch.uint8 and aug.is_contiguous() and aug.numel() == B * 64:
        aug_dev = aug
    else:
        aug_dev = _aug_to_device(torch, aug, B, device)
    crop = None
    if crop_tl is not None:
        if isinstance(crop_tl, torch.Tensor) and crop_tl.is_cuda and crop_tl.dtype == torch.int32 and crop_tl.is_contiguous():
            crop = crop_tl
        else:
            crop = crop_tl if isinstance(crop_tl, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(crop_tl, dtype=np.int32))
            crop = crop.to(device=device, dtype=torch.int32).contiguous()
        if tuple(crop.shape) != (B, 2):
            raise ValueError(f"crop_tl must be ({B}, 2)")
    if out is None:
        out = torch.empty((B, channels, outH, outW), dtype=torch.float32, device=device)
    elif not (out.is_cuda and out.dtype == torch.float32 and out.is_contiguous() and tuple(out.shape) == (B, channels, outH, outW)):
        raise ValueError("out must be a contiguous float32 CUDA tensor (B, channels, outH, outW)")
    lib = _lib.load()
    ws = _lib.workspace.get(torch, 256, device, "hist")
    stream = _lib.stream_ptr(torch, device)
    n = ev.shape[0]
    lut = None
    if value_lut is not None:      # LogTransform / GammaTransform: float32 [256] table of the values written per count
        lut = value_lut.to(device=device, dtype=torch.float32).contiguous()
        if lut.numel() != 256:
            raise ValueError("value_lut must hold 256 float32 values")
    _lib.check(lib.memb_event_pipeline_lut_f32(
        ev.data_ptr() if n else None, n, off.data_ptr(), B, aug_dev.data_ptr(),
        crop.data_ptr() if crop is not None else None, H, W, max(outH - H, 0), max(outW - W, 0), outH, outW, channels,
        float(hot_num_stds) if hot_num_stds is not None else -1.0, int(bool(normalize)),
        lut.data_ptr() if lut is not None else None, out.data_ptr(), ws.data_ptr(), ws.numel(), stream))
    if check:
        _lib.check(lib.memb_hist_status(ws.data_ptr(), stream))
    return out


def post_raster(hist, crop_tl=None, out_hw=None, *, remove_timesurface=True, hot_num_stds=10.0, normalize=False,
                out=None, value_lut=None):
    """``uint8 (B,H,W,C)`` counts -> ``float32 (B,C,outH,outW)``: /255, crop (zero padding like
    ``RandomCrop(pad_if_needed=True)``), RemoveTimesurface, RemoveHotPixels (``hot_num_stds=None`` = off), LogTransform /
    GammaTransform as ``value_lut`` (``value_table(...)``, None = off), NormalizeEvent.

    crop_tl: ``int32 (B,2)`` (top, left) in the padded image (numpy or tensor) or None for (0,0)."""
    torch = _lib.require_cuda()
    if not (isinstance(hist, torch.Tensor) and hist.is_cuda and hist.dtype == torch.uint8 and hist.ndim == 4):
        raise ValueError("hist must be a uint8 CUDA tensor (B,H,W,C)")
    hist = hist.contiguous()
    B, H, W, C = (int(v) for v in hist.shape)
    outH, outW = (H, W) if out_hw is None else (int(out_hw[0]), int(out_hw[1]))
    pad_t, pad_l = max(outH - H, 0), max(outW - W, 0)
    device = hist.device
    with torch.cuda.device(device):
        crop = None
        if crop_tl is not None:
            crop = crop_tl if isinstance(crop_tl, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(crop_tl, dtype=np.int32))
            crop = crop.to(device=device, dtype=torch.int32).contiguous()
            if tuple(crop.shape) != (B, 2):
                raise ValueError(f"crop_tl must be ({B}, 2)")
        if out is None:
            out = torch.empty((B, C, outH, outW), dtype=torch.float32, device=device)
        lib = _lib.load()
        ws = _lib.workspace.get(torch, max(lib.memb_raster_post_workspace_bytes(B), 16), device, "raster_post")
        lut = None
        if value_lut is not None:
            lut = torch.as_tensor(value_lut, dtype=torch.float32).to(device).contiguous()
            if lut.numel() != 256:
                raise ValueError("value_lut must hold 256 float32 values")
        _lib.check(lib.memb_raster_post_lut_f32(hist.data_ptr(), B, H, W, C, crop.data_ptr() if crop is not None else None,
                                                pad_t, pad_l, outH, outW, int(bool(remove_timesurface)),
                                                float(hot_num_stds) if hot_num_stds is not None else -1.0, int(bool(normalize)),
                                                lut.data_ptr() if lut is not None else None,
                                                out.data_ptr(), ws.data_ptr(), ws.numel(), _lib.stream_ptr(torch, device)))
    return out


class EventBatchPipeline:
    """Batched GPU replacement of ``build_transformNPY(is_train, args)`` for fixed-sensor data.

    ``pipe(streams)`` with ``streams`` a list of ``(N_b, 4)`` float64 arrays (or ``(events, offsets)`` already
    concatenated) returns ``float32 (B, C, input_H, input_W)`` on the device, equal to stacking the reference
    transform's outputs when the generators start from the same state and samples are drawn in order."""

    def __init__(self, cfg: PipelineConfig, channels: int = 3, device="cuda", fused=None):
        self.cfg, self.channels, self.device = cfg, channels, device
        fits = cfg.input_H * cfg.input_W <= FUSED_MAX_PIXELS
        if fused and not fits:
            raise ValueError("the output raster does not fit one shared-memory tile; use fused=False")
        if cfg.timesurface:
            # args.timesurface: the one-kernel path rasterises polarity counts only; the time surface takes the
            # rasterise (L2 REDs + last-writer pass) + post-raster pair
            if fused:
                raise NotImplementedError("the fused kernel has no time surface; use fused=False (the default with timesurface)")
            if channels != 3:
                raise NotImplementedError("the time surface is the middle plane of the 3-channel image")
            fused = False
        self.fused = fits if fused is None else bool(fused)   # one kernel when the crop fits a shared-memory tile

    def __call__(self, streams, offsets=None, params=None):
        torch = _lib.require_cuda()
        if offsets is None:
            lens = [len(s) for s in streams]
            offsets = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
            events = np.concatenate([np.asarray(s, dtype=np.float64).reshape(-1, 4) for s in streams], axis=0) \
                if sum(lens) else np.zeros((0, 4))
        else:
            events = streams
            off_host = offsets.cpu().numpy() if isinstance(offsets, torch.Tensor) else np.asarray(offsets)
            lens = np.diff(off_host).tolist()
        cfg = self.cfg
        if params is None:
            params = [draw_params(int(n), cfg) for n in lens]
        aug, crop = pack_params(params)
        H, W = cfg.raster_hw()
        hot = cfg.hotpix_num_stds if cfg.hotpixfilter else None
        lut = value_table(cfg.logtrafo, cfg.gammatrafo, cfg.gamma)
        if self.fused:
            out = pipeline_fused(events, offsets, aug, crop if cfg.is_train else None, H, W,
                                 (cfg.input_H, cfg.input_W) if cfg.is_train else (H, W), self.channels,
                                 hot_num_stds=hot, normalize=cfg.normalize_events, check=not cfg.is_train, value_lut=lut)
            return _apply_randaug(out, params, cfg, self.channels)
        hist = rasterise_augmented(events, offsets, aug, H, W, self.channels,
                                   max_stream_len=int(max(p["count"] for p in params)) if params else 0,
                                   check=not cfg.is_train,   # after the cull every row is inside the sensor
                                   timesurface=cfg.timesurface)
        out = post_raster(hist, crop if cfg.is_train else None, (cfg.input_H, cfg.input_W) if cfg.is_train else None,
                          remove_timesurface=not cfg.timesurface,
                          hot_num_stds=cfg.hotpix_num_stds if cfg.hotpixfilter else None, normalize=cfg.normalize_events,
                          value_lut=lut)
        return _apply_randaug(out, params, cfg, self.channels)


# ------------------------------------------------------------------------- variable sensor size (N-Caltech101 / N-Cars)
@dataclass
class VarPipelineConfig:
    """The ``args`` fields ``build_transformNPY`` reads on its branch without a fixed sensor (datasets.py:614, :623-642);
    ``canvas_H`` / ``canvas_W`` bound the recordings' extent (the sensor: 180 x 240 for N-Caltech101, 100 x 120 for N-Cars)."""
    is_train: bool = True
    canvas_H: int = 180
    canvas_W: int = 240
    input_H: int = 224
    input_W: int = 224
    slice_max_evs: int = 30000
    max_random_shift_evs: int = 15
    timesurface: bool = False
    hotpixfilter: bool = True
    hotpix_num_stds: float = 10
    normalize_events: bool = False
    rand_aug: bool = False
    logtrafo: bool = False
    gammatrafo: bool = False
    gamma: float = 0.5

    def __post_init__(self):
        assert 5000 <= self.slice_max_evs < 200000 and 0 <= self.max_random_shift_evs <= 200     # datasets.py:491, :530
        assert self.canvas_H * self.canvas_W <= FUSED_MAX_PIXELS, "the canvas must fit one shared-memory tile"


def draw_params_var(n_events: int, cfg: VarPipelineConfig) -> dict:
    """One sample's draws in the reference's order: ``random.choice`` (window start, long streams only, datasets.py:495),
    ``np.random.random`` (time flip, :603), ``np.random.random`` (x flip, :518), ``np.random.randint(size=(2,))`` (shift,
    :541).  ``RandomCrop`` draws nothing on this branch: ``Resize`` already produced the crop size."""
    p = dict(scale_x=1.0, scale_y=1.0, start=0, count=n_events, time_flip=False, flip_x=False, flip_w=0, cull=False,
             shift_x=0, shift_y=0, cull_w=0, cull_h=0, top=0, left=0)
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


def pipeline_var_fused(events, offsets, aug, canvas_hw, out_hw, channels=3, *, hot_num_stds=10.0, normalize=False, check=True,
                       out=None, logtrafo=False, gammatrafo=False, gamma=0.5, timesurface=False):
    """Ragged batch of raw streams -> ``float32 (B,C,outH,outW)`` through ``memb_event_pipeline_var_tf_f32`` (sizes inferred
    per stream, anti-aliased bilinear resize, optional log / gamma maps of the resized planes, optional time surface in the
    middle plane).  ``check`` synchronises and
    raises ``ValueError`` where the reference raises (a stream that is empty after the window / shift) or when a recording
    exceeds the canvas."""
    torch = _lib.require_cuda()
    from .process_data import _as_device_events
    device = torch.device(events.device if (isinstance(events, torch.Tensor) and events.is_cuda) else "cuda")
    with torch.cuda.device(device):
        ev, _ = _as_device_events(torch, events, device)
        off = offsets if isinstance(offsets, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(offsets, dtype=np.int64))
        off = off.to(device=device, dtype=torch.int64).contiguous()
        B = int(off.numel()) - 1
        if B < 1:
            raise ValueError("offsets must have B+1 >= 2 entries")
        aug_dev = _aug_to_device(torch, aug, B, device)
        outH, outW = int(out_hw[0]), int(out_hw[1])
        if out is None:
            out = torch.empty((B, channels, outH, outW), dtype=torch.float32, device=device)
        lib = _lib.load()
        ws = _lib.workspace.get(torch, 256, device, "hist")
        stream = _lib.stream_ptr(torch, device)
        n = int(ev.shape[0])
        _lib.check(lib.memb_event_pipeline_var_tf_f32(ev.data_ptr() if n else None, n, off.data_ptr(), B, aug_dev.data_ptr(),
                                                      int(canvas_hw[0]), int(canvas_hw[1]), outH, outW, channels,
                                                      float(hot_num_stds) if hot_num_stds is not None else -1.0,
                                                      int(bool(normalize)), int(bool(logtrafo)), int(bool(gammatrafo)), float(gamma),
                                                      int(bool(timesurface)), out.data_ptr(), ws.data_ptr(), ws.numel(), stream))
        if check:
            _lib.check(lib.memb_hist_status(ws.data_ptr(), stream))
    return out


class EventBatchPipelineVar:
    """Batched GPU replacement of ``build_transformNPY(is_train, args)`` for data without a fixed sensor size
    (N-Caltech101, N-Cars): ``pipe(streams)`` -> ``float32 (B, C, input_H, input_W)`` on the device, equal (to float32
    rounding of the resize filter) to stacking the reference transform's outputs under the same generator state."""

    def __init__(self, cfg: VarPipelineConfig, channels: int = 3):
        if cfg.timesurface and channels != 3:
            raise ValueError("the time surface lives in the middle plane of the 3-channel image")
        self.cfg, self.channels = cfg, channels

    def __call__(self, streams, offsets=None, params=None, check=True):
        torch = _lib.require_cuda()
        if offsets is None:
            lens = [len(s) for s in streams]
            offsets = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
            events = np.concatenate([np.asarray(s).reshape(-1, 4) for s in streams], axis=0) if sum(lens) else np.zeros((0, 4))
        else:
            events = streams
            off_host = offsets.cpu().numpy() if isinstance(offsets, torch.Tensor) else np.asarray(offsets)
            lens = np.diff(off_host).tolist()
        cfg = self.cfg
        if params is None:
            params = [draw_params_var(int(n), cfg) for n in lens]
        aug, _ = pack_params(params)
        out = pipeline_var_fused(events, offsets, aug, (cfg.canvas_H, cfg.canvas_W), (cfg.input_H, cfg.input_W), self.channels,
                                 hot_num_stds=cfg.hotpix_num_stds if cfg.hotpixfilter else None,
                                 normalize=cfg.normalize_events, check=check, logtrafo=cfg.logtrafo, gammatrafo=cfg.gammatrafo,
                                 gamma=cfg.gamma, timesurface=cfg.timesurface)
        return _apply_randaug(out, params, cfg, self.channels)
