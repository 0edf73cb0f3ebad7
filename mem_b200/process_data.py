"""Event stream -> polarity-count histogram on the GPU.

``histogram`` is the name ``BASELINE.json``'s north star uses; the reference
keeps the rasteriser in ``EventArrToImg.__call__`` (``mem/datasets.py:566-595``,
see SURVEY.md D1) and ``mem_b200.datasets.EventArrToImg`` wraps this function
with that class's constructor and call signature.

Results are bit-identical to the reference: truncation of x / y toward zero,
``x + W*y`` flat index with numpy's single negative wrap, ``IndexError`` outside
``[-H*W, H*W)``, exact ``p == +1`` / ``p == -1`` selection, uint8 wrap modulo
256, and the optional last-writer time surface.
"""
from __future__ import annotations

import numpy as np

from . import _lib

__all__ = ["histogram", "histogram_batch", "decode_events", "histogram_raw", "read_ncaltech101_bin", "read_ncars_dat",
           "read_nimagenet_npz", "convert_nimagenet", "RAW_NCALTECH101", "RAW_NCARS"]

RAW_NCALTECH101, RAW_NCARS = _lib.RAW_NCALTECH101, _lib.RAW_NCARS
_RECORD_BYTES = {RAW_NCALTECH101: 5, RAW_NCARS: 8}


def _as_device_events(torch, events, device):
    """(N,4) float64 rows on the device; returns (tensor, came_from_host)."""
    if isinstance(events, torch.Tensor):
        ev = events
        host = not ev.is_cuda
    else:
        arr = np.ascontiguousarray(np.asarray(events))
        if arr.dtype == np.bool_ or arr.dtype.kind not in "iuf":
            arr = arr.astype(np.float64)
        try:
            ev = torch.from_numpy(arr)
        except TypeError:                      # a dtype torch has no tensor type for
            ev = torch.from_numpy(arr.astype(np.float64))
        host = True
    if ev.ndim != 2 or ev.shape[1] != 4:
        raise ValueError(f"events must be (N, 4) rows [x, y, t, p], got {tuple(ev.shape)}")
    ev = ev.contiguous()
    if host:
        # recordings stored in a narrower dtype (N-ImageNet .npz arrays, int16 / float32 exports) cross PCIe as they
        # are and are widened on the device: the rasteriser reads float64 rows (the reference's arithmetic type)
        ev = ev.to(device, non_blocking=False)
    if ev.dtype != torch.float64:
        ev = ev.to(torch.float64)
    return ev, host


def _launch(torch, ev, offsets, B, max_len, H, W, channels, timesurface, strategy, out, check):
    lib = _lib.load()
    n = int(ev.shape[0])
    need = lib.memb_hist_workspace_bytes(B, n, H, W, int(timesurface), strategy)
    ws = _lib.workspace.get(torch, need, ev.device, "hist")
    stream = _lib.stream_ptr(torch, ev.device)
    _lib.check(lib.memb_hist_u8(
        ev.data_ptr() if n else None, n, offsets.data_ptr() if offsets is not None else None, B, max_len,
        H, W, channels, int(timesurface), strategy, out.data_ptr(), ws.data_ptr(), ws.numel(), stream))
    if check:
        _lib.check(lib.memb_hist_status(ws.data_ptr(), stream))


def histogram(events, H=None, W=None, timesurface=False, channels=3, *, device=None,
              strategy=_lib.HIST_AUTO, check=True):
    """Rasterise one event stream.

    events: ``(N, 4)`` rows ``[x, y, t, p]`` (numpy array, CPU tensor or CUDA tensor).
    Returns ``uint8 (H, W, channels)``: ``[pos, time-surface|0, neg]`` for 3 channels,
    ``[pos, neg]`` for 2 (the ``[..., 0::2]`` map).  A numpy array in gives a numpy
    array out; a tensor in gives a CUDA tensor out.  ``check=False`` skips the
    (synchronising) out-of-bounds test.
    """
    torch = _lib.require_cuda()
    if channels not in (2, 3):
        raise ValueError("channels must be 2 or 3")
    if timesurface and channels != 3:
        raise ValueError("the time surface lives in channel 1 of a 3-channel image")
    device = torch.device(device if device is not None else
                          (events.device if isinstance(events, torch.Tensor) and events.is_cuda else "cuda"))
    numpy_in = not isinstance(events, torch.Tensor)
    with torch.cuda.device(device):
        ev, _ = _as_device_events(torch, events, device)
        n = int(ev.shape[0])
        if H is None or W is None:
            if n == 0:
                raise ValueError("zero-size array to reduction operation maximum which has no identity")
            lib = _lib.load()
            ws = _lib.workspace.get(torch, 256, device, "hist")
            ext = (_lib._i64 * 2)()
            _lib.check(lib.memb_hist_extent(ev.data_ptr(), n, ext, ws.data_ptr(), ws.numel(),
                                            _lib.stream_ptr(torch, device)))
            W = int(ext[0]) + 1 if W is None else W
            H = int(ext[1]) + 1 if H is None else H
            if H <= 0 or W <= 0:
                raise ValueError("negative dimensions are not allowed")
        if timesurface and n == 0:
            raise ValueError("zero-size array to reduction operation minimum which has no identity")
        out = torch.empty((H, W, channels), dtype=torch.uint8, device=device)
        _launch(torch, ev, None, 1, n, H, W, channels, timesurface, strategy, out, check)
    return out.cpu().numpy() if numpy_in else out


def histogram_batch(events, offsets, H, W, channels=3, timesurface=False, *, max_stream_len=None,
                    strategy=_lib.HIST_AUTO, check=True, out=None):
    """Rasterise a ragged batch: stream ``b`` is ``events[offsets[b]:offsets[b+1]]``.

    events ``(sum N_b, 4)`` float64 and offsets ``int64 (B+1,)`` may be numpy or tensors;
    returns ``uint8 (B, H, W, channels)`` (numpy if events was numpy, else CUDA tensor).
    """
    torch = _lib.require_cuda()
    numpy_in = not isinstance(events, torch.Tensor)
    device = torch.device(events.device if (not numpy_in and events.is_cuda) else "cuda")
    with torch.cuda.device(device):
        ev, _ = _as_device_events(torch, events, device)
        if isinstance(offsets, torch.Tensor):
            off = offsets.to(device=device, dtype=torch.int64).contiguous()
            if max_stream_len is None:
                max_stream_len = int(ev.shape[0])
        else:
            off_np = np.ascontiguousarray(np.asarray(offsets), dtype=np.int64)
            if max_stream_len is None and len(off_np) > 1:
                max_stream_len = int(np.diff(off_np).max())
            off = torch.from_numpy(off_np).to(device)
        B = int(off.numel()) - 1
        if B < 1:
            raise ValueError("offsets must have B+1 >= 2 entries")
        if out is None:
            out = torch.empty((B, H, W, channels), dtype=torch.uint8, device=device)
        _launch(torch, ev, off, B, int(max_stream_len or 0), H, W, channels, timesurface, strategy, out, check)
    return out.cpu().numpy() if numpy_in else out


# ----------------------------------------------------------------------------- raw recordings
def read_ncaltech101_bin(path) -> np.ndarray:
    """Raw bytes of an N-Caltech101 ``.bin`` recording: a sequence of 5-byte records, no header
    (reference ``process_data/process_dataset.py:47-51`` reads them with ``file.read(5)`` until EOF; a trailing
    partial record would make the reference raise ``IndexError`` -- here it raises ``ValueError``)."""
    raw = np.fromfile(path, dtype=np.uint8)
    if raw.size % 5:
        raise ValueError(f"{path}: {raw.size} bytes is not a whole number of 5-byte records")
    return raw


def read_ncars_dat(path) -> np.ndarray:
    """Raw record bytes of an N-Cars / Prophesee ``.dat`` file: skips the ``%`` header lines and the two
    type / size bytes like the reference (``process_data/process_dataset.py:77-85``), returns the 8-byte records.
    A trailing partial record makes the reference's ``struct.unpack`` raise; here it raises ``ValueError``."""
    with open(path, "rb") as f:
        while True:
            pos = f.tell()
            line = f.readline(256)
            if not line or line[0] != 37:
                f.seek(pos)
                break
        f.read(2)
        raw = np.frombuffer(f.read(), dtype=np.uint8)
    if raw.size % 8:
        raise ValueError(f"{path}: {raw.size} payload bytes is not a whole number of 8-byte records")
    return raw


def read_nimagenet_npz(path) -> np.ndarray:
    """The ``(N, 4)`` event array of an N-ImageNet recording, ``np.load(path)["event_data"]`` in its stored dtype
    (reference ``process_data/process_dataset.py:115``: the conversion is a container change, no arithmetic).  Feed it
    to ``histogram`` / ``histogram_batch`` / the event pipeline as is: it is uploaded in its own dtype and widened to
    float64 on the device."""
    with np.load(path) as z:
        if "event_data" not in z.files:
            raise KeyError(f"{path}: no 'event_data' array (N-ImageNet recordings store their events under that key)")
        data = z["event_data"]
    if data.ndim != 2 or data.shape[1] != 4:
        raise ValueError(f"{path}: event_data must be (N, 4), got {data.shape}")
    return data


def convert_nimagenet(src_npz, dst_npy=None) -> str:
    """One file of the reference's ``nimagenet()`` conversion (process_dataset.py:108-117): ``event_data`` of
    ``src_npz`` saved as ``<name>.npy`` (next to the source unless ``dst_npy`` is given).  Returns the path written."""
    import os
    if dst_npy is None:
        dst_npy = os.path.join(os.path.dirname(src_npz), os.path.basename(src_npz).split(".")[0] + ".npy")
    np.save(dst_npy, read_nimagenet_npz(src_npz))
    return dst_npy


def _raw_on_device(torch, raw, fmt, device):
    if fmt not in _RECORD_BYTES:
        raise ValueError(f"unknown raw record format {fmt}")
    t = raw if isinstance(raw, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(np.asarray(raw), dtype=np.uint8))
    if t.dtype != torch.uint8 or t.ndim != 1:
        raise ValueError("raw records must be a flat uint8 buffer")
    if t.numel() % _RECORD_BYTES[fmt]:
        raise ValueError(f"{t.numel()} bytes is not a whole number of {_RECORD_BYTES[fmt]}-byte records")
    t = t.to(device).contiguous()
    if t.data_ptr() % 16:
        t = t.clone()        # a slice of a larger buffer: the kernels want 16-byte aligned records
    return t, t.numel() // _RECORD_BYTES[fmt]


def decode_events(raw, fmt, *, device=None):
    """Raw record bytes -> ``float64 (N, 4)`` rows ``[col0, col1, t, p]``, the array the reference's
    ``process_dataset.py`` stores as ``.npy`` (N-Caltech101: p in {-1,+1}; N-Cars: p in {0,1}).
    numpy in -> numpy out; tensor in -> CUDA tensor out."""
    torch = _lib.require_cuda()
    numpy_in = not isinstance(raw, torch.Tensor)
    device = torch.device(device if device is not None else (raw.device if (not numpy_in and raw.is_cuda) else "cuda"))
    with torch.cuda.device(device):
        t, n = _raw_on_device(torch, raw, fmt, device)
        out = torch.empty((n, 4), dtype=torch.float64, device=device)
        _lib.check(_lib.load().memb_decode_events_f64(t.data_ptr() if n else None, n, fmt, out.data_ptr() if n else None,
                                                      _lib.stream_ptr(torch, device)))
    return out.cpu().numpy() if numpy_in else out


def histogram_raw(raw, fmt, H, W, channels=3, *, device=None, check=True):
    """Rasterise a recording straight from its raw records: equal to ``histogram(decode_events(raw, fmt), H, W)``
    without the float64 rows ever existing (5 or 8 bytes per event of PCIe / HBM traffic instead of 32)."""
    torch = _lib.require_cuda()
    if channels not in (2, 3):
        raise ValueError("channels must be 2 or 3")
    numpy_in = not isinstance(raw, torch.Tensor)
    device = torch.device(device if device is not None else (raw.device if (not numpy_in and raw.is_cuda) else "cuda"))
    with torch.cuda.device(device):
        t, n = _raw_on_device(torch, raw, fmt, device)
        lib = _lib.load()
        out = torch.empty((H, W, channels), dtype=torch.uint8, device=device)
        ws = _lib.workspace.get(torch, lib.memb_hist_workspace_bytes(1, n, H, W, 0, _lib.HIST_AUTO), device, "hist")
        stream = _lib.stream_ptr(torch, device)
        _lib.check(lib.memb_hist_raw_u8(t.data_ptr() if n else None, n, fmt, H, W, channels, out.data_ptr(), ws.data_ptr(),
                                        ws.numel(), stream))
        if check:
            _lib.check(lib.memb_hist_status(ws.data_ptr(), stream))
    return out.cpu().numpy() if numpy_in else out
