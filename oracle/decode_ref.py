"""Oracle (test infrastructure, not product): raw recording decoders.

Restates ``process_data/process_dataset.py``: ``ncaltech101`` (:24-63) and ``ncars`` (:66-103).  ``*_loop`` follow the
reference statement by statement (byte-by-byte Python, for small inputs); ``*_np`` are the vectorised forms used at
larger sizes and are checked against the loops.  Pinned against the reference's own functions run on synthetic
recordings by ``tests/golden/decode.npz`` (``oracle/make_golden.py::golden_decode``).
"""
from __future__ import annotations

import struct

import numpy as np


def ncaltech101_loop(raw: bytes) -> np.ndarray:
    events = []
    for i in range(0, len(raw), 5):
        data = raw[i:i + 5]
        y = data[0]
        x = data[1]
        p = (data[2] >> 7) & 0x01
        t = (data[2] & 0x7f).to_bytes(1, byteorder="big") + data[3:5]
        t = int.from_bytes(t, byteorder="big")
        p = 2. * p - 1.
        events.append([float(y), float(x), float(t), float(p)])
    return np.array(events).astype(float).reshape(-1, 4)


def ncars_loop(raw: bytes) -> np.ndarray:
    events = []
    for i in range(0, len(raw), 8):
        t = struct.unpack("I", raw[i:i + 4])[0]
        data = int.from_bytes(raw[i + 4:i + 8], byteorder="little")
        y = (data & 0x00003fff)
        x = (data & 0x0fffc000) >> 14
        p = (data & 0x10000000) >> 28
        events.append([y, x, t, bool(p)])
    return np.array(events).astype(float).reshape(-1, 4)


def ncaltech101_np(raw) -> np.ndarray:
    b = np.frombuffer(bytes(raw), dtype=np.uint8).reshape(-1, 5).astype(np.int64)
    t = ((b[:, 2] & 0x7f) << 16) | (b[:, 3] << 8) | b[:, 4]
    p = 2.0 * ((b[:, 2] >> 7) & 1) - 1.0
    return np.stack([b[:, 0].astype(np.float64), b[:, 1].astype(np.float64), t.astype(np.float64), p], axis=1)


def ncars_np(raw) -> np.ndarray:
    w = np.frombuffer(bytes(raw), dtype="<u4").reshape(-1, 2).astype(np.int64)
    d = w[:, 1]
    return np.stack([(d & 0x3fff).astype(np.float64), ((d & 0x0fffc000) >> 14).astype(np.float64),
                     w[:, 0].astype(np.float64), ((d & 0x10000000) >> 28).astype(np.float64)], axis=1)


def skip_dat_header(blob: bytes) -> bytes:
    """Payload of a Prophesee .dat file: '%' header lines, then 2 bytes, are skipped (process_dataset.py:77-85)."""
    pos = 0
    while pos < len(blob) and blob[pos] == 37:
        nl = blob.find(b"\n", pos, pos + 256)
        pos = (nl + 1) if nl >= 0 else min(pos + 256, len(blob))
    return blob[pos + 2:]


def synth_ncaltech101(rng, n, W=240, H=180) -> bytes:
    b = np.zeros((n, 5), dtype=np.uint8)
    b[:, 0] = rng.integers(0, W, n)
    b[:, 1] = rng.integers(0, H, n)
    t = np.sort(rng.integers(0, 1 << 23, n))
    b[:, 2] = ((t >> 16) & 0x7f) | (rng.integers(0, 2, n) << 7)
    b[:, 3] = (t >> 8) & 0xff
    b[:, 4] = t & 0xff
    return b.tobytes()


def synth_ncars(rng, n, W=120, H=100, header=True) -> bytes:
    t = np.sort(rng.integers(0, 1 << 32, n, dtype=np.uint64)).astype("<u4")
    d = (rng.integers(0, W, n) | (rng.integers(0, H, n) << 14) | (rng.integers(0, 2, n) << 28)
         | (rng.integers(0, 8, n) << 29)).astype("<u4")          # top bits are don't-care in the reference's masks
    payload = np.stack([t, d], axis=1).tobytes()
    if not header:
        return payload
    return b"% Data file containing CD events.\n% Version 2\n% Date 2017-01-01 00:00:00\n% end\n" + bytes([0, 8]) + payload
