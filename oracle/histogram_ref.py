"""Oracle (test infrastructure, not product): event stream -> polarity-count image.

numpy restatement of ``EventArrToImg.__call__`` (reference
``mem/datasets.py:566-595``; duplicate at
``mem/semantic_segmentation/backbone/EventDataset.py:767-796``).

Semantics that the CUDA path must reproduce bit for bit:

* rows are ``[x, y, t, p]`` float64; x and y are truncated toward zero
  (``astype(int)``, datasets.py:568-569);
* when H / W are not given they are ``max(trunc(coord)) + 1`` (datasets.py:571-575);
* only rows with ``p == +1`` feed the positive image and rows with ``p == -1``
  the negative one (exact float compare, datasets.py:581-582); anything else
  (e.g. N-Cars' 0/1 polarity) is dropped;
* the flat pixel index is ``x + W*y``; numpy's unbuffered ``add.at`` wraps a
  negative index once (``idx + H*W``) and raises ``IndexError`` outside
  ``[-H*W, H*W)``;
* counters are uint8 and therefore wrap modulo 256;
* optional time surface: every row (any polarity) writes
  ``uint8((t - t.min()) / (t - t.min()).max() * 255)`` at its pixel, the last
  row in stream order winning (datasets.py:587-589);
* the result is ``(H, W, 3)`` = ``[pos, time-surface-or-zero, neg]``.
"""
from __future__ import annotations

import numpy as np


def event_hist_ref(events: np.ndarray, H=None, W=None, timesurface: bool = False) -> np.ndarray:
    ev = np.asarray(events)
    col_x = ev[:, 0].astype(np.int64)
    col_y = ev[:, 1].astype(np.int64)
    col_t = ev[:, 2]
    col_p = ev[:, 3]
    if W is None:
        W = int(col_x.max()) + 1
    if H is None:
        H = int(col_y.max()) + 1
    npix = H * W

    flat = col_x + W * col_y
    planes = np.zeros((3, npix), dtype=np.uint8)          # pos, tss, neg
    for plane, sign in ((0, 1), (2, -1)):
        sel = flat[col_p == sign]
        # np.add.at is the unbuffered scatter the reference uses; it is what
        # defines the wrap / IndexError behaviour, so the oracle keeps it.
        np.add.at(planes[plane], sel, 1)

    if timesurface:
        rel = col_t - col_t.min()
        with np.errstate(invalid="ignore", divide="ignore"):
            val = rel / rel.max() * 255
        planes[1][flat] = val                               # last write wins
    return np.ascontiguousarray(planes.reshape(3, H, W).transpose(1, 2, 0))


def event_hist_batched_ref(events: np.ndarray, offsets: np.ndarray, H: int, W: int,
                           channels: int = 3, timesurface: bool = False) -> np.ndarray:
    """Ragged batch: stream b is ``events[offsets[b]:offsets[b+1]]``.

    ``channels == 2`` keeps planes 0 and 2 (the ``images[:, 0::2]`` map noted
    at reference ``mem/engine_for_finetuning.py:228``).
    """
    B = len(offsets) - 1
    out = np.zeros((B, H, W, channels), dtype=np.uint8)
    for b in range(B):
        img = event_hist_ref(events[offsets[b]:offsets[b + 1]], H, W, timesurface) \
            if offsets[b + 1] > offsets[b] else np.zeros((H, W, 3), np.uint8)
        out[b] = img if channels == 3 else img[..., 0::2]
    return out
