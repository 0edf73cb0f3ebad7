"""Drop-in for the event rasteriser transform of the reference data pipeline.

``EventArrToImg`` keeps the constructor and call signature of the reference class
(``mem/datasets.py:552-595``) so it can sit in the same ``transforms.Compose``;
the work is done by the CUDA rasteriser (``mem_b200.process_data.histogram``).
"""
from __future__ import annotations

from .process_data import histogram, histogram_batch

__all__ = ["EventArrToImg"]


class EventArrToImg:
    """``(N,4)`` float64 events ``[x,y,t,p]`` -> ``(H,W,3)`` uint8 ``[pos, tss|0, neg]``.

    H / W of ``None`` mean "infer from the stream" (``max + 1``), as in the reference
    (``datasets.py:571-575``); the same 100..640 range asserts are kept (``:554-557``).
    """

    def __init__(self, H=None, W=None, timeSurface=False):
        if H is not None:
            assert H >= 100 and H <= 640
        if W is not None:
            assert W >= 100 and W <= 640
        self.H, self.W = H, W
        self.timeSurface = timeSurface
        if timeSurface:
            print("Using Time Surface!")

    def __call__(self, x):
        return histogram(x, self.H, self.W, timesurface=bool(self.timeSurface), channels=3)

    def batch(self, events, offsets, channels=3):
        """GPU-side extension: rasterise a ragged batch in one launch (fixed H, W only)."""
        if self.H is None or self.W is None:
            raise ValueError("batched rasterisation needs a fixed sensor size")
        return histogram_batch(events, offsets, self.H, self.W, channels=channels,
                               timesurface=bool(self.timeSurface))
