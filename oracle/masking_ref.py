"""Oracle (test infrastructure, not product): BEiT blockwise patch masks.

Restatement of ``MaskingGenerator`` (reference ``mem/masking_generator.py:18-81``).
The generator is a rejection sampler driven by Python's global ``random``
stream; to be a drop-in the product must consume that stream in exactly the
same order: per try ``uniform`` (area), ``uniform`` (log aspect) and -- only
when the rectangle fits strictly inside the grid -- ``randint`` (top),
``randint`` (left)  (masking_generator.py:46-52).
"""
from __future__ import annotations

import math
import random

import numpy as np


def blockwise_mask_ref(grid_hw, num_masking_patches, min_num_patches=4, max_num_patches=None,
                       min_aspect=0.3, max_aspect=None, rng=random):
    gh, gw = (grid_hw, grid_hw) if not isinstance(grid_hw, tuple) else grid_hw
    cap = num_masking_patches if max_num_patches is None else max_num_patches
    hi_aspect = max_aspect or 1 / min_aspect
    lo_log, hi_log = math.log(min_aspect), math.log(hi_aspect)

    grid = np.zeros((gh, gw), dtype=np.int64)
    placed = 0
    while placed < num_masking_patches:                       # masking_generator.py:71
        budget = min(num_masking_patches - placed, cap)       # :72-73
        gained = 0
        for _ in range(10):                                   # :46
            area = rng.uniform(min_num_patches, budget)
            ratio = math.exp(rng.uniform(lo_log, hi_log))
            bh = int(round(math.sqrt(area * ratio)))
            bw = int(round(math.sqrt(area / ratio)))
            if bw < gw and bh < gh:                           # :51
                r0 = rng.randint(0, gh - bh)
                c0 = rng.randint(0, gw - bw)
                window = grid[r0:r0 + bh, c0:c0 + bw]
                fresh = bh * bw - int(window.sum())
                if 0 < fresh <= budget:                       # :57
                    window[...] = 1
                    gained += fresh
                if gained > 0:                                # :64
                    break
        if gained == 0:                                       # :76
            break
        placed += gained
    return grid
