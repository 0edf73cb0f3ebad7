"""Blockwise (BEiT) patch-mask generator -- host side of the hot path.

Drop-in for ``MaskingGenerator`` / ``MaskingGeneratorRandomLocation`` (reference
``mem/masking_generator.py:18-116``).  The sampler stays on the host on purpose:
it is a rejection sampler over Python's global ``random`` stream, a 14x14 grid
costs microseconds, and a drop-in must consume that stream in the same order
(``uniform, uniform[, randint, randint]`` per try) so that a seeded run produces
the reference's masks bit for bit.  ``batch`` adds what the GPU path wants:
a pinned ``uint8 [B, P]`` block ready for one async copy.
"""
from __future__ import annotations

import math
import random

import numpy as np

__all__ = ["MaskingGenerator", "MaskingGeneratorRandomLocation"]


class MaskingGenerator:
    def __init__(self, input_size, num_masking_patches, min_num_patches=4, max_num_patches=None,
                 min_aspect=0.3, max_aspect=None):
        if not isinstance(input_size, tuple):
            input_size = (input_size,) * 2
        self.height, self.width = input_size
        self.num_patches = self.height * self.width
        self.num_masking_patches = num_masking_patches
        self.min_num_patches = min_num_patches
        self.max_num_patches = num_masking_patches if max_num_patches is None else max_num_patches
        max_aspect = max_aspect or 1 / min_aspect
        self.log_aspect_ratio = (math.log(min_aspect), math.log(max_aspect))

    def __repr__(self):
        return "Generator(%d, %d -> [%d ~ %d], max = %d, %.3f ~ %.3f)" % (
            self.height, self.width, self.min_num_patches, self.max_num_patches,
            self.num_masking_patches, self.log_aspect_ratio[0], self.log_aspect_ratio[1])

    def get_shape(self):
        return self.height, self.width

    def _mask(self, mask, max_mask_patches):
        """Try up to 10 rectangles; paint the first one that adds 1..max new cells."""
        for _ in range(10):
            area = random.uniform(self.min_num_patches, max_mask_patches)
            aspect = math.exp(random.uniform(*self.log_aspect_ratio))
            h = int(round(math.sqrt(area * aspect)))
            w = int(round(math.sqrt(area / aspect)))
            if not (w < self.width and h < self.height):
                continue
            top = random.randint(0, self.height - h)
            left = random.randint(0, self.width - w)
            block = mask[top:top + h, left:left + w]
            new_cells = h * w - int(block.sum())
            if 0 < new_cells <= max_mask_patches:
                block.fill(1)
                return new_cells
        return 0

    def __call__(self):
        mask = np.zeros(self.get_shape(), dtype=np.int64)
        count = 0
        while count < self.num_masking_patches:
            room = min(self.num_masking_patches - count, self.max_num_patches)
            delta = self._mask(mask, room)
            if delta == 0:
                break
            count += delta
        return mask

    def batch(self, batch_size, pin=True):
        """``uint8 [B, H*W]`` torch tensor of ``batch_size`` consecutive draws (pinned if possible)."""
        import torch
        flat = np.stack([self().reshape(-1) for _ in range(batch_size)]).astype(np.uint8)
        t = torch.from_numpy(flat)
        if pin and torch.cuda.is_available():
            t = t.pin_memory()
        return t


class MaskingGeneratorRandomLocation:
    """``--masking random`` alternative (reference ``masking_generator.py:85-116``): exactly
    ``num_masking_patches`` cells drawn without replacement from the first ``H*W - 1`` cells
    (the reference's ``np.arange(max_idx)`` excludes the last cell)."""

    def __init__(self, input_size, num_masking_patches):
        if not isinstance(input_size, tuple):
            input_size = (input_size,) * 2
        self.height, self.width = input_size
        self.num_patches = self.height * self.width
        self.num_masking_patches = num_masking_patches
        print(f"Masking Ration for RandomLocation-Masker is = {self.num_masking_patches/self.num_patches}")
        assert self.num_masking_patches < self.num_patches

    def __repr__(self):
        return "Generator(patchesY: %d, patchesX %d, numMaskingPatches: %d" % (
            self.height, self.width, self.num_masking_patches)

    def get_shape(self):
        return self.height, self.width

    def __call__(self):
        mask = np.zeros(self.height * self.width, dtype=np.int64)
        population = list(range(self.height * self.width - 1))
        mask[random.sample(population, self.num_masking_patches)] = 1
        return mask.reshape(self.height, self.width)
