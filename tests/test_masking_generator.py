"""CPU: host-side mask generator (product) and its oracle against the reference's golden masks."""
import os
import random

import numpy as np
import pytest

from mem_b200.masking_generator import MaskingGenerator, MaskingGeneratorRandomLocation
from oracle.masking_ref import blockwise_mask_ref


def test_blockwise_masks_match_reference_golden(golden_dir):
    z = np.load(os.path.join(golden_dir, "masks.npz"))
    ci = 0
    while f"block_{ci}_cfg" in z.files:
        gh, gw, nmask, mn = (int(v) for v in z[f"block_{ci}_cfg"])
        want = z[f"block_{ci}_masks"]
        for seed in range(want.shape[0]):
            random.seed(seed)
            got = MaskingGenerator((gh, gw), nmask, min_num_patches=mn)()
            random.seed(seed)
            ora = blockwise_mask_ref((gh, gw), nmask, mn)
            assert np.array_equal(got, want[seed]), (ci, seed)
            assert np.array_equal(ora, want[seed]), (ci, seed)
        ci += 1
    assert ci == 5


def test_consecutive_draws_consume_the_same_random_stream(golden_dir):
    want = np.load(os.path.join(golden_dir, "masks.npz"))["stream_masks"]
    random.seed(1234)
    gen = MaskingGenerator((14, 14), 75, min_num_patches=16)
    got = np.stack([gen() for _ in range(want.shape[0])])
    assert np.array_equal(got, want)


def test_random_location_masks(golden_dir, capsys):
    want = np.load(os.path.join(golden_dir, "masks.npz"))["randloc_masks"]
    for seed in range(want.shape[0]):
        random.seed(seed)
        got = MaskingGeneratorRandomLocation((14, 14), 75)()
        assert np.array_equal(got, want[seed])
        assert got.sum() == 75


def test_mask_statistics_and_api():
    gen = MaskingGenerator((14, 14), 75, min_num_patches=16)
    assert gen.get_shape() == (14, 14)
    assert repr(gen).startswith("Generator(14, 14 -> [16 ~ 75], max = 75")
    random.seed(0)
    counts = np.array([gen().sum() for _ in range(2000)])
    # mean 74.47, min 64, max 75 over 20k draws
    assert counts.max() == 75 and counts.min() >= 60 and 74.0 < counts.mean() < 75.0
    m = gen()
    assert m.dtype == np.int64 and set(np.unique(m)) <= {0, 1}


def test_batch_block():
    gen = MaskingGenerator((14, 14), 75, min_num_patches=16)
    random.seed(5)
    b = gen.batch(6, pin=False)
    random.seed(5)
    want = np.stack([gen().reshape(-1) for _ in range(6)])
    assert tuple(b.shape) == (6, 196) and np.array_equal(b.numpy(), want)
