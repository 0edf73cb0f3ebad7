"""CPU: the numpy oracle against the golden vectors produced by the reference itself
(tests/golden/histogram.npz, made by oracle/make_golden.py from mem/datasets.py:566-595)."""
import os

import numpy as np
import pytest

from oracle.histogram_ref import event_hist_batched_ref, event_hist_ref


def _cases(golden_dir):
    z = np.load(os.path.join(golden_dir, "histogram.npz"))
    names = sorted(k[:-4] for k in z.files if k.endswith("_img"))
    for n in names:
        H, W, tss = (int(v) for v in z[n + "_cfg"])
        yield n, z[n + "_ev"], (None if H < 0 else H), (None if W < 0 else W), bool(tss), z[n + "_img"]


def test_oracle_matches_reference_golden(golden_dir):
    seen = 0
    for name, ev, H, W, tss, want in _cases(golden_dir):
        got = event_hist_ref(ev, H, W, tss)
        assert got.dtype == np.uint8 and got.shape == want.shape, name
        assert np.array_equal(got, want), name
        seen += 1
    assert seen >= 10


def test_golden_contains_wraparound(golden_dir):
    # the fixture really exercises the mod-256 wrap (more hits on a pixel than a byte holds)
    z = np.load(os.path.join(golden_dir, "histogram.npz"))
    ev = z["wrap_100x100_ev"]
    flat = ev[:, 0].astype(np.int64) + 100 * ev[:, 1].astype(np.int64)
    assert np.bincount(flat[ev[:, 3] == 1]).max() > 255


def test_oracle_out_of_bounds_raises():
    ev = np.array([[99.0, 99.0, 0.0, 1.0], [0.0, 100.0, 1.0, -1.0]])
    with pytest.raises(IndexError):
        event_hist_ref(ev, 100, 100)
    # a bad coordinate on a row whose polarity is dropped is never indexed
    ev[1, 3] = 0.0
    assert event_hist_ref(ev, 100, 100).sum() == 1


def test_oracle_batched_two_channel_view():
    rng = np.random.default_rng(3)
    ev = np.stack([rng.integers(0, 120, 900), rng.integers(0, 100, 900), np.arange(900.0),
                   rng.choice([-1.0, 1.0], 900)], 1).astype(np.float64)
    off = np.array([0, 300, 300, 900])
    out3 = event_hist_batched_ref(ev, off, 100, 120, 3)
    out2 = event_hist_batched_ref(ev, off, 100, 120, 2)
    assert np.array_equal(out3[..., 0::2], out2)
    assert out3[1].sum() == 0 and out3[0].sum() == 300 and out3[2].sum() == 600


@pytest.mark.reference
def test_oracle_matches_live_reference():
    from oracle import ref_shims
    ds = ref_shims.ref_module("datasets")
    rng = np.random.default_rng(11)
    for H, W, n in [(180, 240, 5000), (480, 640, 20000)]:
        ev = np.stack([rng.uniform(0, W, n), rng.uniform(0, H, n), np.sort(rng.uniform(0, 1e5, n)),
                       rng.choice([-1.0, 1.0], n)], 1)
        for tss in (False, True):
            assert np.array_equal(ds.EventArrToImg(H, W, tss)(ev.copy()), event_hist_ref(ev, H, W, tss))
