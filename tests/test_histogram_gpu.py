"""GPU parity: CUDA rasteriser (through the C ABI) vs the oracle, bit-exact."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle.histogram_ref import event_hist_batched_ref, event_hist_ref
from oracle.make_golden import synth_events

STRATS = [0, 1, 2, 3, 4, 5, 6, 7]  # auto, global RED, + warp aggregation, smem tile, per-SM private copy, replicated planes, hybrid, sort


def _golden(golden_dir):
    z = np.load(os.path.join(golden_dir, "histogram.npz"))
    for n in sorted(k[:-4] for k in z.files if k.endswith("_img")):
        H, W, tss = (int(v) for v in z[n + "_cfg"])
        yield n, z[n + "_ev"], (None if H < 0 else H), (None if W < 0 else W), bool(tss), z[n + "_img"]


def test_golden_vectors_all_strategies(golden_dir):
    from mem_b200.process_data import histogram
    for name, ev, H, W, tss, want in _golden(golden_dir):
        for s in STRATS:
            got = histogram(ev, H, W, timesurface=tss, strategy=s)
            assert got.dtype == np.uint8 and got.shape == want.shape, (name, s)
            assert np.array_equal(got, want), (name, s, int((got != want).sum()))


def test_event_arr_to_img_dropin(golden_dir):
    from mem_b200.datasets import EventArrToImg
    for name, ev, H, W, tss, want in _golden(golden_dir):
        assert np.array_equal(EventArrToImg(H, W, tss)(ev), want), name


@pytest.mark.parametrize("H,W", [(180, 240), (480, 640), (256, 341)])
@pytest.mark.parametrize("kind", ["uniform", "edge", "hot"])
def test_sweep_shapes_vs_oracle(H, W, kind):
    import torch
    from mem_b200.process_data import histogram
    rng = np.random.default_rng(H * 1000 + W + len(kind))      # (hash() of a str is salted per process)
    for n in (1, 31, 10_000, 300_000):
        ev = synth_events(rng, n, H, W, kind, frac=(kind == "edge"))
        want = event_hist_ref(ev, H, W)
        for s in STRATS:
            got = histogram(torch.from_numpy(ev).cuda(), H, W, strategy=s)
            assert np.array_equal(got.cpu().numpy(), want), (n, s)
        got2 = histogram(ev, H, W, channels=2)
        assert np.array_equal(got2, want[..., 0::2])


def test_heavy_wrap_and_folding_in_tile_strategy():
    # > 61440 rows per stream forces the mod-256 fold between chunks; hot pixels exceed 16 bits
    from mem_b200.process_data import histogram
    rng = np.random.default_rng(7)
    n = 400_000
    ev = synth_events(rng, n, 100, 128, "uniform")
    ev[: n // 2, 0], ev[: n // 2, 1] = 17.0, 23.0          # 200k hits on one pixel
    ev[: n // 2, 3] = np.where(np.arange(n // 2) % 3 == 0, -1.0, 1.0)
    rng.shuffle(ev)
    want = event_hist_ref(ev, 100, 128)
    for s in STRATS:
        assert np.array_equal(histogram(ev, 100, 128, strategy=s), want), s


def test_full_size_10M_events_properties():
    """BASELINE config-2 top size: check against the oracle and via linearity
    (hist(a ++ b) == hist(a) + hist(b) mod 256) and a checksum (sum of counts == #events mod 256...)."""
    import torch
    from mem_b200.process_data import histogram
    rng = np.random.default_rng(1)
    n, H, W = 10_000_000, 480, 640
    ev = synth_events(rng, n, H, W, "uniform")
    hot = rng.random(n) < 0.01
    ev[hot, 0] = rng.integers(0, 4, hot.sum()) * 7.0
    ev[hot, 1] = rng.integers(0, 4, hot.sum()) * 5.0
    d = torch.from_numpy(ev).cuda()
    full = histogram(d, H, W).cpu().numpy()
    a = histogram(d[: n // 3], H, W).cpu().numpy()
    b = histogram(d[n // 3:], H, W).cpu().numpy()
    assert np.array_equal(full, (a.astype(np.uint16) + b).astype(np.uint8))
    want = event_hist_ref(ev, H, W)
    assert np.array_equal(full, want)
    agg = histogram(d, H, W, strategy=2).cpu().numpy()
    assert np.array_equal(agg, want)


@pytest.mark.parametrize("H,W", [(480, 640), (720, 1280), (333, 517)])
def test_hybrid_and_sort_strategies_large_sensors(H, W):
    """HYBRID (hot granules privatised in shared memory, the rest through L2 REDs) and SORT (events binned by pixel class,
    shared-memory counters only) on sensors that do not fit one SM: every distribution, > 65535 hits on one pixel (mod-256
    folds of the private copy / a class far above its share), negative-wrap rows, and AUTO picking one of them for a
    long stream."""
    import torch
    from mem_b200.process_data import histogram
    rng = np.random.default_rng(H * 7 + W)
    for kind in ("uniform", "edge", "hot"):
        n = 2_200_000
        ev = synth_events(rng, n, H, W, kind, frac=(kind == "edge"))
        if kind == "hot":
            ev[: 300_000, 0], ev[: 300_000, 1] = 5.0, 3.0            # 300k hits on one pixel
            ev[: 300_000, 3] = np.where(np.arange(300_000) % 5 == 0, -1.0, 1.0)
            ev[300_000: 300_400, 0] = -ev[300_000: 300_400, 0] - 1.0   # negative flat index: numpy wraps once
            ev[300_000: 300_400, 1] = 0.0
            rng.shuffle(ev)
        want = event_hist_ref(ev, H, W)
        d = torch.from_numpy(ev).cuda()
        for s in (6, 7, 0):
            for C in (3, 2):
                got = histogram(d, H, W, channels=C, strategy=s).cpu().numpy()
                assert np.array_equal(got, want if C == 3 else want[..., 0::2]), (kind, s, C, int((got != (want if C == 3 else want[..., 0::2])).sum()))
    bad = synth_events(rng, 1_100_000, H, W)
    bad[777_777, 1] = H
    for s in (6, 7):
        with pytest.raises(IndexError):
            histogram(torch.from_numpy(bad).cuda(), H, W, strategy=s)


def test_sort_strategy_edges_of_its_layout():
    """SORT: a stream that ends inside a chunk / exactly on a chunk boundary, rows whose polarity is neither +1 nor -1
    (skipped, N-Cars 0/1 convention), an 8-byte-aligned (not 32-byte-aligned) view, a sensor whose pixel count is not a
    multiple of the 8-pixel granule, every event on one pixel (one class holds the whole stream)."""
    import torch
    from mem_b200.process_data import histogram
    rng = np.random.default_rng(11)
    H, W = 301, 533                                   # 160433 pixels: last granule partial
    for n in (8192, 8193, 3 * 8192, 100_001):
        ev = synth_events(rng, n + 1, H, W, "edge", frac=True)
        ev[::7, 3] = 0.0
        want = event_hist_ref(ev[1:], H, W)
        d = torch.from_numpy(ev).cuda()[1:]           # rows start 32 bytes into the allocation: still aligned
        assert np.array_equal(histogram(d, H, W, strategy=7).cpu().numpy(), want), n
        flat = torch.from_numpy(np.concatenate([[0.0], ev[1:].ravel()])).cuda()[1:].view(-1, 4)   # 8-byte aligned only
        assert np.array_equal(histogram(flat, H, W, strategy=7).cpu().numpy(), want), n
    one = np.zeros((700_000, 4))
    one[:, 0], one[:, 1], one[:, 3] = 532.0, 300.0, np.where(np.arange(700_000) % 3 == 0, -1.0, 1.0)
    assert np.array_equal(histogram(torch.from_numpy(one).cuda(), H, W, strategy=7).cpu().numpy(), event_hist_ref(one, H, W))


def test_ragged_batch_training_shape():
    import torch
    from mem_b200.process_data import histogram_batch
    rng = np.random.default_rng(2)
    H, W, B = 256, 341, 24
    lens = rng.integers(0, 30001, B)
    lens[3] = 0
    lens[5] = 30000
    off = np.concatenate([[0], np.cumsum(lens)])
    ev = synth_events(rng, int(off[-1]), H, W, "edge", frac=True)
    for C in (2, 3):
        want = event_hist_batched_ref(ev, off, H, W, C)
        for s in STRATS:
            got = histogram_batch(ev, off, H, W, channels=C, strategy=s)
            assert np.array_equal(got, want), (C, s)
    # device-resident inputs, odd sensor (unaligned rows) and time surface
    H2, W2 = 101, 103
    ev2 = synth_events(rng, int(off[-1]), H2, W2)
    lens2 = np.maximum(lens, 1)
    off2 = np.concatenate([[0], np.cumsum(lens2)])
    ev2 = synth_events(rng, int(off2[-1]), H2, W2)
    want = event_hist_batched_ref(ev2, off2, H2, W2, 3, timesurface=True)
    got = histogram_batch(torch.from_numpy(ev2).cuda(), torch.from_numpy(off2).cuda(), H2, W2, 3, True)
    assert np.array_equal(got.cpu().numpy(), want)
    for s in STRATS:
        got = histogram_batch(ev2, off2, H2, W2, channels=3, strategy=s)
        assert np.array_equal(got, event_hist_batched_ref(ev2, off2, H2, W2, 3)), s


def test_out_of_bounds_raises_index_error_like_numpy():
    from mem_b200.process_data import histogram
    ev = np.array([[99.0, 99.0, 0.0, 1.0], [0.0, 100.0, 1.0, -1.0]])
    for s in STRATS:
        with pytest.raises(IndexError):
            histogram(ev, 100, 100, strategy=s)
    ev[1, 3] = 0.0      # dropped polarity: never indexed, no error
    assert histogram(ev, 100, 100).sum() == 1
    ev[1] = [np.nan, 0.0, 1.0, 1.0]
    with pytest.raises(IndexError):
        histogram(ev, 100, 100)


def test_empty_and_inferred_extent():
    from mem_b200.process_data import histogram
    assert histogram(np.zeros((0, 4)), 100, 120).sum() == 0
    with pytest.raises(ValueError):
        histogram(np.zeros((0, 4)))
    ev = np.array([[5.9, 7.2, 0.0, 1.0], [11.0, 3.0, 1.0, -1.0]])
    out = histogram(ev)
    assert out.shape == (8, 12, 3) and np.array_equal(out, event_hist_ref(ev))


def test_unaligned_event_pointer():
    import torch
    from mem_b200.process_data import histogram
    rng = np.random.default_rng(9)
    ev = synth_events(rng, 5001, 120, 160)
    flat = torch.zeros(5001 * 4 + 1, dtype=torch.float64, device="cuda")
    view = flat[1:].view(5001, 4)            # 8-byte aligned only
    view.copy_(torch.from_numpy(ev))
    assert view.data_ptr() % 32 != 0
    want = event_hist_ref(ev, 120, 160)
    for s in STRATS:
        assert np.array_equal(histogram(view, 120, 160, strategy=s).cpu().numpy(), want)


def test_recordings_in_narrow_dtypes_are_widened_on_the_device():
    """N-ImageNet event_data arrays / integer exports: uploaded in their own dtype, widened to float64 rows on the GPU."""
    from mem_b200.process_data import histogram
    rng = np.random.default_rng(11)
    ev = np.floor(synth_events(rng, 50_000, 480, 640, "edge"))
    want = event_hist_ref(ev, 480, 640)
    for dt in (np.int16, np.int32, np.int64, np.float32):
        assert np.array_equal(histogram(ev.astype(dt), 480, 640), want), dt
    pol01 = ev.copy()
    pol01[:, 3] = (pol01[:, 3] > 0)
    assert np.array_equal(histogram(pol01.astype(np.int16), 480, 640), event_hist_ref(pol01, 480, 640))


def test_hypothesis_small_streams():
    from hypothesis import given, settings, strategies as st
    from mem_b200.process_data import histogram

    @settings(max_examples=40, deadline=None)
    @given(st.integers(0, 2**31), st.integers(1, 700), st.integers(100, 140), st.integers(100, 140),
           st.sampled_from([0, 1, 2, 3]))
    def run(seed, n, H, W, strat):
        rng = np.random.default_rng(seed)
        ev = np.stack([rng.uniform(-0.9, W - 0.01, n), rng.uniform(0, H - 0.01, n), rng.uniform(0, 10, n),
                       rng.choice([-1.0, 1.0, 0.0, 2.0], n)], 1)
        assert np.array_equal(histogram(ev, H, W, strategy=strat), event_hist_ref(ev, H, W))

    run()
