"""GPU parity: fused event augmentation + rasteriser + post-raster transforms (through the C ABI) against
the reference's own build_transformNPY outputs (tests/golden/event_pipeline.npz) and the oracle.  Bit-exact."""
import os
import random

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle.event_pipeline_ref import PipelineCfg, VarPipelineCfg, apply_event_aug, apply_post_raster, pipeline_ref
from oracle.histogram_ref import event_hist_ref
from oracle.make_golden import synth_events

STRATS = [0, 1, 2, 3, 4]


def seed_all(seed):
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)


def product_cfg(cfg: PipelineCfg):
    from mem_b200.event_pipeline import PipelineConfig
    return PipelineConfig(**cfg.__dict__)


def golden_cases(golden_dir):
    z = np.load(os.path.join(golden_dir, "event_pipeline.npz"))
    for name in sorted(k[:-4] for k in z.files if k.endswith("_out")):
        is_train, n, norm, seed = (int(v) for v in z[name + "_meta"])
        kind = str(z[name + "_kind"])
        ev = synth_events(np.random.default_rng(seed), n, 480, 640, kind, frac=(kind == "edge"))
        yield name, ev, PipelineCfg(is_train=bool(is_train), normalize_events=bool(norm)), seed, z[name + "_out"]


def test_reference_golden_single_streams(golden_dir):
    """Same generator seeds as the reference run -> same draws -> identical float32 tensors."""
    from mem_b200.event_pipeline import EventBatchPipeline
    seen = 0
    for name, ev, cfg, seed, want in golden_cases(golden_dir):
        for fused in (True, False):       # one-kernel path and rasterise + post-raster path
            seed_all(seed)
            got = EventBatchPipeline(product_cfg(cfg), fused=fused)([ev])
            assert got.is_cuda and got.dtype == torch.float32 and tuple(got.shape) == (1,) + want.shape, name
            assert np.array_equal(got[0].cpu().numpy(), want), (name, fused, float(np.abs(got[0].cpu().numpy() - want).max()))
        seen += 1
    assert seen >= 6


def test_draws_match_oracle_draws():
    from mem_b200 import event_pipeline as ep
    from oracle import event_pipeline_ref as ref
    for is_train in (True, False):
        for n in (100, 30000, 30001, 45000):
            seed_all(n + is_train)
            a = ep.draw_params(n, ep.PipelineConfig(is_train=is_train))
            seed_all(n + is_train)
            b = ref.draw_params(n, ref.PipelineCfg(is_train=is_train))
            assert a == b


@pytest.mark.parametrize("channels", [2, 3])
def test_batch_against_oracle_all_strategies(channels):
    """Ragged batch (one empty stream, one fully culled, streams shorter and longer than the slice window)."""
    from mem_b200.event_pipeline import (PipelineConfig, draw_params, pack_params, pipeline_fused, post_raster,
                                         rasterise_augmented)
    rng = np.random.default_rng(3)
    cfg = PipelineCfg(is_train=True, normalize_events=True)
    pcfg = product_cfg(cfg)
    H, W = cfg.raster_hw()
    lens = [45000, 0, 7, 30001, 12345, 60000, 1000]
    streams = [synth_events(rng, n, 480, 640, kind, frac=True) if n else np.zeros((0, 4))
               for n, kind in zip(lens, ["edge", "uniform", "uniform", "hot", "edge", "uniform", "uniform"])]
    seed_all(17)
    params = [draw_params(n, pcfg) for n in lens]
    params[6].update(shift_x=400)       # every row of stream 6 leaves the sensor: an all-zero image
    aug, crop = pack_params(params)
    offsets = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    events = np.concatenate(streams, axis=0)
    def raster(s, p):
        # an empty stream makes the reference's time flip raise (x[0, 2] of nothing); the batched path returns zeros
        a = apply_event_aug(s, p) if len(s) else s
        return event_hist_ref(a, H, W) if len(a) else np.zeros((H, W, 3), np.uint8)
    want_hist = np.stack([raster(s, p) for s, p in zip(streams, params)])
    assert want_hist[6].sum() == 0 and want_hist[0].sum() > 0
    for strat in STRATS:
        got = rasterise_augmented(torch.from_numpy(events).cuda(), offsets, aug, H, W, channels, strategy=strat,
                                  max_stream_len=30000)
        want = want_hist if channels == 3 else want_hist[..., 0::2]
        assert np.array_equal(got.cpu().numpy(), want), (strat, int((got.cpu().numpy() != want).sum()))
    out = post_raster(got, crop, (224, 224), hot_num_stds=10.0, normalize=True)
    one = pipeline_fused(torch.from_numpy(events).cuda(), offsets, aug, crop, H, W, (224, 224), channels,
                         hot_num_stds=10.0, normalize=True)
    for b, p in enumerate(params):
        w = apply_post_raster(want_hist[b], p, cfg).numpy()
        w = w if channels == 3 else w[0::2]
        assert np.array_equal(out[b].cpu().numpy(), w), b
        assert np.array_equal(one[b].cpu().numpy(), w), ("fused", b)


@pytest.mark.parametrize("H,W", [(200, 180), (224, 300), (256, 341)])
@pytest.mark.parametrize("hot,norm", [(None, False), (10.0, False), (3.0, True), (None, True)])
def test_post_raster_padding_and_switches(H, W, hot, norm):
    """RandomCrop(pad_if_needed) geometry: the image is padded on both sides by the missing amount."""
    from mem_b200.event_pipeline import post_raster
    rng = np.random.default_rng(H * W)
    B = 5
    hist = (rng.random((B, H, W, 3)) < 0.05).astype(np.uint8) * rng.integers(1, 6, (B, H, W, 3)).astype(np.uint8)
    hist[:, 10, 10, 0] = 255
    hist[:, 150, 100, 2] = 200
    hist[4] = 0                                             # all-zero image: max == 0 -> NormalizeEvent leaves it
    ph, pw = H + 2 * max(224 - H, 0), W + 2 * max(224 - W, 0)
    tl = np.stack([rng.integers(0, ph - 224 + 1, B), rng.integers(0, pw - 224 + 1, B)], axis=1).astype(np.int32)
    cfg = PipelineCfg(is_train=True, hotpixfilter=hot is not None, hotpix_num_stds=hot if hot is not None else 10,
                      normalize_events=norm)
    got = post_raster(torch.from_numpy(hist).cuda(), tl, (224, 224), hot_num_stds=hot, normalize=norm)
    for b in range(B):
        want = apply_post_raster(hist[b], dict(top=int(tl[b, 0]), left=int(tl[b, 1])), cfg).numpy()
        assert np.array_equal(got[b].cpu().numpy(), want), (b, float(np.abs(got[b].cpu().numpy() - want).max()))


def test_fused_padding_small_raster():
    """Raster smaller than the crop (RandomCrop pads both sides): one-kernel path vs rasterise + oracle post."""
    from mem_b200.event_pipeline import AUG_DTYPE, pipeline_fused
    rng = np.random.default_rng(8)
    H, W, B, n = 200, 180, 3, 20000
    ev = np.concatenate([synth_events(rng, n, H, W, "hot") for _ in range(B)], axis=0)
    off = np.arange(B + 1, dtype=np.int64) * n
    aug = np.zeros(B, dtype=AUG_DTYPE)
    aug["scale_x"] = aug["scale_y"] = 1.0
    aug["count"] = -1
    tl = np.array([[0, 0], [24, 44], [12, 20]], dtype=np.int32)     # padded image is 248 x 268
    cfg = PipelineCfg(is_train=True, normalize_events=True, hotpix_num_stds=5)
    got = pipeline_fused(ev, off, aug, tl, H, W, (224, 224), 3, hot_num_stds=5.0, normalize=True)
    for b in range(B):
        hist = event_hist_ref(ev[off[b]:off[b + 1]], H, W)
        want = apply_post_raster(hist, dict(top=int(tl[b, 0]), left=int(tl[b, 1])), cfg).numpy()
        assert np.array_equal(got[b].cpu().numpy(), want), b


def test_full_size_properties():
    """B = 128 streams x 30000 events (the training batch of BASELINE config 3), checked through properties:
    an x flip mirrors the image, a time flip swaps the polarity channels, a shift translates it."""
    from mem_b200.event_pipeline import AUG_DTYPE, rasterise_augmented
    from mem_b200.process_data import histogram_batch
    B, n, H, W = 128, 30000, 256, 341
    g = torch.Generator(device="cuda").manual_seed(0)
    ev = torch.empty(B * n, 4, dtype=torch.float64, device="cuda")
    ev[:, 0] = torch.randint(0, W, (B * n,), generator=g, device="cuda").double()
    ev[:, 1] = torch.randint(0, H, (B * n,), generator=g, device="cuda").double()
    ev[:, 2] = torch.rand(B * n, generator=g, device="cuda", dtype=torch.float64)
    ev[:, 3] = torch.randint(0, 2, (B * n,), generator=g, device="cuda").double() * 2 - 1
    off = torch.arange(B + 1, device="cuda", dtype=torch.int64) * n
    base = histogram_batch(ev, off, H, W, channels=2, max_stream_len=n)
    aug = np.zeros(B, dtype=AUG_DTYPE)
    aug["scale_x"] = aug["scale_y"] = 1.0
    aug["count"] = -1
    aug["flip_w"], aug["cull_w"], aug["cull_h"] = W, W, H
    ident = rasterise_augmented(ev, off, aug, H, W, 2, max_stream_len=n)
    assert torch.equal(ident, base)
    a = aug.copy(); a["flip_x"] = 1
    assert torch.equal(rasterise_augmented(ev, off, a, H, W, 2, max_stream_len=n), base.flip(2))
    a = aug.copy(); a["time_flip"] = 1
    assert torch.equal(rasterise_augmented(ev, off, a, H, W, 2, max_stream_len=n), base.flip(3))
    a = aug.copy(); a["cull"] = 1; a["shift_x"] = 7; a["shift_y"] = -5
    got = rasterise_augmented(ev, off, a, H, W, 2, max_stream_len=n)
    want = torch.zeros_like(base)
    want[:, :H - 5, 7:] = base[:, 5:, :W - 7]
    assert torch.equal(got, want)
    a = aug.copy(); a["start"] = 1000; a["count"] = 5000
    got = rasterise_augmented(ev, off, a, H, W, 2, max_stream_len=5000)
    win = ev.view(B, n, 4)[:, 1000:6000].reshape(-1, 4).contiguous()
    want = histogram_batch(win, torch.arange(B + 1, device="cuda", dtype=torch.int64) * 5000, H, W, channels=2)
    assert torch.equal(got, want)


def test_bad_arguments_fail_loudly():
    from mem_b200 import _lib
    from mem_b200.event_pipeline import AUG_DTYPE, EventBatchPipeline, PipelineConfig, post_raster, rasterise_augmented
    with pytest.raises(NotImplementedError):
        EventBatchPipeline(PipelineConfig(timesurface=True), channels=2)        # the time surface is the middle of 3 channels
    ev = np.zeros((4, 4))
    with pytest.raises(ValueError):
        rasterise_augmented(ev, np.array([0, 4]), np.zeros(2, dtype=AUG_DTYPE), 100, 100)
    with pytest.raises(ValueError):
        post_raster(torch.zeros(1, 8, 8, 3), None)
    # an eval-path stream that indexes outside the sensor raises like the reference's np.add.at
    aug = np.zeros(1, dtype=AUG_DTYPE); aug["scale_x"] = aug["scale_y"] = 1.0; aug["count"] = -1
    bad = np.array([[100.0, 100.0, 0.0, 1.0]])
    with pytest.raises((IndexError, _lib.MembError)):
        rasterise_augmented(bad, np.array([0, 1]), aug, 100, 100)


def test_randomised_parameters_fused_vs_oracle():
    """40 random parameter sets (windows, scales, flips, shifts beyond the reference's range, cull on / off with
    fractional and slightly negative coordinates that exercise numpy's truncation and negative-index wrap) through the
    one-kernel path and the two-stage path, against the oracle chain."""
    from mem_b200.event_pipeline import pack_params, pipeline_fused, post_raster, rasterise_augmented
    rng = np.random.default_rng(2024)
    H, W, oh, ow = 200, 230, 160, 176
    for case in range(40):
        B = int(rng.integers(1, 5))
        lens = [int(rng.integers(0, 4000)) for _ in range(B)]
        streams, params = [], []
        for n in lens:
            cull = bool(rng.integers(0, 2))
            sx, sy = (float(rng.uniform(0.3, 1.0)), float(rng.uniform(0.3, 1.0))) if rng.integers(0, 2) else (1.0, 1.0)
            ev = np.stack([rng.uniform(-0.9 if not cull else -30, (W - 1) / sx if not cull else W / sx + 30, n),
                           rng.uniform(0.0 if not cull else -30, (H - 1) / sy if not cull else H / sy + 30, n),
                           np.sort(rng.uniform(0, 1e5, n)), rng.choice([-1.0, 1.0, 0.0], n, p=[0.45, 0.45, 0.1])], axis=1)
            start = int(rng.integers(0, max(1, n)))
            p = dict(scale_x=sx, scale_y=sy, start=start, count=int(rng.integers(0, n - start + 1)) if n else 0,
                     time_flip=bool(rng.integers(0, 2)), flip_x=bool(rng.integers(0, 2)) and cull, flip_w=W, cull=cull,
                     shift_x=int(rng.integers(-40, 41)) if cull else 0, shift_y=int(rng.integers(-40, 41)) if cull else 0,
                     cull_w=W, cull_h=H, top=int(rng.integers(0, H - oh + 1)), left=int(rng.integers(0, W - ow + 1)))
            streams.append(ev)
            params.append(p)
        aug, crop = pack_params(params)
        off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
        events = np.concatenate(streams, axis=0) if sum(lens) else np.zeros((0, 4))
        hot = [None, 10.0, 2.5][case % 3]
        norm = bool(case % 2)
        cfg = PipelineCfg(is_train=True, input_H=oh, input_W=ow, hotpixfilter=hot is not None,
                          hotpix_num_stds=hot if hot is not None else 10, normalize_events=norm)
        one = pipeline_fused(events, off, aug, crop, H, W, (oh, ow), 3, hot_num_stds=hot, normalize=norm)
        two = post_raster(rasterise_augmented(events, off, aug, H, W, 3), crop, (oh, ow), hot_num_stds=hot, normalize=norm)
        for b, (s, p) in enumerate(zip(streams, params)):
            win = s[p["start"]:p["start"] + p["count"]]
            a = apply_event_aug(s, p) if len(win) else win
            hist = event_hist_ref(a, H, W) if len(a) else np.zeros((H, W, 3), np.uint8)
            want = apply_post_raster(hist, p, cfg).numpy()
            assert np.array_equal(one[b].cpu().numpy(), want), (case, b, "fused")
            assert np.array_equal(two[b].cpu().numpy(), want), (case, b, "two-stage")


# ------------------------------------------------------------------ variable sensor size (N-Caltech101 / N-Cars branch)
def _var_cases(golden_dir):
    from oracle.event_pipeline_ref import VarPipelineCfg
    z = np.load(os.path.join(golden_dir, "event_pipeline_var.npz"))
    for name in sorted(k[:-4] for k in z.files if k.endswith("_out")):
        is_train, n, norm, seed, H, W, pol01 = (int(v) for v in z[name + "_meta"])
        kind = str(z[name + "_kind"])
        ev = np.floor(synth_events(np.random.default_rng(seed), n, H, W, kind, polarity=(0.0, 1.0) if pol01 else (-1.0, 1.0)))
        yield name, ev, VarPipelineCfg(is_train=bool(is_train), normalize_events=bool(norm)), seed, (H, W), z[name + "_out"]


# The resize evaluates ATen's anti-aliased triangle filter in float32 with fused multiply-adds; ATen's CPU kernel rounds
# every product separately: identical taps and weights, last-bit differences in the sums.  Everything before the resize
# is integer-exact, everything after it is a comparison / one multiply.
VAR_ATOL = 4e-7


def _close_images(got, want, name):
    d = np.abs(got - want)
    # a pixel whose value sits within rounding of the hot-pixel threshold may be filtered on one side only
    flipped = (got == 0) != (want == 0)
    assert int(flipped.sum()) <= 2, (name, int(flipped.sum()))
    assert float(d[~flipped].max()) <= VAR_ATOL * max(1.0, float(np.abs(want).max())), (name, float(d[~flipped].max()))


def test_variable_sensor_reference_golden(golden_dir):
    from mem_b200.event_pipeline import EventBatchPipelineVar, VarPipelineConfig
    seen = 0
    for name, ev, cfg, seed, (H, W), want in _var_cases(golden_dir):
        pc = VarPipelineConfig(is_train=cfg.is_train, canvas_H=180, canvas_W=240, normalize_events=cfg.normalize_events)
        seed_all(seed)
        got = EventBatchPipelineVar(pc)([ev])
        assert got.is_cuda and got.dtype == torch.float32 and tuple(got.shape) == (1,) + want.shape, name
        _close_images(got[0].cpu().numpy(), want, name)
        seen += 1
    assert seen == 6


# log(x + 1) / x ** gamma of the resized planes: this side rounds the double-precision value once, torch's float32 kernels
# are within 1 ulp of that per map (gamma = 0.5 is a square root on both sides).  Everything after the resize is
# multiplicative, so the bound is relative: the resize's summation-order differences (a few ulp of all-positive sums),
# stretched by gamma when gamma > 1, plus the maps' own ulp and the normalising multiply (measured worst case over the
# golden set: 1.9e-7).
VAR_TF_RTOL = 5e-7


def _var_loggamma_cases(golden_dir):
    z = np.load(os.path.join(golden_dir, "event_pipeline_var_loggamma.npz"))
    for name in sorted(k[:-4] for k in z.files if k.endswith("_out")):
        is_train, n, norm, lg, gm, seed, H, W, pol01 = (int(v) for v in z[name + "_meta"])
        ev = np.floor(synth_events(np.random.default_rng(seed), n, H, W, str(z[name + "_kind"]),
                                   polarity=(0.0, 1.0) if pol01 else (-1.0, 1.0)))
        yield name, ev, dict(is_train=bool(is_train), normalize_events=bool(norm), logtrafo=bool(lg), gammatrafo=bool(gm),
                             gamma=float(z[name + "_gamma"])), seed, z[name + "_out"]


def test_variable_sensor_log_gamma_reference_golden(golden_dir):
    """args.logtrafo / args.gammatrafo on the variable-sensor branch (tests/golden/event_pipeline_var_loggamma.npz: outputs
    of the reference's own build_transformNPY), gamma in {0.5, 0.7, 1.6}."""
    from mem_b200.event_pipeline import EventBatchPipelineVar, VarPipelineConfig
    seen, worst = 0, {}
    for name, ev, kw, seed, want in _var_loggamma_cases(golden_dir):
        seed_all(seed)
        got = EventBatchPipelineVar(VarPipelineConfig(canvas_H=180, canvas_W=240, **kw))([ev])[0].cpu().numpy()
        assert got.shape == want.shape, name
        flipped = (got == 0) != (want == 0)
        assert int(flipped.sum()) <= 2, (name, int(flipped.sum()))
        rel = (np.abs(got - want) / np.maximum(np.abs(want), 1e-30))[~flipped & (want != 0)]
        worst[name] = float(rel.max())
        seen += 1
    print("var log/gamma worst relative error:", worst)
    assert seen == 5 and max(worst.values()) <= VAR_TF_RTOL, worst
    with pytest.raises(ValueError):
        EventBatchPipelineVar(VarPipelineConfig(gammatrafo=True, gamma=-1.0))([ev])


def test_variable_sensor_time_surface_reference_golden(golden_dir):
    """args.timesurface on the variable-sensor branch (tests/golden/event_pipeline_var_tss.npz: the reference chain's outputs,
    both RandomTimeFlip outcomes): the time-surface bytes are exact, so the middle plane meets the same bound as the
    polarity planes (summation order of the resize); a ragged batch against the oracle with shared draws."""
    from mem_b200.event_pipeline import EventBatchPipelineVar, VarPipelineConfig, draw_params_var
    from oracle.event_pipeline_ref import pipeline_var_ref
    z = np.load(os.path.join(golden_dir, "event_pipeline_var_tss.npz"))
    seen, streams = 0, []
    for name in sorted(k[:-4] for k in z.files if k.endswith("_out")):
        is_train, n, norm, lg, seed, H, W, pol01 = (int(v) for v in z[name + "_meta"])
        ev = np.floor(synth_events(np.random.default_rng(seed), n, H, W, str(z[name + "_kind"]),
                                   polarity=(0.0, 1.0) if pol01 else (-1.0, 1.0)))
        pc = VarPipelineConfig(is_train=bool(is_train), canvas_H=180, canvas_W=240, normalize_events=bool(norm), logtrafo=bool(lg),
                               timesurface=True)
        seed_all(seed)
        got = EventBatchPipelineVar(pc)([ev])[0].cpu().numpy()
        want = z[name + "_out"]
        assert got.shape == want.shape and (got[1] != 0).any(), name
        d = np.abs(got[1] - want[1])
        assert float(d.max()) <= VAR_ATOL, (name, "time surface", float(d.max()))
        if lg:
            rel = np.abs(got[0::2] - want[0::2]) / np.maximum(np.abs(want[0::2]), 1e-30)
            assert float(rel[want[0::2] != 0].max()) <= VAR_TF_RTOL and int(((got[0::2] == 0) != (want[0::2] == 0)).sum()) <= 2, name
        else:
            _close_images(got[0::2], want[0::2], name)
        streams.append(ev)
        seen += 1
    assert seen == 6
    pc = VarPipelineConfig(is_train=True, canvas_H=180, canvas_W=240, normalize_events=True, timesurface=True)
    seed_all(77)
    params = [draw_params_var(len(s), pc) for s in streams]
    assert len({p["time_flip"] for p in params}) == 2
    got = EventBatchPipelineVar(pc)(streams, params=params).cpu().numpy()
    for b, (s, p) in enumerate(zip(streams, params)):
        want = pipeline_var_ref(s, VarPipelineCfg(is_train=True, normalize_events=True, timesurface=True), p).numpy()
        assert float(np.abs(got[b, 1] - want[1]).max()) <= VAR_ATOL, b
        _close_images(got[b, 0::2], want[0::2], b)
    with pytest.raises(ValueError):
        EventBatchPipelineVar(pc, channels=2)


def test_variable_sensor_batch_vs_oracle():
    """A ragged batch (different extents, polarities, lengths; int16 rows uploaded as stored) against the oracle chain
    with shared draws; C = 2; errors where the reference raises."""
    from mem_b200.event_pipeline import EventBatchPipelineVar, VarPipelineConfig, draw_params_var
    from oracle.event_pipeline_ref import VarPipelineCfg, pipeline_var_ref
    rng = np.random.default_rng(5)
    shapes = [(180, 240), (100, 120), (150, 200), (64, 300), (180, 240)]
    streams = [np.floor(synth_events(rng, int(rng.integers(6000, 50000)), h, w, kind))
               for (h, w), kind in zip(shapes, ["edge", "uniform", "hot", "edge", "uniform"])]
    for is_train in (True, False):
        pc = VarPipelineConfig(is_train=is_train, canvas_H=180, canvas_W=240, normalize_events=True)
        use = [s for s, (h, w) in zip(streams, shapes) if w <= 240]
        seed_all(31)
        params = [draw_params_var(len(s), pc) for s in use]
        want = torch.stack([pipeline_var_ref(s, VarPipelineCfg(is_train=is_train, normalize_events=True), p) for s, p in zip(use, params)])
        got = EventBatchPipelineVar(pc)([s.astype(np.int16) for s in use], params=params)
        for b in range(len(use)):
            _close_images(got[b].cpu().numpy(), want[b].numpy(), (is_train, b))
        got2 = EventBatchPipelineVar(pc, channels=2)(use, params=params)
        assert torch.equal(got2, got[:, 0::2])
    # a recording wider than the canvas, and a stream emptied by the shift
    pc = VarPipelineConfig(is_train=True, canvas_H=180, canvas_W=240)
    with pytest.raises(ValueError):
        EventBatchPipelineVar(pc)([streams[3]])
    with pytest.raises(ValueError):
        EventBatchPipelineVar(pc)([np.array([[0.0, 0.0, 1.0, 1.0]])],
                                  params=[dict(draw_params_var(1, pc), shift_x=-5, cull=True)])


def test_log_and_gamma_transforms_reference_golden(golden_dir):
    """args.logtrafo / args.gammatrafo: the fused kernel with the host-evaluated value table reproduces the reference's
    build_transformNPY outputs bit for bit (tests/golden/event_pipeline_loggamma.npz)."""
    from mem_b200.event_pipeline import EventBatchPipeline, PipelineConfig
    z = np.load(os.path.join(golden_dir, "event_pipeline_loggamma.npz"))
    names = sorted(k[:-4] for k in z.files if k.endswith("_out"))
    assert len(names) == 4
    for name in names:
        is_train, n, norm, lg, gm, seed = (int(v) for v in z[name + "_meta"])
        kind = str(z[name + "_kind"])
        ev = synth_events(np.random.default_rng(seed), n, 480, 640, kind, frac=(kind == "edge"))
        cfg = PipelineConfig(is_train=bool(is_train), normalize_events=bool(norm), logtrafo=bool(lg), gammatrafo=bool(gm),
                             gamma=float(z[name + "_gamma"]))
        want = z[name + "_out"]
        for fused in (True, False):           # one kernel / rasterise + post-raster pass: the same table, the same bytes
            seed_all(seed)
            got = EventBatchPipeline(cfg, fused=fused)([ev])
            assert tuple(got.shape) == (1,) + want.shape
            assert np.array_equal(got[0].cpu().numpy(), want), (name, fused, float(np.abs(got[0].cpu().numpy() - want).max()))


def test_time_surface_with_log_gamma_reference_golden(golden_dir):
    """args.timesurface with args.logtrafo / args.gammatrafo (tests/golden/event_pipeline_tss_loggamma.npz, outputs of the
    reference's build_transformNPY): the value table rides on the post-raster pass, the time-surface plane stays c / 255;
    bit-exact."""
    from mem_b200.event_pipeline import EventBatchPipeline, PipelineConfig
    z = np.load(os.path.join(golden_dir, "event_pipeline_tss_loggamma.npz"))
    names = sorted(k[:-4] for k in z.files if k.endswith("_out"))
    assert len(names) == 4
    for name in names:
        is_train, n, norm, lg, gm, seed = (int(v) for v in z[name + "_meta"])
        kind = str(z[name + "_kind"])
        ev = synth_events(np.random.default_rng(seed), n, 480, 640, kind, frac=(kind == "edge"))
        cfg = PipelineConfig(is_train=bool(is_train), normalize_events=bool(norm), logtrafo=bool(lg), gammatrafo=bool(gm),
                             gamma=float(z[name + "_gamma"]), timesurface=True)
        seed_all(seed)
        got = EventBatchPipeline(cfg)([ev])[0].cpu().numpy()
        want = z[name + "_out"]
        assert got.shape == want.shape and (got[1] != 0).any(), name
        assert np.array_equal(got, want), (name, float(np.abs(got - want).max()))


def test_time_surface_with_augmentations_reference_golden(golden_dir):
    """args.timesurface=1 through the rasterise (memb_hist_aug_tss_u8) + post-raster pair: timestamps normalised over the
    rows that survive window / shift / cull, RandomTimeFlip's reversed order and t' = t_last - t, middle channel kept;
    bit-exact with the reference's build_transformNPY outputs.  Also a ragged batch against the oracle with shared draws."""
    from mem_b200.event_pipeline import EventBatchPipeline, PipelineConfig, draw_params
    z = np.load(os.path.join(golden_dir, "event_pipeline_tss.npz"))
    names = sorted(k[:-4] for k in z.files if k.endswith("_out"))
    assert len(names) == 6
    streams = []
    for name in names:
        is_train, n, norm, seed = (int(v) for v in z[name + "_meta"])
        kind = str(z[name + "_kind"])
        ev = synth_events(np.random.default_rng(seed), n, 480, 640, kind, frac=(kind == "edge"))
        streams.append(ev)
        cfg = PipelineConfig(is_train=bool(is_train), normalize_events=bool(norm), timesurface=True)
        seed_all(seed)
        got = EventBatchPipeline(cfg)([ev])
        want = z[name + "_out"]
        assert tuple(got.shape) == (1,) + want.shape
        assert np.array_equal(got[0].cpu().numpy(), want), (name, int((got[0].cpu().numpy() != want).sum()))
    cfg = PipelineConfig(is_train=True, normalize_events=True, timesurface=True)
    seed_all(77)
    params = [draw_params(len(s), cfg) for s in streams]
    assert any(p["time_flip"] for p in params) and not all(p["time_flip"] for p in params)
    got = EventBatchPipeline(cfg)(streams, params=params).cpu().numpy()
    for b, (s, p) in enumerate(zip(streams, params)):
        want = pipeline_ref(s, PipelineCfg(is_train=True, normalize_events=True, timesurface=True), p).numpy()
        assert np.array_equal(got[b], want), b
    with pytest.raises(NotImplementedError):
        EventBatchPipeline(cfg, fused=True)
