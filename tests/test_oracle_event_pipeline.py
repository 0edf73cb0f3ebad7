"""CPU: the event-pipeline oracle (oracle/event_pipeline_ref.py) against golden tensors produced by the reference's own
``build_transformNPY`` chain (tests/golden/event_pipeline.npz, made by oracle/make_golden.py::golden_event_pipeline
from mem/datasets.py:611-660 + mem/transforms.py:225-275), under the same generator seeds."""
import os
import random

import numpy as np
import pytest
import torch

from oracle.event_pipeline_ref import PipelineCfg, apply_event_aug, draw_params, pipeline_ref
from oracle.make_golden import synth_events


def golden_cases(golden_dir):
    z = np.load(os.path.join(golden_dir, "event_pipeline.npz"))
    for name in sorted(k[:-4] for k in z.files if k.endswith("_out")):
        is_train, n, norm, seed = (int(v) for v in z[name + "_meta"])
        kind = str(z[name + "_kind"])
        ev = synth_events(np.random.default_rng(seed), n, 480, 640, kind, frac=(kind == "edge"))
        cfg = PipelineCfg(is_train=bool(is_train), normalize_events=bool(norm))
        yield name, ev, cfg, seed, z[name + "_out"]


def seed_all(seed):
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)


def test_oracle_matches_reference_golden(golden_dir):
    seen = 0
    for name, ev, cfg, seed, want in golden_cases(golden_dir):
        seed_all(seed)
        got = pipeline_ref(ev, cfg).numpy()
        assert got.shape == want.shape and got.dtype == np.float32, name
        assert np.array_equal(got, want), f"{name}: {np.abs(got - want).max()}"
        seen += 1
    assert seen >= 6


def test_draws_follow_the_reference_order():
    cfg = PipelineCfg()
    seed_all(5)
    p = draw_params(45000, cfg)
    seed_all(5)
    start = random.choice(range(45000 - 30000 + 1))
    tf, fx = np.random.random() < 0.5, np.random.random() < 0.5
    sx, sy = np.random.randint(-15, 16, size=(2,))
    top = int(torch.randint(0, 256 - 224 + 1, size=(1,)).item())
    left = int(torch.randint(0, 341 - 224 + 1, size=(1,)).item())
    assert (p["start"], p["count"], p["time_flip"], p["flip_x"], p["shift_x"], p["shift_y"], p["top"], p["left"]) == \
        (start, 30000, tf, fx, int(sx), int(sy), top, left)
    # short streams do not consume Python's generator (datasets.py:495)
    seed_all(5)
    state = random.getstate()
    draw_params(100, cfg)
    assert random.getstate() == state


def test_eval_path_has_no_random_draws():
    cfg = PipelineCfg(is_train=False)
    seed_all(9)
    s0, s1, s2 = random.getstate(), np.random.get_state()[1].copy(), torch.get_rng_state().clone()
    p = draw_params(1000, cfg)
    assert random.getstate() == s0 and np.array_equal(np.random.get_state()[1], s1) and torch.equal(torch.get_rng_state(), s2)
    assert not p["cull"] and p["scale_x"] == 224 / 640 and p["scale_y"] == 224 / 480


def test_cull_and_flip_semantics():
    ev = np.array([[0.0, 0.0, 1.0, 1.0], [639.0, 479.0, 2.0, -1.0], [320.4, 100.7, 3.0, 1.0]])
    p = dict(scale_x=256 / 480, scale_y=256 / 480, start=0, count=3, time_flip=True, flip_x=True, flip_w=341, cull=True,
             shift_x=-1, shift_y=2, cull_w=341, cull_h=256)
    out = apply_event_aug(ev, p)
    # rows reversed, polarity inverted, x mirrored about W-1 then shifted; the mirrored first pixel (x=340-1) stays,
    # the mirrored last pixel (340 - 340.8 - 1 < 0) is culled
    assert out.shape == (2, 4)
    assert np.array_equal(out[:, 3], [-1.0, -1.0])
    assert np.array_equal(out[:, 2], [0.0, 2.0])
    assert out[1, 0] == 341 - 1 - 0.0 - 1 and out[1, 1] == 2.0


def test_product_draws_and_record_layout_match_the_oracle_and_the_header():
    """Host side of mem_b200.event_pipeline (no GPU needed): same generator consumption as the oracle, and the packed
    memb_event_aug records have the layout include/memb.h declares (64 bytes, field order / offsets)."""
    import ctypes
    import re
    from mem_b200 import event_pipeline as ep
    from oracle import event_pipeline_ref as ref
    for is_train in (True, False):
        for n in (100, 30000, 30001, 45000):
            seed_all(n + is_train)
            a = ep.draw_params(n, ep.PipelineConfig(is_train=is_train))
            seed_all(n + is_train)
            b = ref.draw_params(n, ref.PipelineCfg(is_train=is_train))
            assert a == b
    aug, crop = ep.pack_params([a, dict(a, start=7, count=9, time_flip=True, flip_x=True, shift_x=-3, shift_y=4, top=5, left=6)])
    assert aug.dtype.itemsize == ctypes.sizeof(ep.EventAug) == 64 and crop.tolist()[1] == [5, 6]
    rec = ep.EventAug.from_buffer_copy(aug[1].tobytes())
    assert (rec.start, rec.count, rec.time_flip, rec.flip_x, rec.shift_x, rec.shift_y) == (7, 9, 1, 1, -3, 4)
    assert rec.scale_x == a["scale_x"] and rec.flip_w == a["flip_w"] and rec.cull_w == a["cull_w"] and rec.cull_h == a["cull_h"]
    # field order in the header == field order of the ctypes structure
    hdr = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "memb.h")).read()
    body = hdr[hdr.index("typedef struct memb_event_aug"):hdr.index("} memb_event_aug;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = [n.strip() for decl in re.findall(r"(?:double|int64_t|int32_t)\s+([^;]+);", body) for n in decl.split(",")]
    assert names == [f[0] for f in ep.EventAug._fields_]
    with pytest.raises(AssertionError):
        ep.PipelineConfig(slice_max_evs=100)          # the reference's own range check (datasets.py:491)


def var_golden_cases(golden_dir):
    from oracle.event_pipeline_ref import VarPipelineCfg
    z = np.load(os.path.join(golden_dir, "event_pipeline_var.npz"))
    for name in sorted(k[:-4] for k in z.files if k.endswith("_out")):
        is_train, n, norm, seed, H, W, pol01 = (int(v) for v in z[name + "_meta"])
        kind = str(z[name + "_kind"])
        ev = np.floor(synth_events(np.random.default_rng(seed), n, H, W, kind, polarity=(0.0, 1.0) if pol01 else (-1.0, 1.0)))
        yield name, ev, VarPipelineCfg(is_train=bool(is_train), normalize_events=bool(norm)), seed, (H, W), z[name + "_out"]


def test_variable_sensor_oracle_matches_reference_golden(golden_dir):
    """The N-Caltech101 / N-Cars branch of build_transformNPY (H = W = None, per-sample sizes, antialiased bilinear
    Resize): tests/golden/event_pipeline_var.npz holds the reference chain's own outputs."""
    from oracle.event_pipeline_ref import pipeline_var_ref
    seen = 0
    for name, ev, cfg, seed, _, want in var_golden_cases(golden_dir):
        seed_all(seed)
        got = pipeline_var_ref(ev, cfg).numpy()
        assert got.shape == want.shape and got.dtype == np.float32, name
        assert np.array_equal(got, want), f"{name}: {np.abs(got - want).max()}"
        seen += 1
    assert seen == 6
    with pytest.raises(ValueError):                       # every row shifted out: the reference raises from max() of nothing
        pipeline_var_ref(np.array([[0.0, 0.0, 1.0, 1.0]]), cfg, dict(start=0, count=1, time_flip=False, flip_x=False, cull=True,
                                                                    shift_x=-5, shift_y=0))


def var_loggamma_cases(golden_dir):
    from oracle.event_pipeline_ref import VarPipelineCfg
    z = np.load(os.path.join(golden_dir, "event_pipeline_var_loggamma.npz"))
    for name in sorted(k[:-4] for k in z.files if k.endswith("_out")):
        is_train, n, norm, lg, gm, seed, H, W, pol01 = (int(v) for v in z[name + "_meta"])
        kind = str(z[name + "_kind"])
        ev = np.floor(synth_events(np.random.default_rng(seed), n, H, W, kind, polarity=(0.0, 1.0) if pol01 else (-1.0, 1.0)))
        cfg = VarPipelineCfg(is_train=bool(is_train), normalize_events=bool(norm), logtrafo=bool(lg), gammatrafo=bool(gm),
                             gamma=float(z[name + "_gamma"]))
        yield name, ev, cfg, seed, (H, W), z[name + "_out"]


def test_variable_sensor_log_gamma_oracle_matches_reference_golden(golden_dir):
    """LogTransform / GammaTransform on the variable-sensor branch act on the resized float32 image
    (tests/golden/event_pipeline_var_loggamma.npz: the reference chain's own outputs)."""
    from oracle.event_pipeline_ref import pipeline_var_ref
    seen = 0
    for name, ev, cfg, seed, _, want in var_loggamma_cases(golden_dir):
        seed_all(seed)
        got = pipeline_var_ref(ev, cfg).numpy()
        assert np.array_equal(got, want), f"{name}: {np.abs(got - want).max()}"
        seen += 1
    assert seen == 5


def var_tss_cases(golden_dir):
    from oracle.event_pipeline_ref import VarPipelineCfg
    z = np.load(os.path.join(golden_dir, "event_pipeline_var_tss.npz"))
    for name in sorted(k[:-4] for k in z.files if k.endswith("_out")):
        is_train, n, norm, lg, seed, H, W, pol01 = (int(v) for v in z[name + "_meta"])
        ev = np.floor(synth_events(np.random.default_rng(seed), n, H, W, str(z[name + "_kind"]),
                                   polarity=(0.0, 1.0) if pol01 else (-1.0, 1.0)))
        cfg = VarPipelineCfg(is_train=bool(is_train), normalize_events=bool(norm), logtrafo=bool(lg), timesurface=True)
        yield name, ev, cfg, seed, z[name + "_out"]


def test_variable_sensor_time_surface_oracle_matches_reference_golden(golden_dir):
    """args.timesurface on the variable-sensor branch: the time-surface plane is resized with the polarity planes and kept
    (tests/golden/event_pipeline_var_tss.npz: the reference chain's own outputs, both time-flip outcomes)."""
    from oracle.event_pipeline_ref import pipeline_var_ref
    seen = 0
    for name, ev, cfg, seed, want in var_tss_cases(golden_dir):
        seed_all(seed)
        got = pipeline_var_ref(ev, cfg).numpy()
        assert np.array_equal(got, want) and (got[1] != 0).any(), f"{name}: {np.abs(got - want).max()}"
        seen += 1
    assert seen == 6


def _loggamma_cases(golden_dir):
    z = np.load(os.path.join(golden_dir, "event_pipeline_loggamma.npz"))
    for name in sorted(k[:-4] for k in z.files if k.endswith("_out")):
        is_train, n, norm, lg, gm, seed = (int(v) for v in z[name + "_meta"])
        kind = str(z[name + "_kind"])
        ev = synth_events(np.random.default_rng(seed), n, 480, 640, kind, frac=(kind == "edge"))
        cfg = PipelineCfg(is_train=bool(is_train), normalize_events=bool(norm), logtrafo=bool(lg), gammatrafo=bool(gm),
                          gamma=float(z[name + "_gamma"]))
        yield name, ev, cfg, seed, z[name + "_out"]


def test_log_and_gamma_transforms_match_the_reference(golden_dir):
    """LogTransform / GammaTransform (transforms.py:200-222): the oracle chain against the reference's own outputs, and
    the product's route -- a 256-entry table of the transformed ``c / 255`` values evaluated with torch's CPU routines --
    against the oracle: identical float32 images, which is what lets the device path be bit-exact."""
    from mem_b200.event_pipeline import value_table
    seen = 0
    for name, ev, cfg, seed, want in _loggamma_cases(golden_dir):
        seed_all(seed)
        got = pipeline_ref(ev, cfg).numpy()
        assert np.array_equal(got, want), (name, float(np.abs(got - want).max()))
        # table route: the plain chain's image holds only c / 255 values (times the normalisation factor afterwards)
        plain = PipelineCfg(**{**cfg.__dict__, "logtrafo": False, "gammatrafo": False, "normalize_events": False})
        seed_all(seed)
        base = pipeline_ref(ev, plain)
        counts = torch.round(base * 255).long()
        assert torch.equal(counts.float() / 255, base)
        lut = value_table(cfg.logtrafo, cfg.gammatrafo, cfg.gamma)
        x = lut[counts]
        if cfg.normalize_events and x[0::2].max() != 0:
            x[0::2] = x[0::2] * (1.0 / x[0::2].max())
        assert np.array_equal(x.numpy(), want), name
        seen += 1
    assert seen == 4 and value_table(False, False) is None


def tss_loggamma_cases(golden_dir):
    z = np.load(os.path.join(golden_dir, "event_pipeline_tss_loggamma.npz"))
    for name in sorted(k[:-4] for k in z.files if k.endswith("_out")):
        is_train, n, norm, lg, gm, seed = (int(v) for v in z[name + "_meta"])
        kind = str(z[name + "_kind"])
        ev = synth_events(np.random.default_rng(seed), n, 480, 640, kind, frac=(kind == "edge"))
        cfg = PipelineCfg(is_train=bool(is_train), normalize_events=bool(norm), logtrafo=bool(lg), gammatrafo=bool(gm),
                          gamma=float(z[name + "_gamma"]), timesurface=True)
        yield name, ev, cfg, seed, z[name + "_out"]


def test_time_surface_with_log_gamma_matches_the_reference(golden_dir):
    """args.timesurface together with LogTransform / GammaTransform: the maps leave the time-surface plane alone."""
    seen = 0
    for name, ev, cfg, seed, want in tss_loggamma_cases(golden_dir):
        seed_all(seed)
        got = pipeline_ref(ev, cfg).numpy()
        assert np.array_equal(got, want) and (got[1] != 0).any(), (name, float(np.abs(got - want).max()))
        seen += 1
    assert seen == 4


def test_time_surface_chain_matches_the_reference(golden_dir):
    """args.timesurface=1: EventArrToImg(timeSurface=True) after the augmentations, middle channel kept
    (tests/golden/event_pipeline_tss.npz; two of the training cases draw RandomTimeFlip)."""
    z = np.load(os.path.join(golden_dir, "event_pipeline_tss.npz"))
    names = sorted(k[:-4] for k in z.files if k.endswith("_out"))
    assert len(names) == 6
    for name in names:
        is_train, n, norm, seed = (int(v) for v in z[name + "_meta"])
        kind = str(z[name + "_kind"])
        ev = synth_events(np.random.default_rng(seed), n, 480, 640, kind, frac=(kind == "edge"))
        seed_all(seed)
        got = pipeline_ref(ev, PipelineCfg(is_train=bool(is_train), normalize_events=bool(norm), timesurface=True)).numpy()
        assert np.array_equal(got, z[name + "_out"]), (name, float(np.abs(got - z[name + "_out"]).max()))
        assert (got[1] != 0).any()
