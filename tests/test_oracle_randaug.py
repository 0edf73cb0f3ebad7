"""CPU: the EventRandAugment oracle (oracle/randaug_ref.py) against outputs of the unmodified reference module running on
torchvision (tests/golden/randaug.npz), the host-side draws of the drop-in module against the oracle's, and the reference's
whole transform chain with its default ``rand_aug=1`` tail (tests/golden/event_pipeline_randaug.npz)."""
import contextlib
import io
import os
import random

import numpy as np
import pytest
import torch

from oracle import event_pipeline_ref as E
from oracle import randaug_ref as R
from oracle.make_golden import synth_events

GEOMETRIC = ("ShearX", "ShearY", "Rotate")       # bilinear resampling: up to a few pixels one count off (see the oracle's header)


def _z(golden_dir, name="randaug.npz"):
    return np.load(os.path.join(golden_dir, name))


def test_every_operation_matches_the_reference(golden_dir):
    z = _z(golden_dir)
    assert len(z["cases"]) >= 100
    seen = set()
    for tag in (str(t) for t in z["cases"]):
        _, key, name, _, _ = tag.split("_")
        seen.add(name)
        got, want = R.apply_op(z["img_" + key], name, float(z[tag + "_mag"])), z[tag]
        diff = np.abs(got.astype(np.int32) - want.astype(np.int32))
        if name in GEOMETRIC:
            assert diff.max() <= 1 and int((diff != 0).sum()) <= 4, (tag, int((diff != 0).sum()))
        else:
            assert not diff.any(), (tag, int((diff != 0).sum()))
    assert seen == set(R.OPS)


def test_module_under_fixed_seeds_matches_the_reference(golden_dir):
    """Pins the draw order (operation index, magnitude bin, sign: three torch.randint calls per operation)."""
    z = _z(golden_dir)
    for key in "ab":
        img = z["img_" + key]
        for seed in range(12):
            if f"full_{key}_{seed}" not in z.files:
                continue
            torch.manual_seed(1000 + seed)
            ops = R.draw_ops(lambda n: int(torch.randint(n, (1,)).item()), R.OPS, 2, 20, 31, img.shape[1], img.shape[2])
            diff = np.abs(R.rand_augment(img, ops).astype(np.int32) - z[f"full_{key}_{seed}"].astype(np.int32))
            geometric = any(n in GEOMETRIC for n, _ in ops)
            assert diff.max() <= (1 if geometric else 0) and int((diff != 0).sum()) <= (8 if geometric else 0), (key, seed, ops)


def test_to_uint8_to_float32(golden_dir):
    z = _z(golden_dir)
    assert np.array_equal(R.to_uint8(z["tou8_in"]), z["tou8_out"])
    assert np.array_equal(R.to_uint8(z["tou8n_in"]), z["tou8n_out"])
    assert np.array_equal(R.to_float32(z["img_a"]), z["tof32_out"])


def test_dropin_module_draws_like_the_oracle_and_encodes_the_header_layout():
    from mem_b200 import transforms as T
    with contextlib.redirect_stdout(io.StringIO()):
        aug = T.EventRandAugment(small=False, magnitude=20)
        small = T.EventRandAugment(small=True, magnitude=9)
    assert aug.names == R.OPS and small.names == R.SMALL
    for seed in range(20):
        torch.manual_seed(seed)
        mine = aug.draw(224, 224)
        torch.manual_seed(seed)
        ref = R.draw_ops(lambda n: int(torch.randint(n, (1,)).item()), R.OPS, 2, 20, 31, 224, 224)
        assert mine == ref
    assert T.OP_DTYPE.itemsize == 40 and T.OP_DTYPE.fields["theta"][1] == 16        # memb_randaug_op (include/memb.h)
    rec = T.encode_op("Rotate", -14.0)
    assert rec["op"] == T.RA_AFFINE and np.allclose(rec["theta"], np.asarray(R.op_matrix("Rotate", -14.0), dtype=np.float32), rtol=0, atol=0)
    rec = T.encode_op("Brightness", -0.27)
    assert rec["f0"] == np.float32(1.0 - 0.27) and rec["f1"] == np.float32(1.0 - (1.0 - 0.27))
    assert T.encode_op("Posterize", 6.0)["ival"] == 6 and T.encode_op("Solarize", 246.5)["f0"] == np.float32(246.5)
    with pytest.raises(RuntimeError):
        aug(torch.zeros(3, 8, 8, dtype=torch.uint8))             # CPU tensor: no fallback path


def test_whole_chain_with_rand_aug_matches_the_reference(golden_dir):
    """build_transformNPY(is_train=True, rand_aug=1) of the reference vs the oracle chain under the same seeds."""
    z = _z(golden_dir, "event_pipeline_randaug.npz")
    names = sorted(k[:-4] for k in z.files if k.endswith("_out"))
    assert len(names) == 6
    for name in names:
        n, norm, seed, H, W, pol01, fixed = (int(v) for v in z[name + "_meta"])
        kind = str(z[name + "_kind"])
        pol = (0.0, 1.0) if pol01 else (-1.0, 1.0)
        ev = synth_events(np.random.default_rng(seed), n, H, W, kind, polarity=pol, frac=bool(fixed and kind == "edge"))
        if not fixed:
            ev = np.floor(ev)
        random.seed(seed); np.random.seed(seed); torch.manual_seed(seed)
        if fixed:
            got = E.pipeline_ref(ev, E.PipelineCfg(is_train=True, normalize_events=bool(norm), rand_aug=True))
        else:
            got = E.pipeline_var_ref(ev, E.VarPipelineCfg(is_train=True, normalize_events=bool(norm), rand_aug=True))
        got_u8 = np.rint(got.numpy() * 255).astype(np.int32)
        diff = np.abs(got_u8 - z[name + "_out"].astype(np.int32))
        assert diff.max() <= 1 and int((diff != 0).sum()) <= 8, (name, int((diff != 0).sum()), int(diff.max()))
