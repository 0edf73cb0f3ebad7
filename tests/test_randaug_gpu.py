"""GPU parity: EventRandAugment on the device (csrc/randaug.cu through the C ABI, mem_b200/transforms.py) against the oracle
(bit-exact: the kernel and the oracle evaluate every float32 operation in the same order) and against outputs of the
unmodified reference module on torchvision (tests/golden/randaug.npz: bit-exact for the photometric operations, a few pixels
one count off for the bilinear resampling), plus the whole transform chain with the reference's default ``rand_aug=1``."""
import contextlib
import io
import os
import random

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import randaug_ref as R
from oracle.make_golden import synth_event_image, synth_events

GEOMETRIC = ("ShearX", "ShearY", "Rotate")


def _aug(**kw):
    from mem_b200 import transforms as T
    with contextlib.redirect_stdout(io.StringIO()):
        return T.EventRandAugment(**kw)


def test_every_operation_vs_oracle_and_reference(golden_dir):
    from mem_b200 import transforms as T
    z = np.load(os.path.join(golden_dir, "randaug.npz"))
    for key in "ab":
        img = z["img_" + key]
        tags = [str(t) for t in z["cases"] if str(t).split("_")[1] == key]
        ops = np.zeros((len(tags), 1), dtype=T.OP_DTYPE)
        for i, tag in enumerate(tags):
            ops[i, 0] = T.encode_op(tag.split("_")[2], float(z[tag + "_mag"]))
        batch = torch.from_numpy(np.broadcast_to(img, (len(tags),) + img.shape).copy()).cuda()
        got = T.apply_ops(batch, ops).cpu().numpy()
        for i, tag in enumerate(tags):
            name = tag.split("_")[2]
            want_oracle = R.apply_op(img, name, float(z[tag + "_mag"]))
            assert np.array_equal(got[i], want_oracle), (tag, int((got[i] != want_oracle).sum()))
            diff = np.abs(got[i].astype(np.int32) - z[tag].astype(np.int32))
            if name in GEOMETRIC:
                assert diff.max() <= 1 and int((diff != 0).sum()) <= 4, tag
            else:
                assert not diff.any(), tag


def test_module_under_fixed_seeds_vs_reference(golden_dir):
    z = np.load(os.path.join(golden_dir, "randaug.npz"))
    aug = _aug(small=False, magnitude=20)
    for key in "ab":
        img = torch.from_numpy(z["img_" + key]).cuda()
        for seed in range(12):
            if f"full_{key}_{seed}" not in z.files:
                continue
            torch.manual_seed(1000 + seed)
            got = aug(img)
            assert got.dtype == torch.uint8 and got.shape == img.shape
            diff = np.abs(got.cpu().numpy().astype(np.int32) - z[f"full_{key}_{seed}"].astype(np.int32))
            assert diff.max() <= 1 and int((diff != 0).sum()) <= 8, (key, seed)


def test_mixed_batch_two_operations_float_in_float_out():
    """augment_batch = ToUnit8 -> EventRandAugment -> ToFloat32 in one launch: every sample with its own pair of
    operations (all 14 x 14 ordered pairs over the batch), float32 counts / 255 in, float32 out, against the oracle."""
    from mem_b200 import transforms as T
    rng = np.random.default_rng(3)
    H, W = 96, 112
    pairs = [(a, b) for a in R.OPS for b in R.OPS]
    imgs = np.stack([synth_event_image(rng, H, W) for _ in range(len(pairs))])
    x = imgs.astype(np.float32) / np.float32(255)
    x[::3] = x[::3] / np.maximum(x[::3].max(axis=(1, 2, 3), keepdims=True), np.float32(1e-6))     # NormalizeEvent-like inputs
    space = R.augmentation_space(R.OPS, 31, H, W)
    ops = np.zeros((len(pairs), 2), dtype=T.OP_DTYPE)
    chosen = []
    for i, pair in enumerate(pairs):
        row = []
        for k, name in enumerate(pair):
            mags, signed = space[name]
            mag = float(mags[int(rng.integers(0, 21))]) if mags is not None else 0.0
            if signed and rng.integers(0, 2):
                mag = -mag
            ops[i, k] = T.encode_op(name, mag)
            row.append((name, mag))
        chosen.append(row)
    got = T.apply_ops(torch.from_numpy(x).cuda(), ops, out_float=True)
    assert got.dtype == torch.float32
    got = got.cpu().numpy()
    for i, row in enumerate(chosen):
        want = R.to_float32(R.rand_augment(R.to_uint8(x[i]), row))
        assert np.array_equal(got[i], want), (row, int((got[i] != want).sum()))
    # uint8 in place
    u = torch.from_numpy(imgs).cuda()
    again = T.apply_ops(u, ops, out_float=False).cpu().numpy()
    for i in (0, 17, 100, 195):
        assert np.array_equal(again[i], R.rand_augment(imgs[i], chosen[i]))


def test_random_shapes_and_operation_counts_vs_oracle():
    """Odd image sizes (the scalar load / store path: 3 * H * W not a multiple of 4), zero to three operations per sample,
    tiny images (Sharpness leaves images with a side <= 2 alone), a constant image (AutoContrast's non-finite scale, Equalize's
    zero step): the kernel equals the oracle bit for bit."""
    from mem_b200 import transforms as T
    rng = np.random.default_rng(8)
    for H, W, num_ops in ((37, 53, 3), (2, 41, 2), (64, 64, 0), (101, 7, 1), (224, 224, 3)):
        B = 6
        imgs = np.stack([synth_event_image(rng, H, W) for _ in range(B)])
        imgs[1] = 77                                               # constant image
        space = R.augmentation_space(R.OPS, 31, H, W)
        ops = np.zeros((B, num_ops), dtype=T.OP_DTYPE)
        chosen = []
        for b in range(B):
            row = []
            for k in range(num_ops):
                name = R.OPS[int(rng.integers(0, len(R.OPS)))]
                mags, signed = space[name]
                mag = float(mags[int(rng.integers(0, 31))]) if mags is not None else 0.0
                if signed and rng.integers(0, 2):
                    mag = -mag
                ops[b, k] = T.encode_op(name, mag)
                row.append((name, mag))
            chosen.append(row)
        got = T.apply_ops(torch.from_numpy(imgs).cuda(), ops).cpu().numpy()
        for b in range(B):
            want = R.rand_augment(imgs[b], chosen[b])
            assert np.array_equal(got[b], want), (H, W, chosen[b], int((got[b] != want).sum()))


def test_whole_chain_with_rand_aug_vs_reference(golden_dir):
    """The reference's build_transformNPY(is_train=True, rand_aug=1) outputs under fixed seeds, fixed-sensor and
    variable-sensor branch, against EventBatchPipeline / EventBatchPipelineVar with ``rand_aug=True``."""
    from mem_b200.event_pipeline import EventBatchPipeline, EventBatchPipelineVar, PipelineConfig, VarPipelineConfig
    z = np.load(os.path.join(golden_dir, "event_pipeline_randaug.npz"))
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
            got = EventBatchPipeline(PipelineConfig(is_train=True, normalize_events=bool(norm), rand_aug=True))([ev])
        else:
            got = EventBatchPipelineVar(VarPipelineConfig(is_train=True, canvas_H=180, canvas_W=240, normalize_events=bool(norm),
                                                          rand_aug=True))([ev])
        assert got.dtype == torch.float32 and tuple(got.shape) == (1, 3, 224, 224)
        got_u8 = np.rint(got[0].cpu().numpy() * 255).astype(np.int32)
        assert np.array_equal(got_u8.astype(np.float32) / np.float32(255), got[0].cpu().numpy())       # exactly k / 255
        diff = np.abs(got_u8 - z[name + "_out"].astype(np.int32))
        assert diff.max() <= 1 and int((diff != 0).sum()) <= 8, (name, int((diff != 0).sum()), int(diff.max()))


def test_bad_arguments_fail_loudly():
    from mem_b200 import transforms as T
    ops = np.zeros((1, 1), dtype=T.OP_DTYPE)
    with pytest.raises(AssertionError):
        T.apply_ops(torch.zeros(1, 2, 32, 32, dtype=torch.uint8, device="cuda"), ops)
    with pytest.raises(ValueError):
        T.apply_ops(torch.zeros(1, 3, 300, 300, dtype=torch.uint8, device="cuda"), ops)       # 270 KB: no shared-memory tile
    with pytest.raises(ValueError):
        T.encode_op("Hue", 0.1)
    with pytest.raises(NotImplementedError):
        T.EventRandAugment(fill=[0.0])
