"""CPU: raw-recording decoder oracle (oracle/decode_ref.py) against arrays the reference's own
process_data/process_dataset.py produced from synthetic recordings (tests/golden/decode.npz)."""
import os

import numpy as np
import pytest

from oracle import decode_ref


def golden(golden_dir):
    z = np.load(os.path.join(golden_dir, "decode.npz"))
    for k in sorted(f[:-4] for f in z.files if f.endswith("_raw")):
        yield k, z[k + "_raw"].tobytes(), z[k + "_npy"]


def test_oracle_matches_reference_golden(golden_dir):
    seen = 0
    for name, raw, want in golden(golden_dir):
        if name.startswith("ncaltech101"):
            loop, vec = decode_ref.ncaltech101_loop(raw), decode_ref.ncaltech101_np(raw)
        else:
            payload = decode_ref.skip_dat_header(raw)
            loop, vec = decode_ref.ncars_loop(payload), decode_ref.ncars_np(payload)
        assert want.dtype == np.float64 and loop.shape == want.shape, name
        assert np.array_equal(loop, want) and np.array_equal(vec, want), name
        seen += 1
    assert seen == 4


def test_vectorised_equals_loop_on_edge_fields():
    rng = np.random.default_rng(4)
    raw = bytearray(decode_ref.synth_ncaltech101(rng, 64))
    raw[0:5] = bytes([255, 255, 0xff, 0xff, 0xff])      # all fields at their maximum
    raw[5:10] = bytes([0, 0, 0x00, 0x00, 0x00])
    assert np.array_equal(decode_ref.ncaltech101_loop(bytes(raw)), decode_ref.ncaltech101_np(raw))
    got = decode_ref.ncaltech101_np(raw)[:2]
    assert got.tolist() == [[255.0, 255.0, float((1 << 23) - 1), 1.0], [0.0, 0.0, 0.0, -1.0]]
    pay = bytearray(decode_ref.synth_ncars(rng, 64, header=False))
    pay[0:8] = (0xffffffff).to_bytes(4, "little") * 2
    assert np.array_equal(decode_ref.ncars_loop(bytes(pay)), decode_ref.ncars_np(pay))
    assert decode_ref.ncars_np(pay)[0].tolist() == [16383.0, 16383.0, 4294967295.0, 1.0]


def test_file_readers_skip_header(tmp_path):
    from mem_b200.process_data import read_ncaltech101_bin, read_ncars_dat
    rng = np.random.default_rng(1)
    blob = decode_ref.synth_ncars(rng, 10)
    p = tmp_path / "a.dat"
    p.write_bytes(blob)
    assert read_ncars_dat(p).tobytes() == decode_ref.skip_dat_header(blob)
    q = tmp_path / "b.bin"
    q.write_bytes(decode_ref.synth_ncaltech101(rng, 7))
    assert read_ncaltech101_bin(q).size == 35


def test_nimagenet_npz_reader(tmp_path):
    """N-ImageNet recordings are .npz files whose ``event_data`` array the reference stores unchanged as .npy
    (process_dataset.py:108-117)."""
    from mem_b200.process_data import convert_nimagenet, read_nimagenet_npz
    rng = np.random.default_rng(0)
    ev = np.stack([rng.integers(0, 640, 500), rng.integers(0, 480, 500), np.sort(rng.integers(0, 10**6, 500)),
                   rng.integers(0, 2, 500)], axis=1).astype(np.int32)
    src = tmp_path / "n01440764_10026.npz"
    np.savez_compressed(src, event_data=ev)
    got = read_nimagenet_npz(str(src))
    assert got.dtype == np.int32 and np.array_equal(got, ev)
    dst = convert_nimagenet(str(src))
    assert dst.endswith("n01440764_10026.npy") and np.array_equal(np.load(dst), ev)
    np.savez(tmp_path / "bad.npz", other=ev)
    with pytest.raises(KeyError):
        read_nimagenet_npz(str(tmp_path / "bad.npz"))


@pytest.mark.reference
def test_nimagenet_conversion_matches_reference(tmp_path):
    """The unmodified reference ``nimagenet()`` on a temporary dataset tree writes the same .npy as convert_nimagenet."""
    from types import SimpleNamespace
    from mem_b200.process_data import convert_nimagenet
    from oracle import ref_shims
    ref_shims.install()
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_process_dataset", "/root/reference/process_data/process_dataset.py")
    mod = importlib.util.module_from_spec(spec)
    import sys
    # process_dataset.py does ``import utils`` (its own sibling); mem/ has a different ``utils`` that the other
    # live-reference tests import under the same name, so keep this one out of sys.modules afterwards
    saved = {k: sys.modules.get(k) for k in ("utils",)}
    for k in saved:
        sys.modules.pop(k, None)
    sys.path.insert(0, "/root/reference/process_data")
    try:
        spec.loader.exec_module(mod)
    finally:
        sys.path.remove("/root/reference/process_data")
        for k, v in saved.items():
            sys.modules.pop(k, None)
            if v is not None:
                sys.modules[k] = v
    rng = np.random.default_rng(1)
    ev = np.stack([rng.integers(0, 640, 300), rng.integers(0, 480, 300), np.sort(rng.uniform(0, 1e5, 300)),
                   rng.integers(0, 2, 300)], axis=1).astype(np.float64)
    for split in ("extracted_train", "extracted_val"):
        (tmp_path / "in" / split / "n0").mkdir(parents=True)
        np.savez_compressed(tmp_path / "in" / split / "n0" / "rec_1.npz", event_data=ev)
    mod.nimagenet("n0", SimpleNamespace(input=str(tmp_path / "in"), output=str(tmp_path / "out")))
    want = np.load(tmp_path / "out" / "train" / "n0" / "rec_1.npy")
    got = np.load(convert_nimagenet(str(tmp_path / "in" / "extracted_train" / "n0" / "rec_1.npz"), str(tmp_path / "ours.npy")))
    assert got.dtype == want.dtype and np.array_equal(got, want)
