"""CPU: raw-recording decoder oracle (oracle/decode_ref.py) against arrays the reference's own
process_data/process_dataset.py produced from synthetic recordings (tests/golden/decode.npz)."""
import os

import numpy as np

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
