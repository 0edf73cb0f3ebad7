"""GPU parity: raw-recording decoders and the raw rasteriser (through the C ABI), bit-exact against the reference's
own decoder outputs (tests/golden/decode.npz) and the oracle at larger sizes."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import decode_ref
from oracle.histogram_ref import event_hist_ref


def golden(golden_dir):
    z = np.load(os.path.join(golden_dir, "decode.npz"))
    for k in sorted(f[:-4] for f in z.files if f.endswith("_raw")):
        yield k, z[k + "_raw"].tobytes(), z[k + "_npy"]


def test_reference_golden(golden_dir):
    from mem_b200.process_data import RAW_NCALTECH101, RAW_NCARS, decode_events
    for name, raw, want in golden(golden_dir):
        if name.startswith("ncaltech101"):
            got = decode_events(np.frombuffer(raw, np.uint8), RAW_NCALTECH101)
        else:
            got = decode_events(np.frombuffer(decode_ref.skip_dat_header(raw), np.uint8), RAW_NCARS)
        assert got.dtype == np.float64 and got.shape == want.shape, name
        assert np.array_equal(got, want), name


@pytest.mark.parametrize("n", [0, 1, 3, 4095, 4096, 4097, 100_003, 3_000_001])
def test_decoders_vs_oracle(n):
    from mem_b200.process_data import RAW_NCALTECH101, RAW_NCARS, decode_events
    rng = np.random.default_rng(n)
    raw = decode_ref.synth_ncaltech101(rng, n)
    got = decode_events(torch.frombuffer(bytearray(raw), dtype=torch.uint8).cuda() if n else torch.zeros(0, dtype=torch.uint8).cuda(),
                        RAW_NCALTECH101)
    assert got.is_cuda and tuple(got.shape) == (n, 4)
    assert np.array_equal(got.cpu().numpy(), decode_ref.ncaltech101_np(raw))
    pay = decode_ref.synth_ncars(rng, n, header=False)
    got = decode_events(np.frombuffer(pay, np.uint8), RAW_NCARS)
    assert np.array_equal(got, decode_ref.ncars_np(pay))


def test_unaligned_slice_is_handled():
    from mem_b200.process_data import RAW_NCALTECH101, decode_events
    rng = np.random.default_rng(9)
    raw = decode_ref.synth_ncaltech101(rng, 1000)
    buf = torch.frombuffer(bytearray(b"\x00" * 7 + raw), dtype=torch.uint8).cuda()
    got = decode_events(buf[7:], RAW_NCALTECH101)
    assert np.array_equal(got.cpu().numpy(), decode_ref.ncaltech101_np(raw))


@pytest.mark.parametrize("n", [1, 5000, 2_000_003])
def test_raw_rasteriser_equals_decode_then_rasterise(n):
    from mem_b200.process_data import RAW_NCALTECH101, RAW_NCARS, histogram, histogram_raw
    rng = np.random.default_rng(n + 1)
    raw = decode_ref.synth_ncaltech101(rng, n, W=240, H=180)
    want = event_hist_ref(decode_ref.ncaltech101_np(raw), 180, 240)
    got = histogram_raw(np.frombuffer(raw, np.uint8), RAW_NCALTECH101, 180, 240)
    assert np.array_equal(got, want)
    assert np.array_equal(histogram_raw(np.frombuffer(raw, np.uint8), RAW_NCALTECH101, 180, 240, channels=2), want[..., 0::2])
    # N-Cars polarity is {0,1}: only the +1 rows count, like the reference's EventArrToImg on its .npy files
    pay = decode_ref.synth_ncars(rng, n, W=120, H=100, header=False)
    ev = decode_ref.ncars_np(pay)
    want = event_hist_ref(ev, 100, 120)
    got = histogram_raw(np.frombuffer(pay, np.uint8), RAW_NCARS, 100, 120)
    assert np.array_equal(got, want) and want[..., 2].sum() == 0
    assert np.array_equal(histogram(ev, 100, 120), want)


def test_out_of_range_and_bad_sizes_raise():
    from mem_b200.process_data import RAW_NCALTECH101, decode_events, histogram_raw
    rng = np.random.default_rng(2)
    raw = np.frombuffer(decode_ref.synth_ncaltech101(rng, 100, W=240, H=180), np.uint8)
    with pytest.raises(IndexError):
        histogram_raw(raw, RAW_NCALTECH101, 100, 100)       # coordinates up to 239 / 179 on a 100 x 100 sensor
    with pytest.raises(ValueError):
        decode_events(raw[:-1], RAW_NCALTECH101)
    with pytest.raises(ValueError):
        decode_events(raw, 7)
