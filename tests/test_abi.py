"""CPU: the C-ABI library loads and exports every symbol include/memb.h declares
(no compute calls -- there is no GPU here)."""
import os
import re

from mem_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "memb.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(memb_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound(lib):
    names = _declared()
    assert "memb_hist_u8" in names and len(names) >= 7
    for n in names:
        assert hasattr(lib, n), f"{n} declared in memb.h but not exported by libmemb.so"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature in mem_b200/_lib.py"
    for n in _lib.SIGNATURES:
        assert n in names, f"{n} bound in _lib.py but not declared in memb.h"


def test_host_only_calls(lib):
    assert lib.memb_version() >= 100
    assert lib.memb_launch_count() >= 0
    # pure host arithmetic: workspace sizing
    tile = lib.memb_hist_workspace_bytes(128, 128 * 30000, 256, 341, 0, _lib.HIST_AUTO)
    glob = lib.memb_hist_workspace_bytes(1, 10_000_000, 480, 640, 0, _lib.HIST_AUTO)
    assert tile == 256 and glob >= 480 * 640 * 8
    assert lib.memb_hist_workspace_bytes(0, 0, 10, 10, 0, 0) == 0


def test_invalid_arguments_are_rejected_without_a_gpu(lib):
    # argument validation happens before any CUDA call
    rc = lib.memb_hist_u8(None, 0, None, 1, 0, 100, 100, 4, 0, 0, None, None, 0, None)
    assert rc == _lib.MEMB_EINVAL
    assert b"C must be 2 or 3" in lib.memb_last_error()


def test_product_fails_loudly_without_cuda():
    import pytest
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from mem_b200.process_data import histogram
    import numpy as np
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        histogram(np.zeros((1, 4)), 100, 100)
