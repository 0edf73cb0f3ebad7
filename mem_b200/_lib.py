"""ctypes binding of ``libmemb.so`` (the C ABI declared in ``include/memb.h``).

PyTorch is plumbing here: it owns device memory and streams, and this module
hands raw pointers to the library.  There is no CPU or PyTorch fallback: if the
shared library is missing or CUDA is unavailable, calls fail loudly.
"""
from __future__ import annotations

import ctypes
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MEMB_LIB_PATH") or os.path.join(_HERE, "libmemb.so")  # override: debugging builds only

MEMB_OK, MEMB_EINVAL, MEMB_EOOB, MEMB_ECUDA, MEMB_EWORKSPACE = 0, -1, -2, -3, -4

HIST_AUTO, HIST_GLOBAL, HIST_GLOBAL_AGG, HIST_TILE, HIST_PRIVATE, HIST_GLOBAL_REPL, HIST_HYBRID, HIST_SORT = 0, 1, 2, 3, 4, 5, 6, 7
RAW_NCALTECH101, RAW_NCARS = 1, 2

_c = ctypes
_vp, _i32, _i64, _sz, _f32 = _c.c_void_p, _c.c_int, _c.c_int64, _c.c_size_t, _c.c_float

# name -> (restype, argtypes).  Kept in step with include/memb.h; tests/test_abi.py
# checks that every symbol the header declares is exported and listed here.
SIGNATURES = {
    "memb_last_error": (_c.c_char_p, []),
    "memb_version": (_i32, []),
    "memb_launch_count": (_i64, []),
    "memb_hist_workspace_bytes": (_sz, [_i32, _i64, _i32, _i32, _i32, _i32]),
    "memb_hist_u8": (_i32, [_vp, _i64, _vp, _i32, _i64, _i32, _i32, _i32, _i32, _i32, _vp, _vp, _sz, _vp]),
    "memb_hist_status": (_i32, [_vp, _vp]),
    "memb_hist_extent": (_i32, [_vp, _i64, _vp, _vp, _sz, _vp]),
    "memb_decode_events_f64": (_i32, [_vp, _i64, _i32, _vp, _vp]),
    "memb_hist_raw_u8": (_i32, [_vp, _i64, _i32, _i32, _i32, _i32, _vp, _vp, _sz, _vp]),
    "memb_hist_aug_u8": (_i32, [_vp, _i64, _vp, _i32, _i64, _vp, _i32, _i32, _i32, _i32, _vp, _vp, _sz, _vp]),
    "memb_hist_aug_tss_u8": (_i32, [_vp, _i64, _vp, _i32, _i64, _vp, _i32, _i32, _i32, _vp, _vp, _sz, _vp]),
    "memb_event_pipeline_f32": (_i32, [_vp, _i64, _vp, _i32, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _f32, _i32,
                                       _vp, _vp, _sz, _vp]),
    "memb_event_pipeline_lut_f32": (_i32, [_vp, _i64, _vp, _i32, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _f32, _i32,
                                           _vp, _vp, _vp, _sz, _vp]),
    "memb_event_pipeline_var_f32": (_i32, [_vp, _i64, _vp, _i32, _vp, _i32, _i32, _i32, _i32, _i32, _f32, _i32, _vp, _vp, _sz, _vp]),
    "memb_event_pipeline_var_tf_f32": (_i32, [_vp, _i64, _vp, _i32, _vp, _i32, _i32, _i32, _i32, _i32, _f32, _i32, _i32, _i32, _f32,
                                              _i32, _vp, _vp, _sz, _vp]),
    "memb_event_randaug": (_i32, [_vp, _i32, _i32, _i32, _i32, _i32, _vp, _i32, _vp, _i32, _vp]),
    "memb_raster_post_workspace_bytes": (_sz, [_i32]),
    "memb_raster_post_f32": (_i32, [_vp, _i32, _i32, _i32, _i32, _vp, _i32, _i32, _i32, _i32, _i32, _f32, _i32, _vp, _vp,
                                    _sz, _vp]),
    "memb_raster_post_lut_f32": (_i32, [_vp, _i32, _i32, _i32, _i32, _vp, _i32, _i32, _i32, _i32, _i32, _f32, _i32, _vp, _vp, _vp,
                                        _sz, _vp]),
}

DT_BF16, DT_F32 = 0, 1
ATTN_BIAS_FLOATS_PER_HEAD = 52 * 256 * 4  # MEMB_ATTN_BIAS_FLOATS_PER_HEAD
EPI_STORE, EPI_BIAS_GELU, EPI_RESIDUAL, EPI_ATOMIC_ADD, EPI_DGELU, EPI_ARGMAX, EPI_STORE_ROWDOT = 0, 1, 2, 3, 4, 5, 6


class GemmDesc(_c.Structure):
    """Mirror of ``memb_gemm_desc`` (include/memb.h)."""
    _fields_ = [
        ("a", _vp), ("b", _vp), ("lda", _i64), ("ldb", _i64),
        ("m", _i32), ("n", _i32), ("k", _i32),
        ("a_layout", _i32), ("b_layout", _i32), ("in_dtype", _i32), ("out_dtype", _i32),
        ("epilogue", _i32), ("splits", _i32), ("block_n", _i32), ("split_precision", _i32),
        ("act", _i32), ("out_split", _i32),
        ("d", _vp), ("ldd", _i64), ("d2", _vp), ("ldd2", _i64),
        ("bias", _vp), ("aux", _vp), ("ldaux", _i64), ("colscale", _vp), ("rowscale", _vp),
        ("rows_per_group", _i32), ("out_group_rows", _i32), ("out_group_stride", _i32), ("out_row_offset", _i32),
        ("rowmask", _vp), ("maskvec", _vp), ("alpha", _f32), ("alpha_dev", _vp), ("err_flag", _vp), ("colsum", _vp), ("rowdot", _vp),
    ]


SIGNATURES["memb_gemm"] = (_i32, [_c.POINTER(GemmDesc), _vp])
SIGNATURES.update({
    "memb_layernorm_fwd": (_i32, [_vp, _i64, _vp, _vp, _f32, _i32, _i32, _vp, _i64, _vp, _vp, _vp, _vp, _vp]),
    "memb_layernorm_bwd": (_i32, [_vp, _i32, _i64, _vp, _i64, _vp, _vp, _vp, _i32, _i32, _vp, _i64, _vp, _vp, _vp, _vp, _vp]),
    "memb_layernorm_bwd_branch": (_i32, [_vp, _i32, _i64, _vp, _i64, _vp, _vp, _vp, _i32, _i32, _vp, _i64, _vp, _vp,
                                         _vp, _i64, _vp, _vp, _i32, _vp, _i64, _vp, _vp, _vp]),
    "memb_branch_bwd": (_i32, [_vp, _i64, _vp, _i64, _vp, _vp, _i32, _i32, _i32, _vp, _i64, _vp, _vp, _vp]),
    "memb_colsum_bf16": (_i32, [_vp, _i64, _i32, _i32, _vp, _vp]),
    "memb_vbias_chain": (_i32, [_vp, _vp, _i32, _vp, _vp, _vp]),
    "memb_patchify": (_i32, [_vp, _i32, _i32, _i32, _i32, _i32, _vp, _vp]),
    "memb_cls_pos": (_i32, [_vp, _vp, _vp, _i32, _i32, _i32, _vp]),
    "memb_embed_bwd": (_i32, [_vp, _vp, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp]),
    "memb_mask_compact": (_i32, [_vp, _i32, _i32, _vp, _vp, _vp, _i32, _vp]),
    "memb_cross_entropy": (_i32, [_vp, _i64, _vp, _vp, _vp, _i32, _i32, _vp, _i64, _vp, _f32, _vp]),
    "memb_relpos_gather": (_i32, [_vp, _vp, _i32, _i32, _i32, _vp, _vp, _vp]),
    "memb_relpos_scatter": (_i32, [_vp, _i32, _vp, _i32, _i32, _vp, _vp]),
    "memb_batch_reduce_bf16": (_i32, [_vp, _i32, _i64, _vp, _vp]),
    "memb_meanpool_fwd": (_i32, [_vp, _i32, _i32, _i32, _vp, _vp]),
    "memb_meanpool_bwd": (_i32, [_vp, _i32, _i32, _i32, _vp, _vp]),
    "memb_linear_small_fwd": (_i32, [_vp, _vp, _vp, _i32, _i32, _i32, _vp, _vp]),
    "memb_linear_small_bwd": (_i32, [_vp, _vp, _vp, _i32, _i32, _i32, _vp, _vp, _vp, _vp]),
    "memb_attention_pack_bias": (_i32, [_vp, _i32, _i32, _i32, _vp, _vp]),
    "memb_attention_fwd": (_i32, [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _f32, _vp, _vp, _vp]),
    "memb_attention_bwd_workspace_bytes": (_sz, [_i32, _i32, _i32]),
    "memb_rowdot_heads": (_i32, [_vp, _vp, _i32, _i32, _i32, _vp, _vp]),
    "memb_attention_bwd": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _f32, _vp, _vp, _vp, _sz, _vp]),
    "memb_fill_f32": (_i32, [_vp, _i64, _f32, _vp]),
    "memb_cast_bf16": (_i32, [_vp, _vp, _i64, _vp]),
    "memb_sqnorm": (_i32, [_vp, _i64, _f32, _vp, _vp]),
    "memb_sqnorm_groups": (_i32, [_vp, _i64, _f32, _vp, _vp, _vp]),
    "memb_soft_ce": (_i32, [_vp, _i32, _i32, _vp, _vp, _f32, _vp, _vp, _vp]),
    # dVAE training step / decoder (csrc/vae_train.cu)
    "memb_vae_nchw_to_nhwc": (_i32, [_vp, _i32, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp]),
    "memb_vae_nhwc_to_nchw": (_i32, [_vp, _i32, _i32, _i32, _i32, _i32, _vp, _vp]),
    "memb_vae_im2col": (_i32, [_vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _vp]),
    "memb_vae_col2im": (_i32, [_vp, _i64, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _i32, _vp, _vp, _vp, _i32, _vp]),
    "memb_vae_ew_bf16": (_i32, [_i32, _vp, _vp, _vp, _i64, _vp]),
    "memb_vae_gather_rows": (_i32, [_vp, _i32, _i32, _vp, _i64, _i32, _vp, _vp, _vp]),
    "memb_vae_gumbel_fwd": (_i32, [_vp, _vp, _i64, _i32, _f32, _i32, _vp, _vp, _vp, _vp, _vp]),
    "memb_vae_gumbel_bwd": (_i32, [_vp, _vp, _vp, _vp, _i64, _i32, _f32, _vp, _f32, _vp, _vp]),
    "memb_vae_recon_loss": (_i32, [_vp, _vp, _i64, _i32, _i32, _i32, _vp, _vp, _vp]),
    "memb_axpy_f32": (_i32, [_vp, _vp, _vp, _i64, _i32, _vp]),
    "memb_adamw": (_i32, [_vp, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _i32, _f32, _f32, _f32, _i32, _f32, _f32, _vp, _vp]),
})



class ConvDesc(_c.Structure):
    """Mirror of ``memb_conv_desc`` (include/memb.h)."""
    _fields_ = [
        ("a_hi", _vp), ("a_lo", _vp), ("inner", _i32), ("x_slots", _i32), ("r_slots", _i32), ("rows_per_img", _i32),
        ("taps_y", _i32), ("taps_x", _i32), ("tap_y0", _i32), ("tap_x0", _i32),
        ("w", _vp), ("bias", _vp), ("B", _i32), ("OH", _i32), ("OW", _i32), ("Cout", _i32), ("relu", _i32),
        ("aux", _vp), ("d_full", _vp), ("keys", _vp), ("seg_kblocks", _i32), ("d_hi", _vp), ("d_lo", _vp),
        ("sB", _i64), ("sy_major", _i64), ("sy_minor", _i64), ("sx_major", _i64), ("sx_minor", _i64),
        ("pad", _i32), ("shift", _i32), ("err_flag", _vp),
    ]


class Conv16Desc(_c.Structure):
    """Mirror of ``memb_conv16_desc`` (include/memb.h)."""
    _fields_ = [
        ("a_hi", _vp), ("a_lo", _vp), ("inner", _i32), ("x_slots", _i32), ("r_slots", _i32), ("rows_per_img", _i32),
        ("taps_y", _i32), ("taps_x", _i32), ("tap_y0", _i32), ("tap_x0", _i32),
        ("w", _vp), ("bias", _vp), ("B", _i32), ("OH", _i32), ("OW", _i32), ("Cout", _i32), ("relu", _i32),
        ("a_exp", _i32), ("w_exp", _i32), ("out_exp", _i32),
        ("aux", _vp), ("d_full", _vp), ("keys", _vp), ("seg_kblocks", _i32), ("d_hi", _vp), ("d_lo", _vp),
        ("sB", _i64), ("sy_major", _i64), ("sy_minor", _i64), ("sx_major", _i64), ("sx_minor", _i64),
        ("pad", _i32), ("shift", _i32), ("absmax", _vp), ("err_flag", _vp),
    ]


SIGNATURES.update({
    "memb_conv_f16x2": (_i32, [_c.POINTER(Conv16Desc), _vp]),
    "memb_dvae_im2col_l1_f16": (_i32, [_vp, _i32, _i32, _i32, _i32, _i32, _vp, _vp, _i32, _vp, _vp, _vp, _vp]),
    "memb_split_f16": (_i32, [_vp, _i32, _vp, _vp, _i64, _vp]),
    "memb_conv_tf32x3": (_i32, [_c.POINTER(ConvDesc), _vp]),
    "memb_dvae_im2col_l1": (_i32, [_vp, _i32, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp]),
    "memb_split_tf32": (_i32, [_vp, _vp, _vp, _i64, _vp]),
    "memb_argmax_decode": (_i32, [_vp, _vp, _i64, _vp]),
})

_lib = None
_lock = threading.Lock()


class MembError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libmemb error {code}: {msg}")
        self.code = code


def load() -> ctypes.CDLL:
    """Load the library (once).  Raises if it has not been built -- run
    ``python -c 'import __graft_entry__ as g; g.build()'`` at the repo root."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise RuntimeError(
                    f"{LIB_PATH} is missing: the CUDA library has not been built and mem_b200 has "
                    "no CPU fallback (build it with __graft_entry__.build()).")
            lib = ctypes.CDLL(LIB_PATH)
            for name, (res, args) in SIGNATURES.items():
                fn = getattr(lib, name)
                fn.restype, fn.argtypes = res, args
            _lib = lib
    return _lib


def check(code: int) -> None:
    if code == MEMB_OK:
        return
    msg = load().memb_last_error().decode("utf-8", "replace")
    if code == MEMB_EOOB:
        raise IndexError(msg)
    if code == MEMB_EINVAL:
        raise ValueError(msg)
    raise MembError(code, msg)


def launch_count() -> int:
    return int(load().memb_launch_count())


def require_cuda():
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("mem_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    return torch


def stream_ptr(torch, device=None) -> int:
    return int(torch.cuda.current_stream(device).cuda_stream)


class Workspace:
    """Grow-only device scratch buffer owned by the caller side (PyTorch caching allocator)."""

    def __init__(self):
        self._buf = {}

    def get(self, torch, nbytes: int, device, tag: str = "default"):
        key = (tag, str(device))
        buf = self._buf.get(key)
        if buf is None or buf.numel() < nbytes:
            buf = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)
            self._buf[key] = buf
        return buf


workspace = Workspace()
