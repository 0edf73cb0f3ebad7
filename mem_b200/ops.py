"""Thin Python wrappers over the C ABI: tensors in, raw pointers out.

Nothing here computes on the host or falls back to PyTorch kernels; each function
validates shapes, builds the descriptor and calls ``libmemb``.
"""
from __future__ import annotations

import ctypes

from . import _lib
from ._lib import (DT_BF16, DT_F32, EPI_ARGMAX, EPI_ATOMIC_ADD, EPI_BIAS_GELU, EPI_DGELU, EPI_RESIDUAL,
                   EPI_STORE, GemmDesc)

_err_flags = {}


def _err_flag(torch, device):
    key = str(device)
    f = _err_flags.get(key)
    if f is None:
        f = torch.zeros(1, dtype=torch.int32, device=device)
        _err_flags[key] = f
    return f


def _ptr(t):
    return None if t is None else t.data_ptr()


def gemm(a, b, *, out=None, out_dtype=None, a_layout=0, b_layout=0, epilogue=EPI_STORE, bias=None,
         aux=None, d2=None, colscale=None, rowscale=None, rows_per_group=0, alpha=1.0, act=0, splits=0,
         block_n=0, split_precision=False, out_split=False, row_remap=None, rowmask=None, maskvec=None,
         m=None, n=None, k=None, alpha_dev=None, colsum=None, rowdot=None):
    """``D[M,N] = epilogue(A @ B^T)`` on the tcgen05 GEMM.

    a: ``[M,K]`` (a_layout 0) or ``[K,M]`` (a_layout 1); b: ``[N,K]`` (b_layout 0) or ``[K,N]``
    (b_layout 1).  bf16 operands -> kind::f16; fp32 operands -> kind::tf32 (``split_precision``:
    operands are ``[rows, 2K]`` hi|lo halves, 3xTF32).  Returns ``out``.
    """
    import torch
    assert a.is_cuda and b.is_cuda and a.dim() == 2 and b.dim() == 2
    assert a.stride(1) == 1 and b.stride(1) == 1, "operands must be row-major (unit inner stride)"
    in_dt = DT_BF16 if a.dtype == torch.bfloat16 else DT_F32
    assert a.dtype == b.dtype and a.dtype in (torch.bfloat16, torch.float32)
    kdiv = 2 if split_precision else 1
    M = m if m is not None else (a.shape[1] if a_layout else a.shape[0])
    K = k if k is not None else ((a.shape[0] if a_layout else a.shape[1]) // kdiv)
    N = n if n is not None else (b.shape[1] if b_layout else b.shape[0])
    if out is None:
        if out_dtype is None:
            out_dtype = torch.float32 if epilogue in (EPI_RESIDUAL, EPI_ATOMIC_ADD) else a.dtype
        cols = N * (2 if out_split else 1)
        if epilogue == EPI_ARGMAX:
            out = torch.zeros(M, dtype=torch.int64, device=a.device)
        elif epilogue == EPI_ATOMIC_ADD:
            out = torch.zeros(M, cols, dtype=torch.float32, device=a.device)
        else:
            rows = M if row_remap is None else row_remap[3]
            out = torch.empty(rows, cols, dtype=out_dtype, device=a.device)
    g = GemmDesc()
    g.a, g.b, g.lda, g.ldb = a.data_ptr(), b.data_ptr(), a.stride(0), b.stride(0)
    g.m, g.n, g.k = M, N, K
    g.a_layout, g.b_layout, g.in_dtype = a_layout, b_layout, in_dt
    g.out_dtype = DT_F32 if out.dtype in (torch.float32, torch.int64) else DT_BF16
    g.epilogue, g.splits, g.block_n = epilogue, splits, block_n
    g.split_precision, g.act, g.out_split = int(split_precision), act, int(out_split)
    g.d, g.ldd = out.data_ptr(), (out.stride(0) if out.dim() == 2 else 1)
    g.d2, g.ldd2 = _ptr(d2), (d2.stride(0) if d2 is not None else 0)
    g.bias, g.aux, g.ldaux = _ptr(bias), _ptr(aux), (aux.stride(0) if aux is not None else 0)
    g.colscale, g.rowscale, g.rows_per_group = _ptr(colscale), _ptr(rowscale), rows_per_group
    if row_remap is not None:
        g.out_group_rows, g.out_group_stride, g.out_row_offset = row_remap[0], row_remap[1], row_remap[2]
    g.rowmask, g.maskvec = _ptr(rowmask), _ptr(maskvec)
    g.alpha = alpha
    g.alpha_dev = _ptr(alpha_dev)
    g.colsum = _ptr(colsum)          # EPI_DGELU: [N] fp32, += column sums of the output (the fused bias gradient)
    g.rowdot = _ptr(rowdot)          # EPI_STORE_ROWDOT: fp32 [M / rows_per_group, N / 64, rows_per_group] = per-head rowsum(out * aux)
    g.err_flag = _err_flag(torch, a.device).data_ptr()
    lib = _lib.load()
    _lib.check(lib.memb_gemm(ctypes.byref(g), _lib.stream_ptr(torch, a.device)))
    return out
