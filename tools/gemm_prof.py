"""GPU box: run the ViT-B forward/backward GEMM shapes (B=128) a few times each, print CUDA-event times.
Used under ncu (--set full) to profile gemm_tcgen05 variants."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mem_b200 import ops  # noqa: E402
from mem_b200._lib import EPI_ATOMIC_ADD, EPI_BIAS_GELU, EPI_DGELU, EPI_RESIDUAL, EPI_STORE, EPI_STORE_ROWDOT  # noqa: E402

torch.manual_seed(0)
M, D, Hd = 128 * 197, 768, 3072
dev = "cuda"
x = torch.randn(M, D, device=dev).bfloat16()
h = torch.randn(M, Hd, device=dev).bfloat16()
wqkv = torch.randn(3 * D, D, device=dev).bfloat16()
w1 = torch.randn(Hd, D, device=dev).bfloat16()
w2 = torch.randn(D, Hd, device=dev).bfloat16()
wp = torch.randn(D, D, device=dev).bfloat16()
bias3 = torch.randn(3 * D, device=dev)
bias1 = torch.randn(Hd, device=dev)
biasD = torch.randn(D, device=dev)
gamma = torch.randn(D, device=dev)
res = torch.randn(M, D, device=dev)
out_qkv = torch.empty(M, 3 * D, device=dev, dtype=torch.bfloat16)
out_h = torch.empty(M, Hd, device=dev, dtype=torch.bfloat16)
pre_h = torch.empty(M, Hd, device=dev, dtype=torch.bfloat16)
out_res = torch.empty(M, D, device=dev)
br = torch.empty(M, D, device=dev, dtype=torch.bfloat16)
gw = torch.zeros(Hd, D, device=dev)
rd = torch.zeros(128, D // 64, 197, device=dev)
cases = {
    "qkv_store": (lambda: ops.gemm(x, wqkv, out=out_qkv, bias=bias3), 2.0 * M * 3 * D * D),
    "fc1_gelu": (lambda: ops.gemm(x, w1, out=out_h, epilogue=EPI_BIAS_GELU, bias=bias1, d2=pre_h), 2.0 * M * Hd * D),
    "fc2_residual": (lambda: ops.gemm(h, w2, out=out_res, epilogue=EPI_RESIDUAL, bias=biasD, aux=res, d2=br, colscale=gamma), 2.0 * M * Hd * D),
    "fc2_dgrad_dgelu": (lambda: ops.gemm(x, w2, out=out_h, b_layout=1, epilogue=EPI_DGELU, aux=pre_h), 2.0 * M * Hd * D),
    "fc2_dgrad_dgelu_colsum": (lambda: ops.gemm(x, w2, out=out_h, b_layout=1, epilogue=EPI_DGELU, aux=pre_h, colsum=bias1), 2.0 * M * Hd * D),
    "fc1_dgrad": (lambda: ops.gemm(h, w1, out=br, b_layout=1), 2.0 * M * Hd * D),
    "proj_residual": (lambda: ops.gemm(x, wp, out=out_res, epilogue=EPI_RESIDUAL, bias=biasD, aux=res, d2=br, colscale=gamma), 2.0 * M * D * D),
    "proj_dgrad": (lambda: ops.gemm(x, wp, out=br, b_layout=1), 2.0 * M * D * D),
    "proj_dgrad_rowdot": (lambda: ops.gemm(x, wp, out=br, b_layout=1, epilogue=EPI_STORE_ROWDOT, aux=x, rowdot=rd, rows_per_group=197), 2.0 * M * D * D),
    "qkv_dgrad": (lambda: ops.gemm(out_qkv, wqkv, out=br, b_layout=1), 2.0 * M * 3 * D * D),
    "fc1_wgrad": (lambda: ops.gemm(h, x, out=gw, a_layout=1, b_layout=1, epilogue=EPI_ATOMIC_ADD), 2.0 * M * Hd * D),
}
only = sys.argv[1:] or list(cases)
for name in only:
    fn, fl = cases[name]
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"{name:18s} {ms * 1e3:8.1f} us  {fl / ms / 1e9:7.1f} TFLOP/s", flush=True)
