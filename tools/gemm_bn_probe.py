"""GPU: the two epilogue-heavy N = 3072 GEMMs at each CTA-pair tile width (and the single-CTA kernel)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mem_b200 import ops
from mem_b200._lib import EPI_BIAS_GELU, EPI_DGELU, EPI_RESIDUAL
torch.manual_seed(0)
M, D, Hd = 128 * 197, 768, 3072
dev = "cuda"
x = torch.randn(M, D, device=dev).bfloat16(); h = torch.randn(M, Hd, device=dev).bfloat16()
w1 = torch.randn(Hd, D, device=dev).bfloat16(); w2 = torch.randn(D, Hd, device=dev).bfloat16()
bias1 = torch.randn(Hd, device=dev); biasD = torch.randn(D, device=dev); gamma = torch.randn(D, device=dev)
res = torch.randn(M, D, device=dev); out_res = torch.empty(M, D, device=dev); br = torch.empty(M, D, device=dev, dtype=torch.bfloat16)
out_h = torch.empty(M, Hd, device=dev, dtype=torch.bfloat16); pre_h = torch.randn(M, Hd, device=dev).bfloat16()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ts.sort(); return ts[len(ts) // 2] * 1e3
for bn in (0, 128, 192, 256):
    a = t(lambda: ops.gemm(x, w1, out=out_h, epilogue=EPI_BIAS_GELU, bias=bias1, d2=pre_h, block_n=bn))
    c = t(lambda: ops.gemm(h, w2, out=out_res, epilogue=EPI_RESIDUAL, bias=biasD, aux=res, d2=br, colscale=gamma, block_n=bn))
    b = t(lambda: ops.gemm(x, w2, out=out_h, b_layout=1, epilogue=EPI_DGELU, aux=pre_h, block_n=bn)) if bn != 192 else float("nan")
    print(f"block_n {bn:3d}: fc1 bias+gelu {a:6.1f} us   fc2 dgrad dgelu {b:6.1f} us   fc2 residual {c:6.1f} us", flush=True)
