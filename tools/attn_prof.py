#!/usr/bin/env python
"""Time the fused attention kernels at the ViT-B pretraining shape (B=128, N=197, H=12) and check them
against a torch fp32 reference at a small batch. """
import ctypes, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mem_b200 import _lib

L = _lib.load()
sp = lambda: _lib.stream_ptr(torch)
_vp, _i32, _f32 = ctypes.c_void_p, ctypes.c_int, ctypes.c_float
for name in ("memb_attention_fwd_mma", "memb_attention_bwd_mma"):
    if hasattr(L, name):
        fn = getattr(L, name)
        fn.restype = _i32
        fn.argtypes = [a for a in _lib.SIGNATURES[name.replace("_mma", "")][1]]
        if "bwd" in name:
            del fn.argtypes  # legacy signature has no workspace pair
            fn.argtypes = _lib.SIGNATURES["memb_attention_bwd"][1][:-3] + [_vp]


def rel(a, b):
    return ((a - b).norm() / b.norm().clamp_min(1e-20)).item()


def run(B, N, H, check, iters=20):
    torch.manual_seed(0)
    D = H * 64
    ldk = (N + 7) // 8 * 8
    qkv = (torch.randn(B, N, 3 * D, device="cuda") * 0.8).bfloat16()
    bias = torch.zeros(H, N, ldk, device="cuda"); bias[:, :, :N] = torch.randn(H, N, N, device="cuda") * 0.5
    biasT = torch.zeros(H, N, ldk, device="cuda"); biasT[:, :, :N] = bias[:, :, :N].transpose(1, 2)
    PF = _lib.ATTN_BIAS_FLOATS_PER_HEAD
    bias_p = torch.empty(H, PF, device="cuda"); biasT_p = torch.empty(H, PF, device="cuda")
    _lib.check(L.memb_attention_pack_bias(bias.data_ptr(), ldk, N, H, bias_p.data_ptr(), sp()))
    _lib.check(L.memb_attention_pack_bias(biasT.data_ptr(), ldk, N, H, biasT_p.data_ptr(), sp()))
    out = torch.zeros(B, N, D, device="cuda", dtype=torch.bfloat16); lse = torch.zeros(B, H, N, device="cuda")
    dout = (torch.randn(B, N, D, device="cuda") * 0.5).bfloat16()
    dqkv = torch.zeros(B, N, 3 * D, device="cuda", dtype=torch.bfloat16)
    ds = torch.zeros(B, H, N, ldk, device="cuda", dtype=torch.bfloat16)
    scale = 64 ** -0.5
    ws = torch.empty(L.memb_attention_bwd_workspace_bytes(B, N, H), device="cuda", dtype=torch.uint8)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def fwd(fn):
        legacy = fn is not L.memb_attention_fwd
        _lib.check(fn(qkv.data_ptr(), (bias if legacy else bias_p).data_ptr(), ldk, B, N, H, 64, scale, out.data_ptr(), lse.data_ptr(), sp()))

    def bwd(fn):
        legacy = fn is not L.memb_attention_bwd
        _lib.check(fn(qkv.data_ptr(), out.data_ptr(), dout.data_ptr(), lse.data_ptr(), (bias if legacy else bias_p).data_ptr(),
                      (biasT if legacy else biasT_p).data_ptr(), ldk,
                      B, N, H, 64, scale, dqkv.data_ptr(), ds.data_ptr(), *(() if legacy else (ws.data_ptr(), ws.numel())), sp()))

    def timeit(f):
        f(); torch.cuda.synchronize()
        ts = []
        for _ in range(iters):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); f(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        return ts[len(ts) // 2] * 1e3

    res = {}
    fwd(L.memb_attention_fwd); torch.cuda.synchronize(); print("fwd ok", flush=True)
    bwd(L.memb_attention_bwd); torch.cuda.synchronize(); print("bwd ok", flush=True)
    if check:
        q, k, v = qkv.float().view(B, N, 3, H, 64).permute(2, 0, 3, 1, 4)
        qr, kr, vr = q.clone().requires_grad_(True), k.clone().requires_grad_(True), v.clone().requires_grad_(True)
        br = bias[:, :, :N].clone().requires_grad_(True)
        s = (qr * scale) @ kr.transpose(-1, -2) + br.unsqueeze(0)
        ref = (s.softmax(-1) @ vr).transpose(1, 2).reshape(B, N, D)
        ref.backward(dout.float())
        g = dqkv.float().view(B, N, 3, H, 64).permute(2, 0, 3, 1, 4)
        res["err_out"] = rel(out.float(), ref)
        res["err_lse"] = rel(lse, torch.logsumexp(s, -1))
        res["err_dq"], res["err_dk"], res["err_dv"] = rel(g[0], qr.grad), rel(g[1], kr.grad), rel(g[2], vr.grad)
        res["err_dbias"] = rel(ds.float().sum(0)[:, :, :N].transpose(1, 2), br.grad)
    fl = 4.0 * N * N * 64 * B * H
    t = timeit(lambda: fwd(L.memb_attention_fwd)); res["fwd_us"] = round(t, 1); res["fwd_tflops"] = round(fl / t / 1e6, 1)
    t = timeit(lambda: bwd(L.memb_attention_bwd)); res["bwd_us"] = round(t, 1); res["bwd_tflops"] = round(2.5 * fl / t / 1e6, 1)
    if hasattr(L, "memb_attention_fwd_mma") and "legacy" in sys.argv:
        res["fwd_mma_us"] = round(timeit(lambda: fwd(L.memb_attention_fwd_mma)), 1)
        res["bwd_mma_us"] = round(timeit(lambda: bwd(L.memb_attention_bwd_mma)), 1)
    print(f"B={B} N={N} H={H}", res, flush=True)


if __name__ == "__main__":
    if "small" in sys.argv:
        run(2, 197, 2, True, iters=1)
        sys.exit(0)
    if "ncu" in sys.argv:
        run(128, 197, 12, False, iters=2)
        sys.exit(0)
    run(4, 197, 12, True)
    run(2, 100, 3, True)
    run(128, 197, 12, False)
    run(128, 197, 16, False)
