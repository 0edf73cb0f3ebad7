#!/usr/bin/env python
"""Timeline of one CTA of attention_bwd_tc (debug build with -DMEMB_ATTN_TRACE=<block>): prints clock deltas."""
import ctypes, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mem_b200 import _lib
L = _lib.load()
sp = lambda: _lib.stream_ptr(torch)
B, N, H = 128, 197, 12
D = H * 64; ldk = 200
qkv = (torch.randn(B, N, 3 * D, device="cuda") * 0.8).bfloat16()
bias = torch.zeros(H, N, ldk, device="cuda"); bias[:, :, :N] = torch.randn(H, N, N, device="cuda") * 0.5
PF = _lib.ATTN_BIAS_FLOATS_PER_HEAD
bp = torch.empty(H, PF, device="cuda")
_lib.check(L.memb_attention_pack_bias(bias.data_ptr(), ldk, N, H, bp.data_ptr(), sp()))
out = torch.zeros(B, N, D, device="cuda", dtype=torch.bfloat16); lse = torch.zeros(B, H, N, device="cuda")
dout = (torch.randn(B, N, D, device="cuda") * 0.5).bfloat16()
dqkv = torch.zeros(B, N, 3 * D, device="cuda", dtype=torch.bfloat16); ds = torch.zeros(B, H, N, ldk, device="cuda", dtype=torch.bfloat16)
ws = torch.empty(L.memb_attention_bwd_workspace_bytes(B, N, H), device="cuda", dtype=torch.uint8)
_lib.check(L.memb_attention_fwd(qkv.data_ptr(), bp.data_ptr(), ldk, B, N, H, 64, 0.125, out.data_ptr(), lse.data_ptr(), sp()))
buf = (ctypes.c_longlong * 8192)()
L.memb_attention_trace_dump.restype = ctypes.c_int
L.memb_attention_trace_dump.argtypes = [ctypes.c_void_p, ctypes.c_int]
for it in range(2):
    _lib.check(L.memb_attention_bwd(qkv.data_ptr(), out.data_ptr(), dout.data_ptr(), lse.data_ptr(), bp.data_ptr(), bp.data_ptr(), ldk,
                                    B, N, H, 64, 0.125, dqkv.data_ptr(), ds.data_ptr(), ws.data_ptr(), ws.numel(), sp()))
    n = L.memb_attention_trace_dump(buf, 4096)
ev = sorted(((buf[2 * i + 1], buf[2 * i]) for i in range(n) if buf[2 * i + 1] != 0))
t0 = ev[0][0]
names = {1: "start", 2: "loaded"}
for t, e in ev:
    k = e // 100
    nm = names.get(e) or {1: "I1 PD-wait done g=", 2: "I1 issued g=", 3: "I2 PD seen g=", 4: "I2 committed g=", 5: "EW begin g=", 6: "EW S ok g=",
                          7: "EW B2 ok g=", 8: "EW arrived g=", 9: "EW ACC ok t="}[k] + str(e % 100)
    print(f"{t - t0:8d}  {nm}")
