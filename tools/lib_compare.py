"""GPU: same-box library table -> gpurun_out/lib_compare.json.

For each hot op of the MEM step: this repo's kernel (through the C ABI) next to what the reference executes on the same
B200 -- torch / cuBLASLt / cuDNN / ATen running the reference's modules' ops (F.linear, F.gelu, softmax attention with the
additive relative-position bias, nn.Conv2d).  Evidence only: torch stays out of mem_b200/.

    python tools/lib_compare.py            # ViT-B/16, batch 128 (BASELINE config 3 shapes)
"""
import json
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mem_b200 import _lib, modeling_pretrain, ops, registry, vit_engine  # noqa: E402,F401
from mem_b200._lib import EPI_ATOMIC_ADD, EPI_BIAS_GELU, EPI_RESIDUAL  # noqa: E402
from mem_b200.vae_model import DiscreteVAE  # noqa: E402
from oracle import dvae_ref, vit_ref  # noqa: E402

dev = torch.device("cuda")
FLUSH = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def bench(fn, iters=10, warm=3, flush=True):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush:
            FLUSH.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    B, N, D, H, hidden, V = 128, 197, 768, 12, 3072, 8192
    M = B * N
    rows = []
    g = torch.Generator(device=dev).manual_seed(0)

    def rnd(*shape, dtype=torch.bfloat16, scale=1.0):
        return (torch.randn(*shape, device=dev, generator=g) * scale).to(dtype)

    def add(name, ours_ms, lib_ms, flops=None, note=""):
        r = {"op": name, "ours_ms": round(ours_ms, 4), "library_ms": round(lib_ms, 4), "speedup": round(lib_ms / ours_ms, 3), "note": note}
        if flops:
            r["ours_tflops"] = round(flops / ours_ms / 1e9, 1)
            r["library_tflops"] = round(flops / lib_ms / 1e9, 1)
        rows.append(r)
        print(r, flush=True)

    # ---------------- forward GEMMs with their fused epilogues vs F.linear + the unfused ops
    x = rnd(M, D); xres = rnd(M, D, dtype=torch.float32); gamma = rnd(D, dtype=torch.float32, scale=0.1)
    for name, n_out, k_in in (("qkv", 3 * D, D), ("fc1", hidden, D)):
        w = rnd(n_out, k_in, scale=0.02); b = rnd(n_out, dtype=torch.float32, scale=0.02)
        a = x if k_in == D else rnd(M, k_in)
        out = torch.empty(M, n_out, dtype=torch.bfloat16, device=dev); pre = torch.empty_like(out)
        if name == "fc1":
            ours = lambda: ops.gemm(a, w, out=out, epilogue=EPI_BIAS_GELU, bias=b, d2=pre)
            lib = lambda: F.gelu(F.linear(a, w, b.bfloat16()))
            note = "bias + GELU (+ pre-activation kept for backward) fused vs F.linear + F.gelu"
        else:
            ours = lambda: ops.gemm(a, w, out=out, bias=b)
            lib = lambda: F.linear(a, w, b.bfloat16())
            note = "bias fused (both)"
        add(f"fwd {name} [{M}x{k_in}]x[{n_out}x{k_in}]^T", bench(ours), bench(lib), 2.0 * M * n_out * k_in, note)
    for name, k_in in (("proj", D), ("fc2", hidden)):
        w = rnd(D, k_in, scale=0.02); b = rnd(D, dtype=torch.float32, scale=0.02)
        a = rnd(M, k_in)
        out = torch.empty(M, D, dtype=torch.float32, device=dev); br = torch.empty(M, D, dtype=torch.bfloat16, device=dev)
        ours = lambda: ops.gemm(a, w, out=out, epilogue=EPI_RESIDUAL, bias=b, aux=xres, d2=br, colscale=gamma)
        lib = lambda: xres + gamma * F.linear(a, w, b.bfloat16()).float()
        add(f"fwd {name} [{M}x{k_in}]x[{D}x{k_in}]^T + LayerScale + residual", bench(ours), bench(lib), 2.0 * M * D * k_in,
            "bias, gamma, fp32 residual add fused vs F.linear + 2 elementwise kernels")
    # ---------------- wgrad (dW += dY^T X) and dgrad
    dy = rnd(M, hidden); xin = rnd(M, D); gw = torch.zeros(hidden, D, dtype=torch.float32, device=dev)
    add(f"wgrad fc1 [{hidden}x{M}]x[{M}x{D}] (fp32 accumulate into the flat gradient)",
        bench(lambda: ops.gemm(dy, xin, out=gw, a_layout=1, b_layout=1, epilogue=EPI_ATOMIC_ADD)),
        bench(lambda: gw.add_(torch.matmul(dy.t(), xin).float())), 2.0 * M * hidden * D, "split-K red.add epilogue vs matmul + add_")
    w1 = rnd(hidden, D, scale=0.02); dx = torch.empty(M, D, dtype=torch.bfloat16, device=dev)
    add(f"dgrad fc1 [{M}x{hidden}]x[{hidden}x{D}]", bench(lambda: ops.gemm(dy, w1, out=dx, b_layout=1)),
        bench(lambda: torch.matmul(dy, w1)), 2.0 * M * hidden * D, "MN-major B operand (no transpose) vs matmul")
    wl = rnd(V, D, scale=0.02); xm = rnd(9600, D); lg = torch.empty(9600, V, dtype=torch.float32, device=dev); bl = rnd(V, dtype=torch.float32)
    add("lm_head [9600x768]x[8192x768]^T -> fp32 logits", bench(lambda: ops.gemm(xm, wl, out=lg, bias=bl)),
        bench(lambda: F.linear(xm, wl, bl.bfloat16()).float()), 2.0 * 9600 * V * D, "fp32 logits written directly vs linear + cast")

    # ---------------- attention fwd / bwd vs SDPA with the additive bias
    lib_ = _lib.load()
    sp = _lib.stream_ptr(torch, dev)
    qkv = rnd(M, 3 * D, scale=0.5)
    ldk = (N + 7) // 8 * 8
    bias_dense = torch.zeros(H, N, ldk, dtype=torch.float32, device=dev)
    bias_dense[:, :, :N] = torch.randn(H, N, N, device=dev, generator=g) * 0.2
    packed = torch.empty(H * _lib.ATTN_BIAS_FLOATS_PER_HEAD, dtype=torch.float32, device=dev)
    _lib.check(lib_.memb_attention_pack_bias(bias_dense.data_ptr(), ldk, N, H, packed.data_ptr(), sp))
    ao = torch.empty(M, D, dtype=torch.bfloat16, device=dev); lse = torch.empty(B, H, N, dtype=torch.float32, device=dev)
    scale = (D // H) ** -0.5
    ours_f = lambda: _lib.check(lib_.memb_attention_fwd(qkv.data_ptr(), packed.data_ptr(), ldk, B, N, H, D // H, scale, ao.data_ptr(),
                                                        lse.data_ptr(), sp))
    q, k, v = (t.contiguous() for t in qkv.view(B, N, 3, H, D // H).permute(2, 0, 3, 1, 4))
    mask = bias_dense[:, :, :N].unsqueeze(0).bfloat16().contiguous()
    lib_f = lambda: F.scaled_dot_product_attention(q, k, v, attn_mask=mask, scale=scale)
    fl_f = 4.0 * B * H * N * N * (D // H)
    add(f"attention fwd B={B} H={H} N={N} d=64 + rel-pos bias", bench(ours_f), bench(lib_f), fl_f, "tcgen05 fused kernel vs SDPA(attn_mask=bias)")

    def lib_eager():
        s = (q * scale) @ k.transpose(-2, -1) + mask
        return s.softmax(-1) @ v
    add("attention fwd, reference formulation (q@k^T + bias, softmax, @v under autocast)", bench(ours_f), bench(lib_eager), fl_f,
        "what mem/modeling_finetune.py:128-157 executes")
    qs = [t.detach().clone().requires_grad_(True) for t in (q, k, v)]
    o = F.scaled_dot_product_attention(*qs, attn_mask=mask, scale=scale)
    do = rnd(B, H, N, D // H, scale=0.1)
    lib_b = lambda: torch.autograd.grad(o, qs, do, retain_graph=True)
    t_b = bench(lib_b)
    rows.append({"op": "attention bwd: SDPA backward (library side; ours: attention_bwd_tc + attn_bwd_prep = 0.18 ms per layer in "
                       "profiles/r02_pretrain_launch_shares_*.txt)",
                 "library_ms": round(t_b, 4), "library_tflops": round(10.0 * B * H * N * N * (D // H) / t_b / 1e9, 1)})
    print(rows[-1], flush=True)

    # ---------------- whole ViT fwd + CE + bwd: this repo vs the reference formulation under bf16 autocast
    torch.manual_seed(0)
    model = registry.create_model("beit_base_patch16_224_8k_vocab", drop_path_rate=0.0, use_shared_rel_pos_bias=True,
                                  use_abs_pos_emb=False, init_values=0.1, in_chans=2).to(dev).train()
    img, maskp, tokens = vit_ref.synth_inputs(B, 2, 224, 224, 196, V, seed=7, n_mask=75)
    img, maskp, tokens = img.to(dev), maskp.to(dev), tokens.to(dev)
    ours_step = lambda: vit_engine.pretrain_step(model, img, maskp, tokens, cap=B * 75)
    sd = {kk: (vv.detach().clone().requires_grad_(True) if vv.is_floating_point() else vv) for kk, vv in model.state_dict().items()}

    def lib_step():
        with torch.autocast("cuda", dtype=torch.bfloat16):
            loss, _, _ = vit_ref.mem_loss(img, maskp, tokens, sd, 12, 16)
        loss.backward()
    flops = 3.0 * B * (12 * (2 * N * D * 3 * D + 4 * H * N * N * 64 + 2 * N * D * D + 4 * N * D * hidden) + 2 * 75 * D * V + 2 * 196 * 512 * D)
    add("ViT-B/16 masked fwd + CE + bwd, B=128 (no tokenizer, no optimizer)", bench(ours_step, iters=5, warm=2, flush=False),
        bench(lib_step, iters=5, warm=2, flush=False), flops, "libmemb schedule vs the reference's op sequence (oracle/vit_ref.py) under bf16 autocast")

    # ---------------- dVAE tokenizer vs cuDNN fp32 (TF32 off: what 'fp32' means; TF32 on: what the reference gets on Ampere+)
    cfg = dict(input_H=224, input_W=224, num_tokens=8192, codebook_dim=32, num_layers=4, num_resnet_blocks=3, hidden_dim=384, channels=2)
    torch.manual_seed(0)
    vae = DiscreteVAE(**cfg).to(dev)
    imgs = dvae_ref.synth_images(B, 2, 224, 224, seed=6).to(dev)
    vsd = {kk: vv.detach() for kk, vv in vae.state_dict().items()}
    ours_tok = lambda: vae.get_codebook_indices(imgs)
    tok = ours_tok()

    def lib_tok():
        with torch.no_grad():
            return torch.cat([dvae_ref.codebook_indices(imgs[i:i + 32], vsd, 4, 3) for i in range(0, B, 32)])
    t_ours = bench(ours_tok, iters=5, warm=2, flush=False)
    t_fp32 = bench(lib_tok, iters=3, warm=1, flush=False)
    diff_fp32 = int((lib_tok() != tok).sum())
    torch.backends.cudnn.allow_tf32 = True
    t_tf32 = bench(lib_tok, iters=3, warm=1, flush=False)
    diff_tf32 = int((lib_tok() != tok).sum())
    torch.backends.cudnn.allow_tf32 = False
    fl = 24.26e9 * B
    add("dVAE tokenizer B=128 vs cuDNN fp32 (allow_tf32=False)", t_ours, t_fp32, fl, f"token differences vs cuDNN fp32: {diff_fp32} of {tok.numel()}")
    add("dVAE tokenizer B=128 vs cuDNN TF32 (allow_tf32=True)", t_ours, t_tf32, fl,
        f"token differences vs cuDNN TF32: {diff_tf32} of {tok.numel()} (TF32 convolutions are not index-exact)")
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump({"gpu": torch.cuda.get_device_name(0), "torch": torch.__version__, "l2": "256 MB memset between timed launches (single ops)",
               "rows": rows}, open(os.path.join(ROOT, "gpurun_out", "lib_compare.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
