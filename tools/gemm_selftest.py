"""GPU bring-up of the tcgen05 GEMM: each case runs in its own subprocess (a trap or hang in one
case must not take the others down), compares against torch fp32/fp64 matmul and, on mismatch,
prints structure hints (which rows/cols are wrong, error histogram)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CASES = [
    # name, M, N, K, a_layout, b_layout, dtype, extra
    ("nt_bf16_small", 128, 256, 64, 0, 0, "bf16", {}),
    ("nt_bf16_k256", 128, 256, 256, 0, 0, "bf16", {}),
    ("nt_bf16_multi", 384, 768, 768, 0, 0, "bf16", {}),
    ("nt_bf16_tail", 197 * 3, 768, 768, 0, 0, "bf16", {}),
    ("nt_bf16_bn128", 300, 384, 192, 0, 0, "bf16", {"block_n": 128}),
    ("nt_bf16_big", 25216, 2304, 768, 0, 0, "bf16", {}),
    ("nn_bf16 (B MN-major, dgrad)", 384, 768, 512, 0, 1, "bf16", {}),
    ("tn_bf16 (A,B MN-major, wgrad)", 768, 512, 1000, 1, 1, "bf16", {}),
    ("tn_bf16_atomic_splitk", 768, 768, 25216, 1, 1, "bf16", {"epilogue": 3}),
    ("nt_tf32", 256, 256, 96, 0, 0, "tf32", {}),
    ("nt_tf32_bn128", 300, 384, 6144, 0, 0, "tf32", {"block_n": 128}),
    ("nt_3xtf32", 512, 384, 6144, 0, 0, "3xtf32", {"block_n": 128}),
    ("nt_bf16_f32out", 256, 512, 128, 0, 0, "bf16", {"out_f32": True}),
]


def run_case(idx):
    import torch
    from mem_b200 import ops
    name, M, N, K, al, bl, dt, extra = CASES[idx]
    torch.manual_seed(idx)
    dev = "cuda"
    if dt == "bf16":
        A = (torch.randn(M, K, device=dev) * 0.5).to(torch.bfloat16)
        B = (torch.randn(N, K, device=dev) * 0.5).to(torch.bfloat16)
        ref = A.double() @ B.double().t()
        a_in = A.t().contiguous() if al else A
        b_in = B.t().contiguous() if bl else B
        kw = dict(a_layout=al, b_layout=bl, block_n=extra.get("block_n", 0), epilogue=extra.get("epilogue", 0))
        if extra.get("out_f32") or extra.get("epilogue") == 3:
            kw["out_dtype"] = torch.float32
        out = ops.gemm(a_in, b_in, **kw)
        tol = 2e-2 if out.dtype == torch.bfloat16 else 1e-3
    elif dt == "tf32":
        A = torch.randn(M, K, device=dev) * 0.5
        B = torch.randn(N, K, device=dev) * 0.5
        ref = A.double() @ B.double().t()
        out = ops.gemm(A, B, block_n=extra.get("block_n", 0), out_dtype=torch.float32)
        tol = 5e-3
    else:
        A = torch.randn(M, K, device=dev) * 0.5
        B = torch.randn(N, K, device=dev) * 0.5
        ref = A.double() @ B.double().t()

        def split(x):
            hi = (x.view(torch.int32) + 0x1000 & ~0x1fff).view(torch.float32)  # rna to tf32
            lo = x - hi
            lo = (lo.view(torch.int32) + 0x1000 & ~0x1fff).view(torch.float32)
            return torch.cat([hi, lo], 1).contiguous()
        out = ops.gemm(split(A), split(B), split_precision=True, block_n=extra.get("block_n", 0), out_dtype=torch.float32)
        fp32 = (A @ B.t()).double()
        print(f"   (torch fp32 matmul max err vs fp64: {(fp32 - ref).abs().max().item():.3e})")
        tol = 2e-5
    torch.cuda.synchronize()
    err = (out.double() - ref).abs()
    scale = ref.abs().max().item()
    rel = err.max().item() / scale
    ok = rel < tol
    print(f"[{'OK ' if ok else 'BAD'}] {name}: M={M} N={N} K={K} max_abs_err={err.max().item():.4e} rel={rel:.3e} (tol {tol})")
    if not ok:
        bad = err > tol * scale
        rows = bad.any(1).nonzero().flatten()
        cols = bad.any(0).nonzero().flatten()
        print(f"   bad elements {int(bad.sum())}/{bad.numel()}; bad rows {rows.numel()} (first {rows[:8].tolist()}), "
              f"bad cols {cols.numel()} (first {cols[:8].tolist()})")
        print("   out[0,:8] ", out[0, :8].float().tolist())
        print("   ref[0,:8] ", ref[0, :8].float().tolist())
        # does it match with K truncated / permuted?  quick hints
        for kk in (16, 32, 64):
            if kk < K and dt == "bf16":
                part = A[:, :kk].double() @ B[:, :kk].double().t()
                print(f"   err vs K[:{kk}] partial: {(out.double() - part).abs().max().item():.3e}")
    return ok


if __name__ == "__main__":
    if len(sys.argv) > 1:
        sys.exit(0 if run_case(int(sys.argv[1])) else 1)
    results = {}
    for i, c in enumerate(CASES):
        try:
            r = subprocess.run([sys.executable, __file__, str(i)], capture_output=True, text=True, timeout=120)
            out = (r.stdout + r.stderr).strip().splitlines()
            print("\n".join(out[-12:]))
            results[c[0]] = r.returncode
        except subprocess.TimeoutExpired:
            print(f"[TIMEOUT] {c[0]}")
            results[c[0]] = "timeout"
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(results, open(os.path.join(ROOT, "gpurun_out", "gemm_selftest.json"), "w"), indent=1)
    print(results)
