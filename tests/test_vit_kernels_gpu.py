"""GPU: each masked-ViT kernel of libmemb against a plain PyTorch fp32 statement of the same op
(floating point: tolerances written at each assert; index / count outputs: exact)."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def L(lib):
    return lib


def sp():
    return int(torch.cuda.current_stream().cuda_stream)


def ck(rc):
    from mem_b200 import _lib
    _lib.check(rc)


def rel_err(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


@pytest.mark.parametrize("rows,D", [(50, 128), (1000, 768), (333, 1024)])
def test_layernorm_fwd_bwd(L, rows, D):
    torch.manual_seed(0)
    x = torch.randn(rows, D, device="cuda") * 2 + 0.5
    w = torch.randn(D, device="cuda") * 0.2 + 1
    b = torch.randn(D, device="cuda") * 0.1
    y = torch.empty(rows, D, device="cuda", dtype=torch.bfloat16)
    mean = torch.empty(rows, device="cuda"); rstd = torch.empty(rows, device="cuda")
    ck(L.memb_layernorm_fwd(x.data_ptr(), D, w.data_ptr(), b.data_ptr(), 1e-6, rows, D, y.data_ptr(), D, mean.data_ptr(),
                            rstd.data_ptr(), None, None, sp()))
    xr = x.clone().requires_grad_(True); wr = w.clone().requires_grad_(True); br = b.clone().requires_grad_(True)
    yr = torch.nn.functional.layer_norm(xr, (D,), wr, br, 1e-6)
    assert rel_err(y.float(), yr) < 4e-3          # bf16 output rounding
    torch.testing.assert_close(mean, x.mean(1), rtol=1e-5, atol=1e-5)
    dy = torch.randn(rows, D, device="cuda")
    yr.backward(dy)
    for dt in (torch.float32, torch.bfloat16):
        dyk = dy.to(dt)
        dx = torch.full((rows, D), 0.25, device="cuda")     # accumulate semantics: dx += ...
        dw = torch.zeros(D, device="cuda"); db = torch.zeros(D, device="cuda")
        ck(L.memb_layernorm_bwd(dyk.data_ptr(), 0 if dt == torch.bfloat16 else 1, D, x.data_ptr(), D, w.data_ptr(),
                                mean.data_ptr(), rstd.data_ptr(), rows, D, dx.data_ptr(), D, dw.data_ptr(), db.data_ptr(),
                                None, None, sp()))
        tol = 1e-4 if dt == torch.float32 else 6e-3
        assert rel_err(dx - 0.25, xr.grad) < tol
        assert rel_err(dw, wr.grad) < tol and rel_err(db, br.grad) < tol


def test_layernorm_gather_rows_and_count(L):
    torch.manual_seed(1)
    rows, D, cap = 200, 128, 64
    x = torch.randn(rows, D, device="cuda")
    w = torch.ones(D, device="cuda"); b = torch.zeros(D, device="cuda")
    idx = torch.randperm(rows, device="cuda")[:cap].to(torch.int32)
    count = torch.tensor([40], device="cuda", dtype=torch.int32)
    y = torch.full((cap, D), 7.0, device="cuda", dtype=torch.bfloat16)
    mean = torch.empty(cap, device="cuda"); rstd = torch.empty(cap, device="cuda")
    ck(L.memb_layernorm_fwd(x.data_ptr(), D, w.data_ptr(), b.data_ptr(), 1e-6, cap, D, y.data_ptr(), D, mean.data_ptr(),
                            rstd.data_ptr(), idx.data_ptr(), count.data_ptr(), sp()))
    ref = torch.nn.functional.layer_norm(x[idx[:40].long()], (D,), w, b, 1e-6)
    assert rel_err(y[:40].float(), ref) < 4e-3
    assert (y[40:] == 0).all()
    dy = torch.randn(cap, D, device="cuda").bfloat16()
    dx = torch.zeros(rows, D, device="cuda")
    ck(L.memb_layernorm_bwd(dy.data_ptr(), 0, D, x.data_ptr(), D, w.data_ptr(), mean.data_ptr(), rstd.data_ptr(), cap, D,
                            dx.data_ptr(), D, None, None, idx.data_ptr(), count.data_ptr(), sp()))
    xr = x.clone().requires_grad_(True)
    torch.nn.functional.layer_norm(xr[idx[:40].long()], (D,), w, b, 1e-6).backward(dy[:40].float())
    assert rel_err(dx, xr.grad) < 1e-4


def _attn_ref(qkv, bias, B, N, H, scale):
    q, k, v = qkv.float().view(B, N, 3, H, 64).permute(2, 0, 3, 1, 4)
    s = (q * scale) @ k.transpose(-1, -2)
    if bias is not None:
        s = s + bias[:, :N, :N].unsqueeze(0)
    p = s.softmax(-1)
    return (p @ v).transpose(1, 2).reshape(B, N, H * 64), p


@pytest.mark.parametrize("B,N,H,with_bias", [(2, 197, 12, True), (3, 50, 2, True), (2, 17, 2, False), (1, 208, 4, True), (13, 197, 12, True), (5, 129, 3, True)])
def test_attention_fwd_bwd(L, B, N, H, with_bias):
    torch.manual_seed(2)
    D = H * 64
    ldk = (N + 7) // 8 * 8
    qkv = (torch.randn(B, N, 3 * D, device="cuda") * 0.8).bfloat16()
    qkv_r = qkv.float().requires_grad_(True)
    bias = torch.zeros(H, N, ldk, device="cuda")
    bias[:, :, :N] = torch.randn(H, N, N, device="cuda") * 0.5
    bias_r = bias.clone().requires_grad_(True)
    biasT = torch.zeros(H, N, ldk, device="cuda")
    biasT[:, :, :N] = bias[:, :, :N].transpose(1, 2)
    from mem_b200._lib import ATTN_BIAS_FLOATS_PER_HEAD as PF
    bias_p = torch.empty(H, PF, device="cuda"); biasT_p = torch.empty(H, PF, device="cuda")
    ck(L.memb_attention_pack_bias(bias.data_ptr(), ldk, N, H, bias_p.data_ptr(), sp()))
    ck(L.memb_attention_pack_bias(biasT.data_ptr(), ldk, N, H, biasT_p.data_ptr(), sp()))
    v = bias_p.view(H, 52, 256, 4).permute(0, 2, 1, 3).reshape(H, 256, 208)   # [h, row, col]
    assert torch.equal(v[:, :N, :N], bias[:, :, :N] * 1.4426950408889634) and (v[:, :, N:] == float("-inf")).all()
    scale = 64 ** -0.5
    out = torch.empty(B, N, D, device="cuda", dtype=torch.bfloat16)
    lse = torch.empty(B, H, N, device="cuda")
    ck(L.memb_attention_fwd(qkv.data_ptr(), bias_p.data_ptr() if with_bias else None, ldk, B, N, H, 64, scale, out.data_ptr(),
                            lse.data_ptr(), sp()))
    ref, _ = _attn_ref(qkv_r, bias_r if with_bias else None, B, N, H, scale)
    assert rel_err(out.float(), ref) < 1e-2       # bf16 P and bf16 output
    dout = (torch.randn(B, N, D, device="cuda") * 0.5).bfloat16()
    ref.backward(dout.float())
    dqkv = torch.zeros(B, N, 3 * D, device="cuda", dtype=torch.bfloat16)
    ds = torch.zeros(B, H, N, ldk, device="cuda", dtype=torch.bfloat16) if with_bias else None
    ws = torch.empty(L.memb_attention_bwd_workspace_bytes(B, N, H), device="cuda", dtype=torch.uint8)
    ck(L.memb_attention_bwd(qkv.data_ptr(), out.data_ptr(), dout.data_ptr(), lse.data_ptr(),
                            bias_p.data_ptr() if with_bias else None, biasT_p.data_ptr() if with_bias else None, ldk, B, N, H, 64,
                            scale, dqkv.data_ptr(), ds.data_ptr() if with_bias else None, ws.data_ptr(), ws.numel(), sp()))
    g = qkv_r.grad.view(B, N, 3, D)
    got = dqkv.float().view(B, N, 3, D)
    for i, name in enumerate("qkv"):
        assert rel_err(got[:, :, i], g[:, :, i]) < 2e-2, name
    # the caller may hand in rowsum(dO * O) itself (out = NULL): here from the separate pass, bit-identical result
    delta = torch.empty(B, H, N, device="cuda")
    ck(L.memb_rowdot_heads(out.data_ptr(), dout.data_ptr(), B, N, H, delta.data_ptr(), sp()))
    want_delta = (out.float() * dout.float()).view(B, N, H, 64).sum(-1).permute(0, 2, 1)
    assert float((delta - want_delta).abs().max()) < 1e-4 * max(1.0, float(want_delta.abs().max()))
    ws2 = torch.zeros_like(ws)
    ws2.view(torch.float32)[:B * H * N] = delta.flatten()
    dqkv2 = torch.zeros_like(dqkv)
    ds2 = torch.zeros_like(ds) if with_bias else None
    ck(L.memb_attention_bwd(qkv.data_ptr(), None, dout.data_ptr(), lse.data_ptr(),
                            bias_p.data_ptr() if with_bias else None, biasT_p.data_ptr() if with_bias else None, ldk, B, N, H, 64,
                            scale, dqkv2.data_ptr(), ds2.data_ptr() if with_bias else None, ws2.data_ptr(), ws2.numel(), sp()))
    assert torch.equal(dqkv2, dqkv) and (not with_bias or torch.equal(ds2, ds))
    if with_bias:
        dbiasT = torch.zeros(H, N, ldk, device="cuda")   # the kernel emits dS^T[b, h, key, query]
        ck(L.memb_batch_reduce_bf16(ds.data_ptr(), B, H * N * ldk, dbiasT.data_ptr(), sp()))
        assert rel_err(dbiasT[:, :, :N].transpose(1, 2), bias_r.grad[:, :, :N]) < 2e-2
        assert (dbiasT[:, :, N:] == 0).all()


def test_mask_compact_and_cross_entropy(L):
    torch.manual_seed(3)
    B, P, V = 5, 49, 512
    mask = (torch.rand(B, P, device="cuda") < 0.4).to(torch.uint8)
    cap = B * P
    row_index = torch.empty(cap, device="cuda", dtype=torch.int32); patch_index = torch.empty_like(row_index)
    count = torch.zeros(1, device="cuda", dtype=torch.int32)
    ck(L.memb_mask_compact(mask.data_ptr(), B, P, row_index.data_ptr(), patch_index.data_ptr(), count.data_ptr(), cap, sp()))
    nz = mask.view(-1).nonzero().view(-1)
    n = nz.numel()
    assert count.item() == n
    assert torch.equal(patch_index[:n].long(), nz)
    assert torch.equal(row_index[:n].long(), nz // P * (P + 1) + 1 + nz % P)
    logits = torch.randn(cap, V, device="cuda") * 3
    tokens = torch.randint(0, V, (B * P,), device="cuda")
    for r in range(0, n, 3):   # plant some correct predictions
        logits[r, tokens[nz[r]]] = 50.0
    dlogits = torch.empty(cap, V, device="cuda", dtype=torch.bfloat16)
    stats = torch.zeros(4, device="cuda")
    ck(L.memb_cross_entropy(logits.data_ptr(), V, tokens.data_ptr(), patch_index.data_ptr(), count.data_ptr(), cap, V,
                            dlogits.data_ptr(), V, stats.data_ptr(), 1.0, sp()))
    lr = logits[:n].clone().requires_grad_(True)
    labels = tokens[nz]
    loss = torch.nn.functional.cross_entropy(lr, labels)
    loss.backward()
    assert abs(stats[0].item() / n - loss.item()) < 1e-4 * max(1.0, loss.item())
    assert stats[1].item() == (lr.argmax(-1) == labels).sum().item()
    assert stats[2].item() == n
    assert rel_err(dlogits[:n].float(), lr.grad) < 6e-3      # bf16 store
    assert (dlogits[n:] == 0).all()


def test_relpos_gather_scatter(L):
    from mem_b200.modeling_finetune import relative_position_index
    idx, nrel = relative_position_index((7, 7))
    N, H = 50, 3
    ldk = 56
    idx = idx.cuda()
    table = torch.randn(nrel, H, device="cuda")
    bias = torch.empty(H, N, ldk, device="cuda"); biasT = torch.empty(H, N, ldk, device="cuda")
    ck(L.memb_relpos_gather(table.data_ptr(), idx.data_ptr(), N, H, ldk, bias.data_ptr(), biasT.data_ptr(), sp()))
    ref = table[idx.view(-1)].view(N, N, H).permute(2, 0, 1)
    assert torch.equal(bias[:, :, :N], ref) and torch.equal(biasT[:, :, :N], ref.transpose(1, 2))
    assert (bias[:, :, N:] == 0).all()
    dbias = torch.randn(H, N, ldk, device="cuda")
    dtable = torch.zeros(nrel, H, device="cuda")
    ck(L.memb_relpos_scatter(dbias.data_ptr(), ldk, idx.data_ptr(), N, H, dtable.data_ptr(), sp()))
    tr = table.clone().requires_grad_(True)
    (tr[idx.view(-1)].view(N, N, H).permute(2, 0, 1) * dbias[:, :, :N]).sum().backward()
    torch.testing.assert_close(dtable, tr.grad, rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("D,with_ls,with_dp", [(768, True, True), (128, True, False), (384, False, True)])
def test_layernorm_bwd_branch_equals_the_two_launches(L, D, with_ls, with_dp):
    """memb_layernorm_bwd_branch == memb_layernorm_bwd followed by memb_branch_bwd on the updated residual gradient
    (same dx bit for bit; dz from the same fp32 row; column sums up to summation order)."""
    torch.manual_seed(9)
    B, N = 5, 197
    rows = B * N
    x = torch.randn(rows, D, device="cuda") * 1.5
    w = torch.randn(D, device="cuda") * 0.2 + 1
    mean = x.mean(1).contiguous(); rstd = (x.var(1, unbiased=False) + 1e-6).rsqrt().contiguous()
    dy = torch.randn(rows, D, device="cuda").bfloat16()
    g0 = torch.randn(rows, D, device="cuda")
    branch = torch.randn(rows, D, device="cuda").bfloat16()
    gamma = (torch.randn(D, device="cuda") * 0.1) if with_ls else None
    rs = torch.tensor([0.0, 1.25, 1.25, 0.0, 1.25], device="cuda") if with_dp else None
    p = lambda t: None if t is None else t.data_ptr()  # noqa: E731

    def run(fused):
        dx = g0.clone()
        dw = torch.zeros(D, device="cuda"); db = torch.zeros(D, device="cuda")
        dz = torch.empty(rows, D, device="cuda", dtype=torch.bfloat16)
        dgam = torch.zeros(D, device="cuda") if with_ls else None
        dbias = torch.zeros(D, device="cuda")
        if fused:
            ck(L.memb_layernorm_bwd_branch(dy.data_ptr(), 0, D, x.data_ptr(), D, w.data_ptr(), mean.data_ptr(), rstd.data_ptr(), rows, D,
                                           dx.data_ptr(), D, dw.data_ptr(), db.data_ptr(), branch.data_ptr(), D, p(gamma), p(rs), N,
                                           dz.data_ptr(), D, p(dgam), dbias.data_ptr(), sp()))
        else:
            ck(L.memb_layernorm_bwd(dy.data_ptr(), 0, D, x.data_ptr(), D, w.data_ptr(), mean.data_ptr(), rstd.data_ptr(), rows, D,
                                    dx.data_ptr(), D, dw.data_ptr(), db.data_ptr(), None, None, sp()))
            ck(L.memb_branch_bwd(dx.data_ptr(), D, branch.data_ptr(), D, p(gamma), p(rs), N, rows, D, dz.data_ptr(), D, p(dgam),
                                 dbias.data_ptr(), sp()))
        return dx, dw, db, dz, dgam, dbias

    a, b = run(True), run(False)
    assert torch.equal(a[0], b[0]) and torch.equal(a[3], b[3])
    for u, v in zip(a[1:3] + a[4:], b[1:3] + b[4:]):
        if u is not None:
            assert rel_err(u, v) < 1e-5


def test_vbias_chain_is_the_column_sum_of_dv(L):
    """dv_bias = colsum(dZ) W_proj equals the column sum of dV = P^T dAO when the rows of P sum to one (what the engine
    relies on instead of a pass over dV), and the kernel is that matrix-vector product."""
    torch.manual_seed(12)
    D = 768
    t = torch.randn(D, device="cuda")
    W = torch.randn(D, D, device="cuda") * 0.05
    pb = torch.full((D,), 0.5, device="cuda"); vb = torch.full((D,), -0.25, device="cuda")
    ck(L.memb_vbias_chain(t.data_ptr(), W.data_ptr(), D, pb.data_ptr(), vb.data_ptr(), sp()))
    assert rel_err(pb - 0.5, t) < 1e-6
    assert rel_err(vb + 0.25, (t.double() @ W.double()).float()) < 1e-5
    # the identity itself, in float64: one head, softmax rows
    N, d = 50, 64
    P = torch.softmax(torch.randn(N, N, dtype=torch.float64), -1)
    dAO = torch.randn(N, d, dtype=torch.float64)
    assert torch.allclose((P.T @ dAO).sum(0), dAO.sum(0), atol=1e-10)


def test_branch_bwd_colsum_patchify_embed(L):
    torch.manual_seed(4)
    B, N, D = 3, 50, 128
    rows = B * N
    gout = torch.randn(rows, D, device="cuda")
    branch = torch.randn(rows, D, device="cuda").bfloat16()
    gamma = torch.randn(D, device="cuda") * 0.1
    rs = torch.tensor([0.0, 1.25, 1.25], device="cuda")
    dz = torch.empty(rows, D, device="cuda", dtype=torch.bfloat16)
    dgam = torch.zeros(D, device="cuda"); dbias = torch.zeros(D, device="cuda")
    ck(L.memb_branch_bwd(gout.data_ptr(), D, branch.data_ptr(), D, gamma.data_ptr(), rs.data_ptr(), N, rows, D, dz.data_ptr(), D,
                         dgam.data_ptr(), dbias.data_ptr(), sp()))
    rsr = rs.repeat_interleave(N).view(rows, 1)
    ref_dz = rsr * gamma * gout
    assert rel_err(dz.float(), ref_dz) < 4e-3
    assert rel_err(dgam, (rsr * gout * branch.float()).sum(0)) < 1e-5
    assert rel_err(dbias, ref_dz.sum(0)) < 1e-5
    x = torch.randn(777, 384, device="cuda").bfloat16()
    out = torch.zeros(384, device="cuda")
    ck(L.memb_colsum_bf16(x.data_ptr(), 384, 777, 384, out.data_ptr(), sp()))
    assert rel_err(out, x.float().sum(0)) < 1e-5
    img = torch.randn(2, 3, 32, 48, device="cuda")
    pat = torch.empty(2 * 2 * 3, 3 * 256, device="cuda", dtype=torch.bfloat16)
    ck(L.memb_patchify(img.data_ptr(), 2, 3, 32, 48, 16, pat.data_ptr(), sp()))
    ref = torch.nn.functional.unfold(img, 16, stride=16).transpose(1, 2).reshape(12, 768)
    assert torch.equal(pat, ref.bfloat16())
    # embed_bwd
    P = N - 1
    mask = (torch.rand(B, P, device="cuda") < 0.4).to(torch.uint8)
    g0 = torch.randn(B, N, D, device="cuda")
    dpatch = torch.empty(B * P, D, device="cuda", dtype=torch.bfloat16)
    dmt = torch.zeros(D, device="cuda"); dcls = torch.zeros(D, device="cuda"); dpb = torch.zeros(D, device="cuda")
    dpos = torch.zeros(N, D, device="cuda")
    ck(L.memb_embed_bwd(g0.data_ptr(), mask.data_ptr(), B, P, D, dpatch.data_ptr(), dmt.data_ptr(), dcls.data_ptr(), dpb.data_ptr(),
                        dpos.data_ptr(), sp()))
    m = mask.bool().view(B, P, 1)
    gp = g0[:, 1:]
    assert torch.equal(dpatch.view(B, P, D), (gp * ~m).bfloat16())
    assert rel_err(dmt, (gp * m).sum((0, 1))) < 1e-5 and rel_err(dcls, g0[:, 0].sum(0)) < 1e-5
    assert rel_err(dpb, (gp * ~m).sum((0, 1))) < 1e-5 and rel_err(dpos, g0.sum(0)) < 1e-5


def test_adamw_matches_torch(L):
    torch.manual_seed(5)
    n = 5 * 1024
    p = torch.randn(n, device="cuda"); g = torch.randn(n, device="cuda") * 3
    m = torch.zeros(n, device="cuda"); v = torch.zeros(n, device="cuda")
    shadow = torch.empty(n, device="cuda", dtype=torch.bfloat16)
    groups = torch.tensor([0, 1, 1, 255, 0], device="cuda", dtype=torch.uint8)
    import ctypes
    lr = (ctypes.c_float * 2)(1e-2, 5e-3); wd = (ctypes.c_float * 2)(0.05, 0.0)
    pr = [torch.nn.Parameter(p[i * 1024:(i + 1) * 1024].clone()) for i in range(5)]
    opt = torch.optim.AdamW([dict(params=[pr[0], pr[4]], lr=1e-2, weight_decay=0.05), dict(params=[pr[1], pr[2]], lr=5e-3, weight_decay=0.0)],
                            betas=(0.9, 0.95), eps=1e-8)
    max_norm = 1.0
    for step in range(1, 4):
        sq = torch.zeros(1, device="cuda")
        live = torch.cat([g[:3072], g[4096:]])
        ck(L.memb_sqnorm(live.data_ptr(), live.numel(), 1.0, sq.data_ptr(), sp()))
        for i in (0, 1, 2, 4):
            pr[i].grad = g[i * 1024:(i + 1) * 1024].clone()
        total = torch.nn.utils.clip_grad_norm_([pr[i] for i in (0, 1, 2, 4)], max_norm)
        assert abs(math.sqrt(sq.item()) - total.item()) < 1e-3 * total.item()
        opt.step()
        ck(L.memb_adamw(p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), shadow.data_ptr(), n, groups.data_ptr(), lr, wd, 2,
                        0.9, 0.95, 1e-8, step, 1.0, max_norm, sq.data_ptr(), sp()))
        g = g * 0.7 + 0.1
    for i in (0, 1, 2, 4):
        torch.testing.assert_close(p[i * 1024:(i + 1) * 1024], pr[i].data, rtol=2e-5, atol=2e-6)
    assert torch.equal(p[3072:4096], torch.cat([x.data for x in pr])[3072:4096])  # frozen chunk untouched
    assert torch.equal(shadow[:1024], p[:1024].bfloat16())
