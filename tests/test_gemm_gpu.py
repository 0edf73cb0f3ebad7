"""GPU: the tcgen05 GEMMs (single-CTA kernel and the CTA-pair kernel with the TMA-staged epilogue) against a
plain PyTorch fp32 statement of the same fused op.  bf16 tensor-core operands with fp32 accumulation: outputs
stored as bf16 are compared at bf16 rounding (relative L2 <= 4e-3), fp32 outputs at 1e-3 of the output scale.
Reference ops: Mlp / Attention / Block of mem/modeling_finetune.py:66-71,128-157,182-189."""
import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(180)]

from mem_b200._lib import EPI_BIAS_GELU, EPI_DGELU, EPI_RESIDUAL, EPI_STORE, EPI_STORE_ROWDOT  # noqa: E402


def rel_err(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def operands(M, N, K, b_mn, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    A = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    W = (torch.randn(N, K, device="cuda", generator=g) / K ** 0.5).bfloat16()
    acc = A.float() @ W.float().t()
    b_in = W.t().contiguous() if b_mn else W
    return A, b_in, acc, g


# (M, N, K): the ViT-B shapes at a few images (M tail: 197*16 = 3152 = 12*256 + 80) and an odd small case
SHAPES = [(3152, 768, 768), (3152, 2304, 768), (3152, 3072, 768), (3152, 768, 3072), (1000, 384, 200 * 8)]


@pytest.mark.parametrize("M,N,K", SHAPES)
@pytest.mark.parametrize("b_mn", [0, 1])
@pytest.mark.parametrize("block_n", [0, 128, 256])
def test_store_bias(lib, M, N, K, b_mn, block_n):
    from mem_b200 import ops
    A, B, acc, g = operands(M, N, K, b_mn, 1)
    bias = torch.randn(N, device="cuda", generator=g)
    out = ops.gemm(A, B, b_layout=b_mn, bias=bias, block_n=block_n)
    torch.cuda.synchronize()
    assert out.dtype == torch.bfloat16 and rel_err(out, acc + bias) < 4e-3


@pytest.mark.parametrize("M,N,K", SHAPES[:3])
@pytest.mark.parametrize("block_n", [0, 192])
def test_bias_gelu(lib, M, N, K, block_n):
    from mem_b200 import ops
    A, B, acc, g = operands(M, N, K, 0, 2)
    bias = torch.randn(N, device="cuda", generator=g)
    pre = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    act = ops.gemm(A, B, epilogue=EPI_BIAS_GELU, bias=bias, d2=pre, block_n=block_n)
    torch.cuda.synchronize()
    ref_pre = acc + bias
    ref = torch.nn.functional.gelu(ref_pre)  # exact-erf GELU
    assert rel_err(pre, ref_pre) < 4e-3
    assert rel_err(act, ref) < 4e-3
    # the logistic-polynomial GELU itself: <= 3e-5 absolute before the bf16 store (half a bf16 ulp is 2^-9 relative)
    assert (act.float() - ref).abs().max().item() <= 3e-5 + ref.abs().max().item() * 2 ** -8


@pytest.mark.parametrize("M,N,K", [(3152, 768, 768), (3152, 768, 3072), (1000, 384, 1600)])
@pytest.mark.parametrize("with_scales", [False, True])
def test_residual(lib, M, N, K, with_scales):
    from mem_b200 import ops
    A, B, acc, g = operands(M, N, K, 0, 3)
    bias = torch.randn(N, device="cuda", generator=g)
    xin = torch.randn(M, N, device="cuda", generator=g)
    rows_per_group = 197
    gamma = torch.randn(N, device="cuda", generator=g) if with_scales else None
    rs = (torch.rand((M + rows_per_group - 1) // rows_per_group, device="cuda", generator=g) + 0.5) if with_scales else None
    br = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    out = ops.gemm(A, B, epilogue=EPI_RESIDUAL, bias=bias, aux=xin, d2=br, colscale=gamma, rowscale=rs,
                   rows_per_group=rows_per_group)
    torch.cuda.synchronize()
    branch = acc + bias
    ref = branch
    if with_scales:
        ref = ref * gamma * rs.repeat_interleave(rows_per_group)[:M, None]
    ref = xin + ref
    assert out.dtype == torch.float32
    assert (out - ref).abs().max().item() < 1e-3 * ref.abs().max().item()
    assert rel_err(br, branch) < 4e-3


@pytest.mark.parametrize("M,N,K", [(3152, 3072, 768), (1000, 384, 1600)])
@pytest.mark.parametrize("block_n", [0, 128])
def test_dgelu(lib, M, N, K, block_n):
    from mem_b200 import ops
    A, B, acc, g = operands(M, N, K, 1, 4)  # dgrad: B = W as stored ([K, N] row-major)
    pre = (torch.randn(M, N, device="cuda", generator=g) * 1.5).bfloat16()
    out = ops.gemm(A, B, b_layout=1, epilogue=EPI_DGELU, aux=pre, block_n=block_n)
    torch.cuda.synchronize()
    x = pre.float().requires_grad_(True)
    torch.nn.functional.gelu(x).backward(acc)
    assert rel_err(out, x.grad) < 4e-3


@pytest.mark.parametrize("M,N,K", [(3152, 3072, 768), (1000, 384, 1600), (300, 768, 768)])
def test_dgelu_fused_column_sums(lib, M, N, K):
    """The DGELU epilogue's optional bias gradient: += column sums of its own output over the M rows (M tails, N not a
    multiple of the tile, accumulation into a non-zero buffer; M = 300 takes the single-CTA kernel + separate pass)."""
    from mem_b200 import ops
    A, B, acc, g = operands(M, N, K, 1, 7)
    pre = (torch.randn(M, N, device="cuda", generator=g) * 1.5).bfloat16()
    start = torch.randn(N, device="cuda", generator=g)
    cs = start.clone()
    out = ops.gemm(A, B, b_layout=1, epilogue=EPI_DGELU, aux=pre, colsum=cs)
    plain = ops.gemm(A, B, b_layout=1, epilogue=EPI_DGELU, aux=pre)
    torch.cuda.synchronize()
    assert torch.equal(out, plain)                           # the extra reduction does not touch the stored tile
    x = pre.float().requires_grad_(True)
    torch.nn.functional.gelu(x).backward(acc)
    want = x.grad.double().sum(0)
    got = (cs - start).double()
    scale = x.grad.double().abs().sum(0)                     # error budget: bf16-level error per term (fitted GELU derivative)
    assert float(((got - want).abs() / scale).max()) < 4e-3
    # against the sum of what was stored (bf16-rounded terms): rounding noise only, it averages out over the rows
    assert float(((got - out.double().sum(0)).abs() / scale).max()) < 5e-4


@pytest.mark.parametrize("Bimg,Ntok,N,K,block_n", [(16, 197, 768, 768, 0), (16, 197, 768, 768, 256), (5, 200, 1024, 512, 0),
                                                    (2, 150, 128, 256, 0)])
def test_store_rowdot(lib, Bimg, Ntok, N, K, block_n):
    """The proj-dgrad epilogue that also leaves the attention backward's rowsum(dO * O): d = bf16(acc) as the plain store,
    rowdot[img][head][token] = sum over the head's 64 columns of d * aux (M tails, both tile widths, rows of one tile in
    different images; M = 300 takes the single-CTA kernel + separate pass)."""
    from mem_b200 import ops
    M = Bimg * Ntok
    A, B, acc, g = operands(M, N, K, 1, 9)
    aux = torch.randn(M, N, device="cuda", generator=g).bfloat16()
    rd = torch.full((Bimg, N // 64, Ntok), float("nan"), device="cuda")
    out = ops.gemm(A, B, b_layout=1, epilogue=EPI_STORE_ROWDOT, aux=aux, rowdot=rd, rows_per_group=Ntok, block_n=block_n)
    plain = ops.gemm(A, B, b_layout=1, block_n=block_n)
    torch.cuda.synchronize()
    assert torch.equal(out, plain)
    want = (out.float() * aux.float()).view(Bimg, Ntok, N // 64, 64).sum(-1).permute(0, 2, 1)
    assert torch.isfinite(rd).all()
    assert float((rd - want).abs().max()) < 2e-5 * float((out.float().abs() * aux.float().abs()).view(Bimg, Ntok, -1, 64).sum(-1).max())
    with pytest.raises(ValueError):
        ops.gemm(A, B, b_layout=1, epilogue=EPI_STORE_ROWDOT, aux=aux, rowdot=rd, rows_per_group=Ntok + 1)


def test_pair_and_single_cta_agree(lib):
    """M below the pair kernel's threshold runs on the single-CTA kernel: same math, same GELU."""
    from mem_b200 import ops
    A, B, acc, g = operands(3152, 768, 768, 0, 5)
    bias = torch.randn(768, device="cuda", generator=g)
    big = ops.gemm(A, B, epilogue=EPI_BIAS_GELU, bias=bias)
    small = ops.gemm(A[:300], B, epilogue=EPI_BIAS_GELU, bias=bias)
    torch.cuda.synchronize()
    assert rel_err(small, big[:300]) < 1e-3
