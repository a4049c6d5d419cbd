"""-m gpu: the tcgen05 GEMM (through the C-ABI) against a plain fp32 PyTorch reference of the same op."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref_act(name, x):
    if name == "quickgelu":
        return x * torch.sigmoid(1.702 * x)
    if name == "gelu":
        return torch.nn.functional.gelu(x)
    if name == "sqrelu":
        return torch.relu(x) ** 2
    if name == "relu":
        return torch.relu(x)
    return x


def _mk(shape, scale=1.0, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(shape, device="cuda", generator=g) * scale).to(torch.bfloat16)


def _rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


@pytest.mark.parametrize("trans_a,trans_b", [(False, False), (False, True), (True, False), (True, True)])
@pytest.mark.parametrize("bn", [64, 128, 256])
def test_gemm_majors(trans_a, trans_b, bn):
    from prismer_b200 import ops
    M, N, K = 384, 512, 320
    a = _mk((K, M) if trans_a else (M, K), seed=1)
    b = _mk((K, N) if trans_b else (N, K), seed=2)
    out = ops.gemm(a, b, trans_a=trans_a, trans_b=trans_b, out_dtype=torch.float32, force_bn=bn)
    A = a.float().t() if trans_a else a.float()
    B = b.float().t() if trans_b else b.float()
    ref = A @ B.t()
    torch.cuda.synchronize()
    assert _rel(out, ref) < 1e-5, (trans_a, trans_b, bn, _rel(out, ref))


@pytest.mark.parametrize("M,N,K", [(8320, 2304, 768), (960, 768, 3072), (100, 72, 40), (129, 50265, 768), (257, 264, 16),
                                   (37632, 1536, 768), (1, 8, 8)])
def test_gemm_shapes_tails(M, N, K):
    from prismer_b200 import ops
    a, b = _mk((M, K), seed=3), _mk((N, K), seed=4)
    ldc = (N + 7) // 8 * 8
    buf = torch.zeros((M, ldc), dtype=torch.float32, device="cuda")
    out = buf[:, :N]
    ops.gemm(a, b, out=out)
    ref = a.float() @ b.float().t()
    torch.cuda.synchronize()
    assert _rel(out, ref) < 1e-5
    if ldc > N:
        assert float(buf[:, N:].abs().max()) == 0.0  # nothing written past N


@pytest.mark.parametrize("act", ["none", "quickgelu", "gelu", "sqrelu", "relu"])
def test_gemm_epilogue_fwd(act):
    from prismer_b200 import ops
    M, N, K = 300, 776, 256
    a, b = _mk((M, K), seed=5), _mk((N, K), 0.1, seed=6)
    bias = torch.randn(N, device="cuda")
    res = _mk((M, N), seed=7)
    aux = torch.empty((M, N), dtype=torch.bfloat16, device="cuda")
    out = ops.gemm(a, b, bias=bias, residual=res, act=act, aux_out=aux)
    z = a.float() @ b.float().t() + bias
    ref = _ref_act(act, z) + res.float()
    torch.cuda.synchronize()
    assert _rel(aux.float(), z) < 3e-3          # bf16 output rounding (2^-9 max relative)
    assert _rel(out.float(), ref) < 3e-3
    # fp32 output: fp32-accumulate accuracy
    out32 = ops.gemm(a, b, bias=bias, residual=res, act=act, out_dtype=torch.float32)
    assert _rel(out32, ref) < 2e-5


@pytest.mark.parametrize("act", ["quickgelu", "gelu", "sqrelu", "relu"])
def test_gemm_epilogue_actgrad_accumulate(act):
    from prismer_b200 import ops
    M, N, K = 260, 512, 192
    dy, w = _mk((M, K), seed=8), _mk((K, N), 0.1, seed=9)   # dgrad: dy[M,K] . W[K,N] (W stored MN-major for this product)
    z = _mk((M, N), seed=10)
    out = ops.gemm(dy, w, trans_b=True, act_grad=act, aux_in=z, out_dtype=torch.float32)
    zf = z.float().requires_grad_(True)
    _ref_act(act, zf).backward(dy.float() @ w.float())
    torch.cuda.synchronize()
    assert _rel(out, zf.grad) < 2e-4
    # accumulate into fp32 (wgrad of shared weights)
    acc = torch.ones((M, N), dtype=torch.float32, device="cuda")
    ops.gemm(dy, w, trans_b=True, out=acc, accumulate=True, alpha=0.5)
    assert _rel(acc, 1 + 0.5 * (dy.float() @ w.float())) < 2e-5


def test_gemm_dropout_mask_is_reproducible_and_unbiased():
    from prismer_b200 import ops
    M, N, K = 512, 768, 64
    a = torch.ones((M, K), dtype=torch.bfloat16, device="cuda")
    b = torch.ones((N, K), dtype=torch.bfloat16, device="cuda") / K
    seed = torch.tensor([1234567], dtype=torch.int64, device="cuda")
    o1 = ops.gemm(a, b, drop_p=0.1, seed=seed, rng_stream=3, out_dtype=torch.float32)
    o2 = ops.gemm(a, b, drop_p=0.1, seed=seed, rng_stream=3, out_dtype=torch.float32, force_bn=64)
    o3 = ops.gemm(a, b, drop_p=0.1, seed=seed, rng_stream=4, out_dtype=torch.float32)
    torch.cuda.synchronize()
    assert torch.equal(o1, o2)            # mask depends on (seed, stream, element) only, not on the tiling
    assert not torch.equal(o1, o3)
    keep = (o1 != 0).float().mean().item()
    assert abs(keep - 0.9) < 5e-3
    assert abs(o1.mean().item() - 1.0) < 5e-3   # inverted dropout is unbiased
    vals = o1[o1 != 0]
    assert torch.allclose(vals, torch.full_like(vals, 1 / 0.9), rtol=1e-6)


@pytest.mark.parametrize("M,N,K,splits", [(768, 768, 8320, 0), (192, 864, 20000, 0), (304, 520, 4100, 7), (128, 64, 512, 2)])
def test_gemm_split_k_wgrad(M, N, K, splits):
    """wgrad shape (both operands MN-major, fp32 accumulate): split-K partial tiles are red.added into the output."""
    from prismer_b200 import ops
    dy, x = _mk((K, M), 0.1, seed=11), _mk((K, N), 0.1, seed=12)
    ref = dy.float().t() @ x.float()
    acc = torch.full((M, N), 0.5, dtype=torch.float32, device="cuda")
    ops.gemm(dy, x, trans_a=True, trans_b=True, out=acc, accumulate=True, force_splits=splits)
    off = torch.full((M, N), 0.5, dtype=torch.float32, device="cuda")
    ops.gemm(dy, x, trans_a=True, trans_b=True, out=off, accumulate=True, force_splits=1)
    torch.cuda.synchronize()
    assert _rel(acc - 0.5, ref) < 5e-5 and _rel(off - 0.5, ref) < 5e-5
