"""-m gpu: LayerNorm fwd/bwd kernels vs fp32 PyTorch (utils.py:14-19 semantics)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


@pytest.mark.parametrize("rows,D", [(8320, 768), (960, 768), (37, 256), (1000, 1024), (5, 1280)])
def test_layernorm_fwd_bwd(rows, D):
    from prismer_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(0)
    x = (torch.randn(rows, D, device="cuda", generator=g) * 2 + 0.5).to(torch.bfloat16)
    gamma = 1 + 0.1 * torch.randn(D, device="cuda", generator=g)
    beta = 0.1 * torch.randn(D, device="cuda", generator=g)
    y, mean, rstd = ops.layernorm_fwd(x, gamma, beta)
    xf = x.float().requires_grad_(True)
    gf, bf = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    ref = F.layer_norm(xf, (D,), gf, bf, 1e-5)
    torch.cuda.synchronize()
    assert _rel(y.float(), ref) < 3e-3
    assert _rel(mean, xf.mean(-1)) < 1e-5
    dy = torch.randn(rows, D, device="cuda", generator=g).to(torch.bfloat16)
    dres = torch.randn(rows, D, device="cuda", generator=g).to(torch.bfloat16)
    ref.backward(dy.float())
    dgamma = torch.zeros(D, device="cuda"); dbeta = torch.zeros(D, device="cuda")
    dx, dz = ops.layernorm_bwd(dy, x, mean, rstd, gamma, dres=dres, dgamma=dgamma, dbeta=dbeta, dz=True)
    torch.cuda.synchronize()
    assert _rel(dz.float(), xf.grad) < 4e-3
    assert _rel(dx.float(), xf.grad + dres.float()) < 4e-3
    assert _rel(dgamma, gf.grad) < 1e-4
    assert _rel(dbeta, bf.grad) < 1e-4


@pytest.mark.parametrize("rows,D", [(8320, 768), (77, 256), (1030, 1024)])
def test_layernorm_bwd_frozen_gamma(rows, D):
    """Frozen LayerNorm sites (24 of the 36 ViT sites under freeze_vision): no dgamma / dbeta, dx only (+ residual gradient)."""
    from prismer_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn(rows, D, device="cuda", generator=g).to(torch.bfloat16)
    gamma = 1 + 0.1 * torch.randn(D, device="cuda", generator=g)
    _, mean, rstd = ops.layernorm_fwd(x, gamma, torch.zeros(D, device="cuda"))
    dy = torch.randn(rows, D, device="cuda", generator=g).to(torch.bfloat16)
    xf = x.float().requires_grad_(True)
    F.layer_norm(xf, (D,), gamma, None, 1e-5).backward(dy.float())
    dx, dz = ops.layernorm_bwd(dy, x, mean, rstd, gamma)
    torch.cuda.synchronize()
    assert dz is None and _rel(dx.float(), xf.grad) < 4e-3


def test_layernorm_bwd_dropout_matches_gemm_mask():
    """dz of the LN backward must use the very mask the forward GEMM epilogue applied (same seed/stream/element)."""
    from prismer_b200 import ops
    M, D, K = 96, 768, 64
    seed = torch.tensor([42], dtype=torch.int64, device="cuda")
    a = torch.ones((M, K), dtype=torch.bfloat16, device="cuda")
    b = torch.ones((D, K), dtype=torch.bfloat16, device="cuda") / K
    fwd = ops.gemm(a, b, drop_p=0.1, seed=seed, rng_stream=9, out_dtype=torch.float32)   # = mask/(1-p)
    x = torch.randn(M, D, device="cuda").to(torch.bfloat16)
    gamma, beta = torch.ones(D, device="cuda"), torch.zeros(D, device="cuda")
    _, mean, rstd = ops.layernorm_fwd(x, gamma, beta)
    dy = torch.randn(M, D, device="cuda").to(torch.bfloat16)
    dx, dz = ops.layernorm_bwd(dy, x, mean, rstd, gamma, dz=True, drop_p=0.1, seed=seed, rng_stream=9)
    torch.cuda.synchronize()
    assert torch.equal(dz == 0, (fwd == 0) | (dx == 0))
    sel = fwd != 0
    assert _rel(dz.float()[sel], dx.float()[sel] / 0.9) < 5e-3
