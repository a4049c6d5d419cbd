"""-m gpu: BatchNorm statistics, im2col and the fused BatchNorm/ReLU/col2im backward kernels vs fp32 PyTorch."""
from types import SimpleNamespace

import pytest
import torch
import torch.nn.functional as F

from tests.helpers import rel_l2

pytestmark = pytest.mark.gpu


def _nhwc(t):  # [B,C,H,W] -> [B*H*W, C]
    return t.permute(0, 2, 3, 1).reshape(-1, t.shape[1]).contiguous()


@pytest.mark.parametrize("B,C,H,W,ksz,stride", [(2, 256, 4, 4, 1, 1), (2, 32, 32, 32, 3, 2), (3, 96, 14, 14, 3, 1), (2, 64, 9, 7, 3, 2)])
def test_bn_relu_col2im_backward(B, C, H, W, ksz, stride):
    from prismer_b200 import ops
    g0 = torch.Generator(device="cuda").manual_seed(0)
    rt = lambda t: t.bfloat16().float()
    y = rt(torch.randn(B, C, H, W, device="cuda", generator=g0)).requires_grad_(True)
    gamma = (1 + 0.1 * torch.randn(C, device="cuda", generator=g0)).requires_grad_(True)
    beta = (0.05 * torch.randn(C, device="cuda", generator=g0)).requires_grad_(True)
    bn = SimpleNamespace(weight=gamma.detach(), bias=beta.detach(), running_mean=torch.zeros(C, device="cuda"),
                         running_var=torch.ones(C, device="cuda"), eps=1e-5, momentum=0.1)
    y16 = _nhwc(y.detach()).to(torch.bfloat16)
    scale, shift, mean, rstd = ops.bn_stats(y16, bn, True)
    n = F.batch_norm(y, None, None, gamma, beta, True, 0.1, 1e-5)
    a = torch.relu(n)
    a = a + (rt(a) - a).detach()
    yd = y.detach()
    assert rel_l2(mean, yd.mean((0, 2, 3))) < 1e-5 and rel_l2(rstd, torch.rsqrt(yd.var((0, 2, 3), unbiased=False) + 1e-5)) < 1e-5
    assert rel_l2(bn.running_var, 0.9 + 0.1 * yd.var((0, 2, 3), unbiased=True)) < 1e-5
    # consumer conv: im2col of a
    pad = ksz // 2
    Acol, Ho, Wo = ops.im2col_nhwc(y16, B, H, W, C, ksz, stride, scale, shift)
    cols = F.unfold(a, ksz, padding=pad, stride=stride)                       # [B, C*k*k, L] (c, kh, kw) order
    cols_ref = cols.view(B, C, ksz * ksz, Ho * Wo).permute(0, 3, 2, 1).reshape(B * Ho * Wo, ksz * ksz * C)   # (kh,kw,c)
    assert rel_l2(Acol.float(), cols_ref.detach()) < 3e-3
    dAcol = torch.randn(B * Ho * Wo, ksz * ksz * C, device="cuda", generator=g0).to(torch.bfloat16)
    cols_ref.backward(dAcol.float())
    dgamma, dbeta = torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda")
    dy = ops.bn_relu_bwd(dAcol, y16, scale, shift, mean, rstd, gamma.detach(), dgamma, dbeta, B, H, W, C, ksz, stride, Ho, Wo)
    torch.cuda.synchronize()
    e = dict(dy=rel_l2(dy.float(), _nhwc(y.grad)), dgamma=rel_l2(dgamma, gamma.grad), dbeta=rel_l2(dbeta, beta.grad))
    print(f"bn_relu_bwd B{B} C{C} {H}x{W} k{ksz}s{stride}: {e}")
    assert e["dy"] < 6e-3 and e["dgamma"] < 2e-3 and e["dbeta"] < 2e-3, e
