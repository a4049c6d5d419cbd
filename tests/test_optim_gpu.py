"""-m gpu: fused AdamW kernel vs torch.optim.AdamW (same update rule as train_caption.py:111-112)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_fused_adamw_matches_torch():
    from prismer_b200 import ops
    n = 1_000_003 // 4 * 4
    g0 = torch.Generator(device="cuda").manual_seed(0)
    p = torch.randn(n, device="cuda", generator=g0)
    ref = torch.nn.Parameter(p.clone())
    opt = torch.optim.AdamW([ref], lr=5e-5, weight_decay=0.05)
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    p16 = torch.empty(n, dtype=torch.bfloat16, device="cuda")
    for step in range(1, 4):
        g = torch.randn(n, device="cuda", generator=g0)
        ref.grad = g.clone() * 0.5
        opt.step()
        ops.adamw_step(p, g, m, v, p16, 5e-5, 0.9, 0.999, 1e-8, 0.05, step, grad_scale=0.5)
    torch.cuda.synchronize()
    assert torch.allclose(p, ref.data, rtol=1e-5, atol=1e-7)
    assert torch.equal(p16, p.to(torch.bfloat16))
