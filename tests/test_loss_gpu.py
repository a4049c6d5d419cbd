"""Label-smoothed LM loss kernels (csrc/embed_loss.cu) against fp32 torch: the shifted cross entropy of roberta.py:379-387
(``CrossEntropyLoss(reduction='none', label_smoothing=0.1)`` on logits[:, :-1] / labels[:, 1:], summed per sample) and the
weighted batch mean of prismer_caption.py:33 / prismer_vqa.py:40-41; forward value and the bf16 dlogits fed to the GEMMs."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _reference(logits, labels, V, weights, smoothing):
    B, T = labels.shape
    x = logits[:, :V].float().view(B, T, V).clone().requires_grad_(True)
    tok = torch.nn.functional.cross_entropy(x[:, :-1].reshape(-1, V), labels[:, 1:].reshape(-1), reduction="none",
                                            label_smoothing=smoothing, ignore_index=-100)
    sample = tok.view(B, T - 1).sum(1)
    mean = ((weights if weights is not None else 1.0) * sample).sum() / B
    mean.backward()
    return mean.detach(), sample.detach(), x.grad.view(B * T, V)


# (B, T, V, ld): BASE vocabulary with its padded leading dimension (128-bit path, 1 tail column); a vocabulary that is a multiple of 4;
# 3 tail columns; an odd leading dimension (scalar path)
@pytest.mark.parametrize("B,T,V,ld", [(4, 9, 50265, 50272), (3, 5, 1000, 1000), (2, 6, 1003, 1008), (2, 4, 517, 519)])
@pytest.mark.parametrize("weighted", [False, True])
def test_ce_loss_fwd_bwd(B, T, V, ld, weighted):
    from prismer_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(B * 1000 + V)
    logits = torch.randn(B * T, ld, device="cuda", generator=g) * 3.0
    logits[:, V:] = 1e4                                             # padding columns must not be read
    labels = torch.randint(0, V, (B, T), device="cuda", generator=g)
    labels[0, 2] = -100                                             # ignored position (prompt / padding, prismer_caption.py:27-30)
    labels[-1, -1] = -100
    weights = (torch.rand(B, device="cuda", generator=g) + 0.5) if weighted else None
    mean, sample, lse = ops.ce_loss_fwd(logits, labels, V, weights=weights, smoothing=0.1)
    d = ops.ce_loss_bwd(logits, labels, lse, V, weights=weights, smoothing=0.1)
    rmean, rsample, rgrad = _reference(logits, labels, V, weights, 0.1)
    assert torch.allclose(sample, rsample, rtol=1e-5, atol=1e-4), (sample - rsample).abs().max()
    assert torch.allclose(mean[0], rmean, rtol=1e-5, atol=1e-4)
    ref_lse = torch.logsumexp(logits[:, :V].float(), dim=1)
    assert torch.allclose(lse, ref_lse, rtol=2e-6, atol=2e-5)
    assert d.shape == (B * T, ld) and (ld == V or float(d[:, V:].float().abs().max()) == 0.0)          # padded columns feed the GEMMs as zeros
    got = d[:, :V].float()
    err = (got - rgrad).abs().max() / rgrad.abs().max()
    assert float(err) < 5e-3, float(err)                            # bf16 rounding of the stored gradient
    rows = torch.arange(B * T, device="cuda").view(B, T)
    assert float(got[rows[:, -1]].abs().max()) == 0.0               # the last position of every sample has no target
