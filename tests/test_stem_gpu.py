"""-m gpu: conv stems (resample -> im2col -> GEMM -> BatchNorm(batch stats) -> ReLU ... -> 1x1) forward and backward
against plain fp32 PyTorch (F.conv2d / F.batch_norm / F.interpolate), i.e. the ops of vit.py:88-120."""
import pytest
import torch
import torch.nn.functional as F

from tests.helpers import rel_l2

pytestmark = pytest.mark.gpu


def _rt(t):
    """round to bf16 in the forward, identity in the backward: puts the fp32 reference on the product's storage grid so
    that ReLU masks agree (a flipped mask element changes noise-like gradient sums such as dbeta by O(sqrt(flips/M)))."""
    return t + (t.bfloat16().float() - t).detach()


ws_dbg, ws_act = [], []


def _ref_stem(stem, x, training):
    if stem.scale_factor != 1.0:
        x = F.interpolate(x, scale_factor=stem.scale_factor, mode="bilinear", align_corners=True)
    x = _rt(x)
    ws = []
    ws_dbg.clear(); ws_act.clear()
    for i, s in enumerate(stem.strides):
        conv, bn = stem[str(1 + 3 * i)], stem[str(2 + 3 * i)]
        w = conv.weight.detach().clone().float().requires_grad_(True)
        g = bn.weight.detach().clone().requires_grad_(True)
        b = bn.bias.detach().clone().requires_grad_(True)
        ws += [w, g, b]
        x = _rt(F.conv2d(x, _rt(w), stride=s, padding=1))
        x.retain_grad(); ws_dbg.append(x)
        x = F.batch_norm(x, bn.running_mean.clone(), bn.running_var.clone(), g, b, training, 0.1, 1e-5)
        x = _rt(torch.relu(x))
        x.retain_grad(); ws_act.append(x)
    w13 = stem["13"].weight.detach().clone().float().requires_grad_(True)
    ws.append(w13)
    return F.conv2d(x, _rt(w13)), ws


@pytest.mark.parametrize("domain,B,size,patch", [("depth", 2, 64, 16), ("depth", 8, 128, 16), ("normal", 4, 128, 16),
                                                 ("seg_coco", 8, 224, 16), ("ocr_detection", 4, 112, 14), ("edge", 2, 112, 14)])
def test_stem_fwd_bwd(domain, B, size, patch):
    from prismer_b200 import engine, modeling, synthetic
    vit = modeling.build_encoder(256, 1, patch, 64 if patch == 16 else 56, [domain])
    vit.load_state_dict(synthetic.synth_state_dict(vit.state_dict(), 5))
    vit.cuda().train()
    engine.prepare(vit)
    key = "seg" if "seg" in domain else domain
    stem = vit.conv1[key]
    cin = stem["1"].weight.shape[1]
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn(B, cin, size, size, device="cuda", generator=g)
    ref, ws = _ref_stem(stem, x, True)
    tok, gh, gw, sv = engine._stem_fwd(stem, x, True, True)
    out = tok.view(B, gh, gw, -1).permute(0, 3, 1, 2).float()
    ferr = rel_l2(out, ref.detach())
    dtok = torch.randn(tok.shape, device="cuda", generator=g).to(torch.bfloat16)
    ref.backward(dtok.float().view(B, gh, gw, -1).permute(0, 3, 1, 2))
    st = engine._store(vit)
    st.zero_grad()
    sv.debug = True
    engine._stem_bwd(stem, sv, dtok)
    torch.cuda.synchronize()
    for i, L in enumerate(sv.layers):
        yref, aref = ws_dbg[i], ws_act[i]
        Bc, C, Hh, Ww = yref.shape
        to = lambda t: t.view(Bc, Hh, Ww, C).permute(0, 3, 1, 2).float()
        mask_mine = (to(L.y) * L.scale.view(1, -1, 1, 1) + L.shift.view(1, -1, 1, 1)) > 0
        mask_ref = aref > 0
        print(f"   layer{i}: y {rel_l2(to(L.y), yref.detach()):.1e} dy {rel_l2(to(L.dy), yref.grad):.1e} "
              f"mask-mismatch {(mask_mine != mask_ref).float().mean().item():.2e}"
              + (f" dA {rel_l2(to(L.dA), aref.grad):.1e}" if L.dA.shape[1] == C else ""))
    names = []
    for i in range(4):
        names += [f"{1 + 3 * i}.weight", f"{2 + 3 * i}.weight", f"{2 + 3 * i}.bias"]
    names.append("13.weight")
    params = dict(stem.named_parameters())
    errs = {n: rel_l2(params[n]._g32, w.grad) for n, w in zip(names, ws)}
    print(f"stem {domain} B={B} {size}px p{patch}: fwd {ferr:.2e} | " + " ".join(f"{n}:{e:.1e}" for n, e in errs.items()))
    # Residual after aligning the reference to the bf16 grid: the raw conv outputs of two fp32-accumulating implementations
    # differ by one bf16 ulp on ~0.2% of the elements; where such an element sits at the ReLU threshold the mask flips
    # (measured rate 1e-4..1e-3, printed above) and each flip injects an O(1) element error into noise-like gradients:
    # rel-L2 ~ sqrt(2 * flip_rate) = 2-5%.  The kernels themselves are exact to 2e-3 (tests/test_bn_kernels_gpu.py).
    assert ferr < 1.5e-2
    # with very few samples per channel in the last BatchNorm (B*gh*gw < 200) a single flip moves the statistics themselves,
    # so the bound is wider there; the large cases bound the systematic error
    assert max(errs.values()) < (0.15 if B * gh * gw < 200 else 9e-2), errs
