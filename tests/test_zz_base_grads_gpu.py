"""-m gpu (runs late): full-size Prismer-BASE parity beyond the loss scalar (round-1 VERDICT weak #1b-d).

  * per-tensor parameter GRADIENTS of one caption fine-tune step (freeze_vision, BatchNorm batch statistics, dropout off) against
    autograd through the CPU oracle on the same fp32 weights -- cosine and rel-L2 per tensor, bounds per parameter family;
  * logits / encoder states against the oracle evaluated on the BF16 GRID (matrix weights, embeddings and inputs rounded to bf16 --
    what the engine's compute copies hold), i.e. the error that is left when weight rounding is taken out: activation storage in bf16;
  * the bench.py configuration itself (B = 32, T = 30, freeze_vision) with dropout off: loss vs the oracle.

Stated tolerances (bf16 storage of activations and weights, fp32 accumulation; measured on B200, bounds ~2x measured) are asserted
AND printed; BASELINE.md section 4 carries the table."""
import random

import pytest
import torch

from prismer_b200 import synthetic
from tests.helpers import rel_l2

pytestmark = pytest.mark.gpu
EXPERTS = synthetic.DEFAULT_EXPERTS
CFG = {"experts": EXPERTS, "prismer_model": "prismer_base", "image_resolution": 224, "freeze": "freeze_vision"}


def _family(name):
    if name.endswith((".self.query.weight", ".self.query.bias", ".self.key.weight")) and "text_decoder" in name:
        return "decoder q/k proj"
    if "conv1." in name:
        return "stems"
    if "resampler" in name:
        return "resampler"
    if "expert_encoder" in name:
        return "vit-adaptors/pos"
    if "embeddings" in name or "lm_head" in name:
        return "embeddings/head"
    return "decoder"


# rel-L2 / cosine bounds per family.  Stems: ReLU-mask flips at the BatchNorm threshold between any two implementations put noise-like
# differences into gradient sums over 10^5..10^6 positions (tests/test_bn_kernels_gpu.py shows the kernels are exact given identical
# inputs), so the bound there is on the cosine.
# Decoder query / key projections: their gradient is a second-order-small difference (dS = P * (dP - delta)) over T <= 30 keys, the
# noisiest tensors of the model in bf16 (BASE: 4.7e-2; LARGE at B = 2, T = 12: 1.9e-1 / cosine 0.984 on the output layer's query weight).
BOUNDS = {"decoder q/k proj": (2.5e-1, 0.97), "decoder": (6e-2, 0.998), "embeddings/head": (6e-2, 0.998), "vit-adaptors/pos": (6e-2, 0.998), "resampler": (6e-2, 0.998),
          "stems": (2.5e-1, 0.97)}


@pytest.mark.parametrize("name,cfg,patch,heads,B,T,min_tensors", [
    ("BASE", CFG, 16, 12, 2, 30, 150),
    ("LARGE", {"experts": EXPERTS, "prismer_model": "prismer_large", "image_resolution": 224, "freeze": "freeze_lang_vision"}, 14, 16, 2, 12, 150),
])
def test_gradients_match_oracle_autograd(name, cfg, patch, heads, B, T, min_tensors):
    """BASE (freeze_vision, BASELINE config 3) and LARGE (freeze_lang_vision, BASELINE config 5: ViT-L/14, S = 320, 24 + 1 decoder layers)."""
    from oracle import prismer_oracle as O
    from prismer_b200 import engine
    from prismer_b200.prismer_caption import PrismerCaption
    torch.manual_seed(1)
    m = PrismerCaption(cfg)
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    train_keys = [n for n, p in m.named_parameters() if p.requires_grad]
    m.cuda()
    ex = synthetic.synth_experts(B, 224, EXPERTS, 224, 11)
    ids, mask = synthetic.synth_tokens(B, T, 50265, 11, ragged=True)
    labels = ids.masked_fill(ids == 1, -100); labels[:, :4] = -100
    m.train(); m.text_decoder.eval()                     # BatchNorm batch statistics, dropout off
    random.seed(4)
    loss = engine.train_loss(m, synthetic.experts_to(ex, "cuda"), ids.cuda(), mask.cuda(), labels.cuda())
    loss.backward()
    torch.cuda.synchronize()
    got = {n: p.grad.detach().float().cpu() for n, p in m.named_parameters() if p.requires_grad}
    # oracle autograd on the same fp32 weights
    for k in train_keys:
        sd[k].requires_grad_(True)
    random.seed(4)
    ref, _, _ = O.caption_train_loss(ex, ids, mask, 4, sd, patch, heads, training_bn=True)
    ref.backward()
    assert abs(float(loss) - float(ref)) / abs(float(ref)) < 5e-3
    worst, bad = {}, []
    n_checked = n_zero = 0
    for k in train_keys:
        if k in ("text_decoder.lm_head.decoder.weight", "text_decoder.lm_head.decoder.bias"):
            continue
        if k.endswith(".key.bias"):
            # softmax is invariant to a per-query constant: q.(k_j + b) = q.k_j + q.b for every key j, so d(loss)/d(key.bias) is exactly 0
            # in exact arithmetic -- both sides hold rounding noise only (reference ~1e-9 of the weight gradients); nothing to compare
            n_zero += 1
            continue
        g_ref = sd[k].grad
        if g_ref is None or float(g_ref.norm()) == 0.0:
            continue
        g = got[k]
        r = rel_l2(g, g_ref)
        c = float(torch.nn.functional.cosine_similarity(g.double().flatten(), g_ref.double().flatten(), dim=0))
        fam = _family(k)
        w = worst.setdefault(fam, [0.0, 1.0, "", 0])
        if r > w[0]:
            w[0], w[2] = r, k
        w[1] = min(w[1], c)
        w[3] += 1
        n_checked += 1
        if not (r < BOUNDS[fam][0] and c > BOUNDS[fam][1]):
            bad.append((k, fam, round(r, 4), round(c, 5)))
    for fam, (r, c, k, n) in sorted(worst.items()):
        print(f"{name} grads [{fam:16s}] {n:3d} tensors: worst rel-L2 {r:.2e} ({k}), min cosine {c:.5f}  (bounds {BOUNDS[fam]})")
    print(f"{name} grads: {n_checked} tensors compared, {n_zero} key.bias tensors skipped (analytically zero gradient)")
    assert not bad, bad
    assert n_checked >= min_tensors, n_checked          # hundreds of trainable tensors; none silently skipped


def test_base_logits_vs_oracle_on_the_bf16_grid():
    """north_star: 'forward logits within 1e-3 rel of the reference on identical weights/inputs'.  With bf16 STORAGE of activations
    (one rounding = 2^-9 = 2e-3 relative) that bar is not reachable end to end; this test measures and pins what is: the error against
    the fp32 oracle on fp32 weights, and against the oracle on bf16-rounded weights / inputs, printed for BASELINE.md."""
    from oracle import prismer_oracle as O
    from prismer_b200.prismer_caption import PrismerCaption
    torch.manual_seed(1)
    m = PrismerCaption(CFG)
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    m.cuda().eval()
    B, T = 2, 30
    ex = synthetic.synth_experts(B, 224, EXPERTS, 224, 11)
    ids, mask = synthetic.synth_tokens(B, T, 50265, 11, ragged=True)
    random.seed(3)
    with torch.no_grad():
        enc = m.expert_encoder(synthetic.experts_to(ex, "cuda"))
        out = m.text_decoder(ids.cuda(), attention_mask=mask.cuda(), encoder_hidden_states=enc.transpose(0, 1))
    grid = lambda t: t.to(torch.bfloat16).float()
    sd16 = {k: (grid(v) if v.dtype.is_floating_point and v.dim() >= 2 else v) for k, v in sd.items()}
    ex16 = {k: ({kk: (grid(vv) if vv.dtype.is_floating_point else vv) for kk, vv in v.items()} if isinstance(v, dict) else grid(v))
            for k, v in ex.items()}
    res = {}
    for tag, s, e in (("fp32", sd, ex), ("bf16-grid", sd16, ex16)):
        esd, dsd = O.split_state_dict(s)
        random.seed(3)
        with torch.no_grad():
            enc_ref = O.encoder_forward(e, esd, 16)
            logits_ref, _ = O.decoder_forward(ids, mask, enc_ref.transpose(0, 1), dsd, 12)
        lg = out.logits.cpu()
        res[tag] = (rel_l2(enc.float().cpu(), enc_ref), rel_l2(lg, logits_ref), float((lg - logits_ref).abs().max()),
                    float(logits_ref.abs().mean()))
        top2 = logits_ref.topk(2, dim=-1).values
        decisive = (top2[..., 0] - top2[..., 1]) > 0.25
        agree = lg.argmax(-1) == logits_ref.argmax(-1)
        print(f"BASE eval vs oracle[{tag:9s}]: enc rel-L2 {res[tag][0]:.2e}, logits rel-L2 {res[tag][1]:.2e}, max |dlogit| {res[tag][2]:.3f} "
              f"(mean |logit| {res[tag][3]:.3f}); argmax equal at {int(agree.sum())}/{agree.numel()} positions, asserted on the "
              f"{int(decisive.sum())} with fp32 top-1/top-2 margin > 0.25")
        assert int(decisive.sum()) >= 10, "too few decisive positions for the token-id check to mean anything"
        assert bool(agree[decisive].all())
    assert res["fp32"][1] < 4e-2 and res["bf16-grid"][1] < 4e-2 and res["bf16-grid"][0] < 3e-2


def test_bench_config_loss_matches_oracle():
    """bench.py's workload (per-GPU batch 32, T = 30, 6 experts, freeze_vision) with dropout off: the loss the timed step computes
    is the reference's loss (oracle port on the host cores; ~10 s)."""
    from oracle import prismer_oracle as O
    from prismer_b200 import engine
    from prismer_b200.prismer_caption import PrismerCaption
    torch.manual_seed(0)
    m = PrismerCaption(CFG)
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    m.cuda()
    B, T = 32, 30
    ex = synthetic.synth_experts(B, 224, EXPERTS, 224, 1000)
    ids, mask = synthetic.synth_tokens(B, T, 50265, 1000)
    labels = ids.masked_fill(ids == 1, -100); labels[:, :4] = -100
    m.train(); m.text_decoder.eval()
    random.seed(5)
    with torch.no_grad():
        loss = engine.train_loss(m, synthetic.experts_to(ex, "cuda"), ids.cuda(), mask.cuda(), labels.cuda())
    random.seed(5)
    torch.set_num_threads(min(32, torch.get_num_threads()))
    with torch.no_grad():
        ref, _, _ = O.caption_train_loss(ex, ids, mask, 4, sd, 16, 12, training_bn=True)
    err = abs(float(loss) - float(ref)) / abs(float(ref))
    print(f"bench config (B=32): cuda loss {float(loss):.4f} oracle {float(ref):.4f} rel {err:.2e}")
    assert err < 2e-3
