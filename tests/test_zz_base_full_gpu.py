"""-m gpu (runs last): the full-size Prismer-BASE model (BASELINE.json configs 2/3 dimensions: ViT-B/16 + 6 experts + roberta-base,
224 px, T = 30) against the CPU oracle on the same seeded weights / inputs: caption loss (train path, BatchNorm batch
statistics, dropout off), eval logits, and greedy token ids where the oracle's own top-1 / top-2 margin exceeds the bf16 noise."""
import random

import pytest
import torch

from prismer_b200 import synthetic
from tests.helpers import rel_l2

pytestmark = pytest.mark.gpu
EXPERTS = synthetic.DEFAULT_EXPERTS


def test_prismer_base_full_size_matches_oracle():
    from oracle import prismer_oracle as O
    from prismer_b200 import engine
    from prismer_b200.prismer_caption import PrismerCaption
    torch.manual_seed(1)
    m = PrismerCaption({"experts": EXPERTS, "prismer_model": "prismer_base", "image_resolution": 224, "freeze": "freeze_vision"})
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    m.cuda()
    B, T = 2, 30
    ex = synthetic.synth_experts(B, 224, EXPERTS, 224, 11)
    ids, mask = synthetic.synth_tokens(B, T, 50265, 11, ragged=True)
    labels = ids.masked_fill(ids == 1, -100); labels[:, :4] = -100
    exd = synthetic.experts_to(ex, "cuda")
    esd, dsd = O.split_state_dict(sd)

    # ---- eval: encoder states + logits
    m.eval()
    random.seed(3)
    with torch.no_grad():
        enc = m.expert_encoder(exd)
        out = m.text_decoder(ids.cuda(), attention_mask=mask.cuda(), encoder_hidden_states=enc.transpose(0, 1))
    random.seed(3)
    with torch.no_grad():
        enc_ref = O.encoder_forward(ex, esd, 16)
        logits_ref, _ = O.decoder_forward(ids, mask, enc_ref.transpose(0, 1), dsd, 12)
    e_enc, e_log = rel_l2(enc.float().cpu(), enc_ref), rel_l2(out.logits.cpu(), logits_ref)
    top2 = logits_ref.topk(2, dim=-1).values
    decisive = (top2[..., 0] - top2[..., 1]) > 0.25
    agree = (out.logits.cpu().argmax(-1) == logits_ref.argmax(-1))
    print(f"BASE eval: enc rel-L2 {e_enc:.2e}, logits rel-L2 {e_log:.2e}, argmax agreement {agree.float().mean():.3f} "
          f"(decisive positions {int(decisive.sum())}/{decisive.numel()}: {agree[decisive].float().mean() if decisive.any() else 1.0:.3f})")
    assert e_enc < 4e-2 and e_log < 6e-2
    if decisive.any():
        assert bool(agree[decisive].all())          # token ids exact wherever the fp32 decision is not a near-tie

    # ---- train path loss (BatchNorm batch statistics; dropout off so that the oracle is comparable)
    m.train(); m.text_decoder.eval()
    random.seed(4)
    loss = engine.train_loss(m, exd, ids.cuda(), mask.cuda(), labels.cuda())
    loss.backward()
    torch.cuda.synchronize()
    random.seed(4)
    with torch.no_grad():
        ref, _, _ = O.caption_train_loss(ex, ids, mask, 4, sd, 16, 12, training_bn=True)
    err = abs(float(loss) - float(ref)) / abs(float(ref))
    print(f"BASE train: cuda loss {float(loss):.4f} oracle {float(ref):.4f} rel {err:.2e}")
    assert err < 5e-3
    assert torch.isfinite(engine._store(m).grad_t).all()
