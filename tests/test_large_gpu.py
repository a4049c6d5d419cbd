"""-m gpu: Prismer-LARGE dimensions (ViT-L/14: D=1024, 24 layers, patch 14 -> bilinear 224->256/64 stems, resampler head dim 128;
roberta-large: H=1024, 24+1 layers, 16 heads) run forward + backward + one fused optimizer step and agree with the CPU oracle on
the loss (BASELINE.json configs 4/5 use this model family)."""
import random

import pytest
import torch

from prismer_b200 import synthetic

pytestmark = pytest.mark.gpu


def test_prismer_large_forward_backward_matches_oracle_loss():
    from oracle import prismer_oracle as O
    from prismer_b200 import engine
    from prismer_b200.optim import FusedAdamW
    from prismer_b200.prismer_caption import PrismerCaption
    experts = ["depth", "normal", "seg_coco", "edge", "obj_detection", "ocr_detection"]
    torch.manual_seed(0)
    m = PrismerCaption({"experts": experts, "prismer_model": "prismer_large", "image_resolution": 224, "freeze": "freeze_lang_vision"})
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    m.cuda().train()
    m.text_decoder.eval()                      # dropout off for the oracle comparison; BatchNorm on batch statistics
    B, T = 2, 12
    ex = synthetic.synth_experts(B, 224, experts, 224, 3)
    ids, mask = synthetic.synth_tokens(B, T, 50265, 3, ragged=True)
    labels = ids.masked_fill(ids == 1, -100); labels[:, :4] = -100
    opt = FusedAdamW(m, lr=1e-4, weight_decay=0.05)
    random.seed(5)
    loss = engine.train_loss(m, synthetic.experts_to(ex, "cuda"), ids.cuda(), mask.cuda(), labels.cuda())
    loss.backward()
    opt.step()
    torch.cuda.synchronize()
    st = engine._store(m)
    assert torch.isfinite(st.grad_t).all() and float(st.grad_t.abs().sum()) > 0
    frozen = [n for n, p in m.named_parameters() if not p.requires_grad]
    assert any("encoder.layer.3.0.attention" in n for n in frozen) and any("transformer.resblocks.5.0.mlp" in n for n in frozen)
    random.seed(5)
    with torch.no_grad():
        ref, _, _ = O.caption_train_loss(ex, ids, mask, 4, sd, 14, 16, training_bn=True)
    err = abs(float(loss) - float(ref)) / abs(float(ref))
    print(f"LARGE: cuda loss {float(loss):.4f} oracle {float(ref):.4f} rel {err:.2e}")
    assert err < 5e-3
