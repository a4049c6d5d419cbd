"""TEST INFRASTRUCTURE ONLY -- generates ``tests/golden/*.npz`` from the UNMODIFIED reference.

Run in the build container (needs ``/root/reference``):  ``python oracle/gen_golden.py``

The reference ships no golden vectors for the hot path, so the fixtures
are outputs of the reference's own modules (``model/modules/{vit,resampler,roberta}.py``) on
deterministic weights/inputs from ``prismer_b200.synthetic`` -- reproducible anywhere from
``(seed, name, shape)`` -- at a tiny width so the files stay small.  Each fixture stores the
config, the seeds and the reference outputs; weights and inputs are NOT stored.
"""
import os
import random
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import reference_shim  # noqa: E402
from prismer_b200 import synthetic  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")

TINY_DEC = {
    "attention_probs_dropout_prob": 0.1, "bos_token_id": 0, "eos_token_id": 2, "hidden_act": "gelu",
    "hidden_dropout_prob": 0.1, "hidden_size": 256, "vision_hidden_size": 256, "initializer_range": 0.02,
    "intermediate_size": 1024, "layer_norm_eps": 1e-05, "max_position_embeddings": 514,
    "model_name": "roberta-tiny", "num_attention_heads": 4, "num_hidden_layers": 2, "pad_token_id": 1,
    "type_vocab_size": 1, "vocab_size": 1000, "num_decoder_layers": 4, "is_decoder": True,
}


def expert_dict(names):
    """model/prismer.py:18-27"""
    d = {"rgb": 3}
    for e in names:
        if e in ("depth", "edge"):
            d[e] = 1
        elif e == "normal":
            d[e] = 3
        elif "seg" in e:
            d["seg"] = 64
        elif e in ("obj_detection", "ocr_detection"):
            d[e] = 64
    return d


def build(ns, width, layers, patch, res, experts, seed):
    vit = ns.VisionTransformer(res, patch, width, layers, width // 64, expert_dict(experts))
    dec = ns.build_decoder(TINY_DEC)
    sd = {}
    sd.update({"expert_encoder." + k: v for k, v in vit.state_dict().items()})
    sd.update({"text_decoder." + k: v for k, v in dec.state_dict().items()})
    sd = synthetic.synth_state_dict(sd, seed)
    vit.load_state_dict({k[len("expert_encoder."):]: v for k, v in sd.items() if k.startswith("expert_encoder.")})
    dec.load_state_dict({k[len("text_decoder."):]: v for k, v in sd.items() if k.startswith("text_decoder.")})
    assert dec.lm_head.decoder.weight.data_ptr() == dec.roberta.embeddings.word_embeddings.weight.data_ptr()
    return vit, dec, sd


def main():
    ns = reference_shim.load()
    torch.manual_seed(0)
    os.makedirs(GOLD, exist_ok=True)
    full = synthetic.DEFAULT_EXPERTS
    out = {}

    # ---- fixture A: Prismer-tiny (6 experts), patch 16 ---------------------------------------
    cfg = dict(width=256, layers=2, patch=16, res=64, label=64, B=2, T=8, seed=7, in_seed=11, py_seed=1234)
    vit, dec, sd = build(ns, cfg["width"], cfg["layers"], cfg["patch"], cfg["res"], full, cfg["seed"])
    ex = synthetic.synth_experts(cfg["B"], cfg["res"], full, cfg["label"], cfg["in_seed"])
    ids, mask = synthetic.synth_tokens(cfg["B"], cfg["T"], TINY_DEC["vocab_size"], cfg["in_seed"], ragged=True)
    vit.eval(); dec.eval()
    with torch.no_grad():
        random.seed(cfg["py_seed"])
        enc = vit(ex)                                     # [S,B,D]
        encb = enc.transpose(0, 1).contiguous()
        labels = ids.masked_fill(ids == 1, -100)
        labels[:, :3] = -100
        o = dec(ids, attention_mask=mask, encoder_hidden_states=encb, labels=labels, return_dict=True)
        prefix = ids[:, :4].clone()
        prefix[prefix == 1] = 5
        prefix[prefix == 2] = 6
        g1 = dec.generate(input_ids=prefix, encoder_hidden_states=encb, attention_mask=torch.ones_like(prefix),
                          num_beams=1, do_sample=False, max_length=12, min_length=8)
        g3 = dec.generate(input_ids=prefix, encoder_hidden_states=encb, attention_mask=torch.ones_like(prefix),
                          num_beams=3, max_length=12, min_length=8)
    out["A"] = dict(cfg=cfg, enc=enc.numpy(), logits=o.logits.numpy(), loss=o.loss.numpy(), ids=ids.numpy(),
                    mask=mask.numpy(), labels=labels.numpy(), prefix=prefix.numpy(), greedy=g1.numpy(), beam3=g3.numpy())

    # train-mode encoder (batch-stat BatchNorm) + gradients of the caption loss (eval-mode decoder: no dropout)
    vit.train(); dec.eval()
    random.seed(cfg["py_seed"])
    enc_t = vit(ex)
    o = dec(ids, attention_mask=mask, encoder_hidden_states=enc_t.transpose(0, 1), labels=labels, return_dict=True)
    loss = o.loss.mean()
    loss.backward()
    gnames = ["conv1.depth.1.weight", "conv1.seg.4.weight", "conv1.obj_detection.13.weight", "conv1.rgb.weight",
              "conv1.depth.2.weight", "conv1.depth.2.bias", "instance_embedding", "positional_embedding",
              "resampler.latents", "resampler.perceiver_blocks.0.attn.in_proj_weight",
              "resampler.perceiver_blocks.3.mlp.c_fc.weight", "resampler.perceiver_blocks.1.ln_2.weight",
              "transformer.resblocks.0.0.attn.in_proj_weight", "transformer.resblocks.1.1.adaptor.down_proj.weight",
              "transformer.resblocks.0.0.mlp.c_proj.bias", "ln_pre.weight", "ln_post.bias"]
    def _g(t):  # keep fixtures small: leading 2048 elements + the L2 norm of the full gradient
        return np.concatenate([[float(t.norm())], t.flatten()[:2048].numpy()]).astype(np.float32)
    grads = {"E." + n: _g(dict(vit.named_parameters())[n].grad) for n in gnames}
    dnames = ["roberta.embeddings.word_embeddings.weight", "roberta.embeddings.position_embeddings.weight",
              "roberta.encoder.layer.0.0.attention.self.query.weight", "roberta.encoder.layer.1.1.self.key.weight",
              "roberta.encoder.layer.0.2.adaptor.up_proj.weight", "roberta.encoder.layer.1.0.output.LayerNorm.weight",
              "roberta.encoder.output_layer.intermediate.dense.bias", "lm_head.dense.weight", "lm_head.bias"]
    grads.update({"D." + n: _g(dict(dec.named_parameters())[n].grad) for n in dnames})
    bn = {k: v.numpy() for k, v in vit.state_dict().items() if "running" in k and ("depth.2" in k or "seg.11" in k)}
    out["A_train"] = dict(enc=enc_t.detach().numpy(), loss=loss.detach().numpy(), **{"g." + k: v for k, v in grads.items()},
                          **{"bn." + k: v for k, v in bn.items()})

    # ---- fixture B: patch 14 (bilinear 56->64 / 56->16 resample, bicubic pos-emb identity) -----------
    cfgb = dict(width=256, layers=1, patch=14, res=56, label=56, B=2, seed=3, in_seed=5, py_seed=99)
    vitb, _, _ = build(ns, cfgb["width"], cfgb["layers"], cfgb["patch"], cfgb["res"], ["depth", "seg_coco", "obj_detection"], cfgb["seed"])
    exb = synthetic.synth_experts(cfgb["B"], cfgb["res"], ["depth", "seg_coco", "obj_detection"], cfgb["label"], cfgb["in_seed"])
    vitb.eval()
    with torch.no_grad():
        random.seed(cfgb["py_seed"])
        encb_ = vitb(exb)
    out["B"] = dict(cfg=cfgb, enc=encb_.numpy())

    # ---- fixture C: 112 px image with 64 px labels -> bicubic pos-emb interpolation 7x7 -> 4x4 -------
    cfgc = dict(width=256, layers=1, patch=16, res=112, label=64, B=1, seed=4, in_seed=6, py_seed=5)
    vitc, _, _ = build(ns, cfgc["width"], cfgc["layers"], cfgc["patch"], cfgc["res"], ["normal", "edge", "ocr_detection"], cfgc["seed"])
    exc = synthetic.synth_experts(cfgc["B"], cfgc["res"], ["normal", "edge", "ocr_detection"], cfgc["label"], cfgc["in_seed"])
    vitc.eval()
    with torch.no_grad():
        random.seed(cfgc["py_seed"])
        encc = vitc(exc)
    out["C"] = dict(cfg=cfgc, enc=encc.numpy())

    # ---- fixture Z: PrismerZ (experts 'none' -> rgb only, no resampler), BASELINE config 1 shape ------
    cfgz = dict(width=256, layers=2, patch=16, res=64, B=1, seed=9, in_seed=2)
    vitz, decz, _ = build(ns, cfgz["width"], cfgz["layers"], cfgz["patch"], cfgz["res"], [], cfgz["seed"])
    exz = synthetic.synth_experts(1, cfgz["res"], [], 64, cfgz["in_seed"])
    vitz.eval(); decz.eval()
    with torch.no_grad():
        encz = vitz(exz)
        pz = torch.tensor([[0, 250, 217, 9]])
        gz = decz.generate(input_ids=pz, encoder_hidden_states=encz.transpose(0, 1).contiguous(),
                           attention_mask=torch.ones_like(pz), num_beams=1, do_sample=False, max_length=20, min_length=8)
    out["Z"] = dict(cfg=cfgz, enc=encz.numpy(), prefix=pz.numpy(), greedy=gz.numpy())

    for name, d in out.items():
        flat = {}
        for k, v in d.items():
            if k == "cfg":
                for ck, cv in v.items():
                    flat["cfg." + ck] = np.asarray(cv)
            else:
                flat[k] = np.asarray(v)
        np.savez_compressed(os.path.join(GOLD, f"prismer_tiny_{name}.npz"), **flat)
        print(name, {k: v.shape for k, v in flat.items() if not k.startswith("cfg.")})


if __name__ == "__main__":
    main()
