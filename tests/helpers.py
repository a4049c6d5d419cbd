"""Shared helpers for the test-suite (test infrastructure; may import ``oracle``)."""
import os
import random

import numpy as np
import torch

from prismer_b200 import synthetic

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")

TINY_DEC = {
    "attention_probs_dropout_prob": 0.1, "bos_token_id": 0, "eos_token_id": 2, "hidden_act": "gelu",
    "hidden_dropout_prob": 0.1, "hidden_size": 256, "vision_hidden_size": 256, "initializer_range": 0.02,
    "intermediate_size": 1024, "layer_norm_eps": 1e-05, "max_position_embeddings": 514,
    "model_name": "roberta-tiny", "num_attention_heads": 4, "num_hidden_layers": 2, "pad_token_id": 1,
    "type_vocab_size": 1, "vocab_size": 1000, "num_decoder_layers": 4, "is_decoder": True,
}


def load_golden(name):
    z = np.load(os.path.join(GOLD, f"prismer_tiny_{name}.npz"))
    cfg = {k[4:]: z[k].item() for k in z.files if k.startswith("cfg.")}
    return cfg, {k: z[k] for k in z.files if not k.startswith("cfg.")}


def rel_l2(a, b):
    a = torch.as_tensor(a).double().flatten()
    b = torch.as_tensor(b).double().flatten()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


class Holder(torch.nn.Module):
    """expert_encoder + text_decoder pair with the reference's attribute names (a config-free ``Prismer``)."""

    def __init__(self, enc, dec=None):
        super().__init__()
        self.expert_encoder = enc
        if dec is not None:
            self.text_decoder = dec


def build_model(width, layers, patch, res, experts, dec_cfg, seed, device="cuda", freeze=None):
    from prismer_b200 import modeling
    enc = modeling.build_encoder(width, layers, patch, res, experts)
    dec = modeling.build_decoder(dec_cfg) if dec_cfg is not None else None
    m = Holder(enc, dec)
    sd = synthetic.synth_state_dict(m.state_dict(), seed)
    m.load_state_dict(sd)
    if freeze == "freeze_vision":
        for n, p in m.named_parameters():
            p.requires_grad = not ("transformer.resblocks" in n and "adaptor" not in n)
    m.to(device)
    return m, sd
