"""``Prismer`` base class -- same constructor contract as ``model/prismer.py:15-37`` (config keys ``experts``,
``prismer_model``, ``image_resolution``, ``freeze``), same attributes (``tokenizer``, ``expert_encoder``,
``text_decoder``, ``ignored_modules``) and the same freeze policy; built without network access."""
import json
import os

import torch.nn as nn

from .modules.roberta import RobertaConfig, load_decoder
from .modules.vit import load_encoder
from .tokenizer import build_tokenizer

_CONFIG_JSON = os.path.join(os.path.dirname(os.path.abspath(__file__)), "configs", "prismer.json")


def expert_channels(experts):
    """model/prismer.py:18-27 (note: the string 'none' iterates as characters and matches nothing -> PrismerZ)."""
    out = {"rgb": 3}
    for exp in experts:
        if exp in ("depth", "edge"):
            out[exp] = 1
        elif exp in ("normal",):
            out[exp] = 3
        elif "seg" in exp:
            out["seg"] = 64
        elif exp in ("obj_detection", "ocr_detection"):
            out[exp] = 64
    return out


class Prismer(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.experts = expert_channels(config["experts"])
        pc = config.get("prismer_config")
        if pc is None:
            path = "configs/prismer.json" if os.path.exists("configs/prismer.json") else _CONFIG_JSON
            pc = json.load(open(path, "r"))[config["prismer_model"]]
        roberta_config = RobertaConfig.from_dict(pc["roberta_model"])
        self.tokenizer = build_tokenizer(pc["roberta_model"]["model_name"], roberta_config.vocab_size)
        self.expert_encoder = load_encoder(pc["vit_model"], experts=self.experts, image_resolution=config["image_resolution"]) \
            if "vit_dims" not in pc else _custom_encoder(pc, self.experts, config["image_resolution"])
        self.text_decoder = load_decoder(pc["roberta_model"]["model_name"], config=roberta_config)
        self.prepare_to_train(config["freeze"])
        self.ignored_modules = self.get_ignored_modules(config["freeze"])

    def prepare_to_train(self, mode="none"):
        """model/prismer.py:39-59 -- name-matched freeze policy."""
        lang = lambda n: "encoder.layer" in n and all(k not in n for k in ["1.self", "1.output", "adaptor"])
        vis = lambda n: "transformer.resblocks" in n and "adaptor" not in n
        for name, p in self.named_parameters():
            if mode == "freeze_lang":
                p.requires_grad = not lang(name)
            elif mode == "freeze_vision":
                p.requires_grad = not vis(name)
            elif mode == "freeze_lang_vision":
                p.requires_grad = not (lang(name) or vis(name))
            else:
                p.requires_grad = True

    def get_ignored_modules(self, mode="none"):
        """model/prismer.py:61-94 (FSDP ignore list; kept for API compatibility)."""
        mods = []
        if mode in ("freeze_lang", "freeze_lang_vision"):
            for layer in self.text_decoder.roberta.encoder.layer:
                mods += [layer[0].attention, layer[0].intermediate, layer[0].output]
        if mode in ("freeze_vision", "freeze_lang_vision"):
            for blk in self.expert_encoder.transformer.resblocks:
                mods += [blk[0].attn, blk[0].mlp, blk[0].ln_1, blk[0].ln_2]
        return mods if mode in ("freeze_lang", "freeze_vision", "freeze_lang_vision") else None


def _custom_encoder(pc, experts, image_resolution):
    from .modules.vit import VisionTransformer
    patch, width, layers = pc["vit_dims"]
    return VisionTransformer(image_resolution, patch, width, layers, width // 64, experts)
