"""Construction helpers: tiny / custom-width models for tests and the reference-layout state_dict template."""
from collections import OrderedDict

import torch

from .modules.roberta import RobertaConfig, RobertaForCausalLMModified
from .modules.vit import VisionTransformer
from .prismer import expert_channels


def build_encoder(width, layers, patch, res, experts):
    return VisionTransformer(res, patch, width, layers, width // 64, expert_channels(experts))


def build_decoder(dec_cfg):
    return RobertaForCausalLMModified(RobertaConfig.from_dict(dec_cfg))


def template_state_dict(width, layers, patch, res, experts, dec_cfg=None) -> "OrderedDict[str, torch.Tensor]":
    """Names / shapes / dtypes of the reference-layout checkpoint (``expert_encoder.*`` + ``text_decoder.*``), built on
    the meta device (no memory)."""
    with torch.device("meta"):
        sd = OrderedDict(("expert_encoder." + k, v) for k, v in build_encoder(width, layers, patch, res, experts).state_dict().items())
        if dec_cfg is not None:
            sd.update(("text_decoder." + k, v) for k, v in build_decoder(dec_cfg).state_dict().items())
    return sd
