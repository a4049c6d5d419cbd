"""Vision encoder -- parameter containers mirroring ``model/modules/vit.py`` (state_dict keys / shapes identical,
SURVEY.md section 8b); the forward / backward arithmetic lives in ``prismer_b200.engine`` (sm_100a kernels)."""
import torch
import torch.nn as nn

from .resampler import PerceiverResampler, _MLP
from .utils import Adaptor, LayerNorm, QuickGELU

# (in-channel multiplier table) per-layer strides of the two stem families, vit.py:88-120
STEM_STRIDES = {True: (2, 2, 1, 1), False: (2, 2, 2, 2)}   # key: 64-channel label-map stem?
WIDE_STEMS = ("seg", "obj_detection", "ocr_detection")


class ResidualAttentionBlock(nn.Module):
    def __init__(self, d_model: int, n_head: int):
        super().__init__()
        self.n_head = n_head
        self.attn = nn.MultiheadAttention(d_model, n_head)
        self.mlp = _MLP(d_model, QuickGELU())
        self.ln_1, self.ln_2 = LayerNorm(d_model), LayerNorm(d_model)


class Transformer(nn.Module):
    def __init__(self, width: int, layers: int, heads: int):
        super().__init__()
        self.resblocks = nn.Sequential(*[nn.ModuleList([ResidualAttentionBlock(width, heads), Adaptor(width)])
                                         for _ in range(layers)])


def _make_stem(in_ch: int, width: int, patch_size: int, wide: bool) -> nn.Module:
    """Keys follow the reference's nn.Sequential indices: 0 resample, 1/4/7/10 conv3x3, 2/5/8/11 BN, 13 conv1x1."""
    chans = [in_ch, width // 8, width // 4, width // 2, width]
    mods = {}
    for i, s in enumerate(STEM_STRIDES[wide]):
        mods[str(1 + 3 * i)] = nn.Conv2d(chans[i], chans[i + 1], kernel_size=3, stride=s, padding=1, bias=False)
        mods[str(2 + 3 * i)] = nn.BatchNorm2d(chans[i + 1])
    mods["13"] = nn.Conv2d(width, width, kernel_size=1, bias=False)
    stem = nn.ModuleDict(mods)
    stem.scale_factor = (4 if wide else 16) / patch_size
    stem.strides = STEM_STRIDES[wide]
    return stem


class VisionTransformer(nn.Module):
    def __init__(self, input_resolution: int, patch_size: int, width: int, layers: int, heads: int, experts: dict):
        super().__init__()
        self.experts = experts
        self.patch_size, self.width, self.heads, self.input_resolution = patch_size, width, heads, input_resolution
        self.conv1 = nn.ModuleDict()
        for e, ch in experts.items():
            if e == "rgb":
                self.conv1[e] = nn.Conv2d(ch, width, kernel_size=patch_size, stride=patch_size, bias=False)
            else:
                self.conv1[e] = _make_stem(64 if e in WIDE_STEMS else ch, width, patch_size, e in WIDE_STEMS)
        scale = width ** -0.5
        self.positional_embedding = nn.Parameter(scale * torch.randn((input_resolution // patch_size) ** 2, width))
        if "obj_detection" in experts:
            self.instance_embedding = nn.Parameter(scale * torch.randn(128, width))
        self.transformer = Transformer(width, layers, heads)
        if len(experts) > 1:
            self.resampler = PerceiverResampler(width=width, layers=4, heads=8, num_latents=64)
        self.ln_pre, self.ln_post = LayerNorm(width), LayerNorm(width)

    def forward(self, x: dict):
        """dict of expert tensors -> [S, B, D] seq-first, as ``vit.py:133-172``."""
        from .. import engine
        return engine.encoder_apply(self, x)


VIT_CONFIGS = {  # name -> (patch, width, layers); heads = width // 64 (vit.py:211-214)
    "ViT-B/16": (16, 768, 12), "ViT-B/32": (32, 768, 12), "ViT-L/14": (14, 1024, 24), "ViT-L/14@336px": (14, 1024, 24),
    "ViT-H/14": (14, 1280, 32),
}


def load_encoder(name: str, experts: dict, image_resolution: int) -> VisionTransformer:
    """Same signature as ``vit.py:175``.  There is no network here: the architecture is built from the model name with
    its random init; pretrained CLIP weights come in through ``load_state_dict`` (reference key layout)."""
    if name not in VIT_CONFIGS:
        raise RuntimeError(f"Model {name} not found")
    patch, width, layers = VIT_CONFIGS[name]
    return VisionTransformer(image_resolution, patch, width, layers, width // 64, experts)
