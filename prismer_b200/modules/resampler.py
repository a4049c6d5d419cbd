"""Experts Resampler -- parameter containers mirroring ``model/modules/resampler.py`` (keys / shapes identical)."""
import torch
import torch.nn as nn

from .utils import LayerNorm, SquaredReLU


class _MLP(nn.Module):
    def __init__(self, d_model, act):
        super().__init__()
        self.c_fc = nn.Linear(d_model, d_model * 4)
        setattr(self, "sq_relu" if isinstance(act, SquaredReLU) else "gelu", act)
        self.c_proj = nn.Linear(d_model * 4, d_model)
        self.act = act.act


class PerceiverAttentionBlock(nn.Module):
    """latents += MHA(LN1(latents), cat(LN1(latents), LN2(x))); latents += MLP(LNff(latents))  (resampler.py:33-36)."""

    def __init__(self, d_model: int, n_heads: int):
        super().__init__()
        self.n_heads = n_heads
        self.attn = nn.MultiheadAttention(d_model, n_heads)  # container for in_proj_weight/bias + out_proj
        self.mlp = _MLP(d_model, SquaredReLU())
        self.ln_1, self.ln_2, self.ln_ff = LayerNorm(d_model), LayerNorm(d_model), LayerNorm(d_model)


class PerceiverResampler(nn.Module):
    def __init__(self, width: int, layers: int, heads: int, num_latents: int):
        super().__init__()
        self.latents = nn.Parameter(width ** -0.5 * torch.randn(num_latents, width))
        self.perceiver_blocks = nn.Sequential(*[PerceiverAttentionBlock(width, heads) for _ in range(layers)])
        self.heads = heads
