"""Shared blocks -- parameter containers mirroring ``model/modules/utils.py`` of the reference.

The arithmetic of these blocks is executed by fused sm_100a kernels (``prismer_b200.engine``): activations live in
GEMM epilogues, LayerNorm in ``prismer_layernorm_fwd/bwd``.  Calling them stand-alone routes to the same kernels.
"""
import torch
import torch.nn as nn


class LayerNorm(nn.LayerNorm):
    """fp32-statistics LayerNorm (reference ``utils.py:14-19``), executed by ``prismer_layernorm_fwd``."""

    def forward(self, x: torch.Tensor):
        from .. import engine
        return engine.standalone_layernorm(self, x)


class QuickGELU(nn.Module):
    """x * sigmoid(1.702 x) (``utils.py:23-25``): fused into the c_fc GEMM epilogue (PRISMER_ACT_QUICKGELU)."""
    act = "quickgelu"

    def forward(self, x):
        raise RuntimeError("QuickGELU is fused into the producing GEMM epilogue; it has no stand-alone path")


class SquaredReLU(nn.Module):
    """relu(x)^2 (``utils.py:28-30``): fused into the GEMM epilogue (PRISMER_ACT_SQRELU)."""
    act = "sqrelu"

    def forward(self, x):
        raise RuntimeError("SquaredReLU is fused into the producing GEMM epilogue; it has no stand-alone path")


def interpolation_matrix(orig_len: int, target_len: int) -> torch.Tensor:
    """The bicubic (align_corners=False) resize of ``utils.py:34-44`` is linear in the embedding, so it is a fixed
    [target_len, orig_len] matrix; built once on the host at construction time (not on the hot path) and applied /
    back-propagated through with the GEMM kernel."""
    import torch.nn.functional as F
    o, n = int(orig_len ** 0.5), int(target_len ** 0.5)
    eye = torch.eye(o * o, dtype=torch.float64).reshape(1, o, o, o * o).permute(0, 3, 1, 2)
    m = F.interpolate(eye, size=(n, n), mode="bicubic", align_corners=False)
    return m.permute(0, 2, 3, 1).reshape(n * n, o * o).float()


def interpolate_pos_embed(orig_pos_embed: torch.Tensor, target_len: int) -> torch.Tensor:
    """Checkpoint-time helper with the reference's signature (``train_caption.py:98-99``); runs on whatever device
    the checkpoint tensor lives on (load time, not the hot path)."""
    if int(orig_pos_embed.shape[0] ** 0.5) == int(target_len ** 0.5):
        return orig_pos_embed
    m = interpolation_matrix(orig_pos_embed.shape[0], target_len).to(orig_pos_embed)
    return m @ orig_pos_embed


class _Projection(nn.Module):
    """down_proj -> sq_relu -> up_proj holder so the keys read ``adaptor.down_proj.* / adaptor.up_proj.*``."""

    def __init__(self, dim):
        super().__init__()
        self.down_proj = nn.Linear(dim, dim)
        self.sq_relu = SquaredReLU()
        self.up_proj = nn.Linear(dim, dim)


class Adaptor(nn.Module):
    """Bottleneck-free adaptor (``utils.py:48-65``): norm-early in the ViT, norm-late in the decoder."""

    def __init__(self, embed_dim: int, norm_late: bool = False):
        super().__init__()
        self.norm_late = norm_late
        self.adaptor = _Projection(embed_dim)
        self.adaptor_ln = LayerNorm(embed_dim)
