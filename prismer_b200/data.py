"""Host side of the compact expert-input format.

The reference's data workers run ``post_label_process`` (dataset/utils.py:117-160) per sample: depth / normal / edge are
min/max-remapped to [-1, 1]; seg_coco / seg_ade / obj_detection / ocr_detection label maps are in-painted into a
``[64, H, W]`` fp32 stack of CLIP-PCA feature rows (background id 255).  ``compact_label_process`` is the drop-in that keeps
the uint8 map and builds the <= 256-row table instead -- a ``CompactMap`` -- and the model's stems expand it on the GPU
(csrc/compact_inputs.cu).  Per image and modality that is 50 KB + a table instead of 12.8 MB through the worker pipes,
pinned memory and PCIe.

``CompactMap.expand()`` materialises exactly the tensor the reference builds (on whatever device the map lives on, through
the CUDA kernel when on the GPU); the values are bit-identical because the table rows ARE the reference's feature rows and
the min/max remap is evaluated with the reference's own fp32 expression on the 256 possible grey levels."""
from __future__ import annotations

from typing import Dict, Optional

import torch

FEATURE_DIM = 64
BACKGROUND_ID = 255


class CompactMap:
    """uint8 map ``u8`` [Cin, H, W] (or batched [B, Cin, H, W]) + fp32 ``table`` [256, C] (or [B, 256, C])."""

    __slots__ = ("u8", "table")

    def __init__(self, u8: torch.Tensor, table: torch.Tensor):
        assert u8.dtype == torch.uint8 and table.dtype == torch.float32 and table.shape[-2] == 256
        self.u8, self.table = u8, table

    @property
    def device(self):
        return self.u8.device

    def to(self, device, non_blocking: bool = False) -> "CompactMap":
        return CompactMap(self.u8.to(device, non_blocking=non_blocking), self.table.to(device, non_blocking=non_blocking))

    def pin_memory(self) -> "CompactMap":
        return CompactMap(self.u8.pin_memory(), self.table.pin_memory())

    @property
    def shape(self):
        """Shape of the tensor this map stands for."""
        *lead, cin, h, w = self.u8.shape
        return torch.Size([*lead, cin * self.table.shape[-1], h, w])

    def size(self, dim=None):
        return self.shape if dim is None else self.shape[dim]

    def expand(self) -> torch.Tensor:
        """The reference-format fp32 tensor, built by the CUDA kernel (the map must live on the GPU)."""
        from . import ops
        ops._req_cuda(self.u8, self.table)       # no CPU fallback: raises PrismerError (expand_on_host() exists for data-loader checks)
        batched = self.u8.dim() == 4
        u8 = (self.u8 if batched else self.u8[None]).contiguous()
        out = ops.expand_labels(u8, self.table.contiguous())
        return out if batched else out[0]

    def expand_on_host(self) -> torch.Tensor:
        """Index arithmetic on CPU tensors -- what the reference's workers do (dataset/utils.py:120-158).  For data-pipeline
        checks and tests only; the model never calls it."""
        assert not self.u8.is_cuda
        batched = self.u8.dim() == 4
        u8 = self.u8 if batched else self.u8[None]
        t = self.table if self.table.dim() == 3 else self.table[None].expand(u8.shape[0], -1, -1)
        idx = u8.long()                                                          # [B, Cin, H, W]
        rows = torch.stack([t[b][idx[b]] for b in range(u8.shape[0])])           # [B, Cin, H, W, C]
        out = rows.permute(0, 1, 4, 2, 3).reshape(u8.shape[0], -1, *u8.shape[2:]).contiguous()
        return out if batched else out[0]

    @staticmethod
    def collate(maps) -> "CompactMap":
        """Batch per-sample maps (what a DataLoader collate_fn does for tensors)."""
        u8 = torch.stack([m.u8 for m in maps])
        same = all(m.table is maps[0].table for m in maps)
        return CompactMap(u8, maps[0].table if same else torch.stack([m.table for m in maps]))


def minmax_table(u8: torch.Tensor, eps: float = 1e-6) -> torch.Tensor:
    """[256, 1] LUT of ``2 * (x - x.min()) / (x.max() - x.min() + eps) - 1`` with x = u8 / 255 (dataset/utils.py:120-121 after
    ``to_tensor``, :59): the reference's own fp32 expression evaluated on the 256 possible grey levels."""
    x = torch.arange(256, dtype=torch.float32).div(255)
    lo, hi = x[int(u8.min())], x[int(u8.max())]
    return (2 * (x - lo) / (hi - lo + eps) - 1).view(256, 1)


def feature_table(ids, rows: Dict[int, torch.Tensor], background: torch.Tensor) -> torch.Tensor:
    """[256, 64] table: row l = rows[l] for the ids present, row 255 = background, absent ids zero (never read)."""
    t = torch.zeros(256, background.numel(), dtype=torch.float32)
    for l in ids:
        t[l] = background if l == BACKGROUND_ID else rows[l]
    return t


def compact_label_process(inputs: Dict, labels_info: Optional[Dict], features: Dict[str, torch.Tensor]) -> Dict:
    """Drop-in for ``post_label_process`` (dataset/utils.py:117-160), same arguments plus the feature tables
    ``{'coco', 'ade', 'detection': [N, 64], 'background': [64]}`` (the reference loads them as module globals, :17-20).

    ``inputs``: what ``Transform`` hands over, but as uint8 (``pil_to_tensor``) instead of float / int64 tensors: depth / edge
    [1, H, W], normal [3, H, W], label maps [1, H, W].  Returns the same dict with every expert replaced by a ``CompactMap``
    (``obj_detection`` -> ``{'label': CompactMap, 'instance': int64 map}`` as the reference)."""
    bg = features["background"].float()
    out = dict(inputs)
    for exp, v in inputs.items():
        if exp == "rgb":
            continue
        assert v.dtype == torch.uint8, exp
        if exp in ("depth", "normal", "edge"):
            out[exp] = CompactMap(v, minmax_table(v))
            continue
        ids = torch.unique(v).tolist()
        if exp == "seg_coco":
            rows = {l: features["coco"][l] for l in ids if l != BACKGROUND_ID}
        elif exp == "seg_ade":
            rows = {l: features["ade"][l] for l in ids if l != BACKGROUND_ID}
        elif exp == "obj_detection":
            label_map = labels_info[exp]
            rows = {l: features["detection"][label_map[str(l)]] for l in ids if l != BACKGROUND_ID}
        elif exp == "ocr_detection":
            label_map = labels_info[exp]
            rows = {l: label_map[l]["features"] for l in ids if l != BACKGROUND_ID}
        else:
            raise KeyError(exp)
        cm = CompactMap(v, feature_table(ids, rows, bg))
        out[exp] = {"label": cm, "instance": v.long()} if exp == "obj_detection" else cm
    return out


def collate_experts(samples):
    """collate_fn for dicts that may hold ``CompactMap`` values."""
    out = {}
    for k in samples[0]:
        vs = [s[k] for s in samples]
        if isinstance(vs[0], CompactMap):
            out[k] = CompactMap.collate(vs)
        elif isinstance(vs[0], dict):
            out[k] = collate_experts(vs)
        else:
            out[k] = torch.stack(vs)
    return out
