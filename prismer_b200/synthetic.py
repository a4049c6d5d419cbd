"""Deterministic synthetic weights and inputs (no network, no datasets, no checkpoints).

Everything is drawn from ``numpy.random.RandomState`` (the frozen legacy MT19937 stream, which is
bit-stable across numpy versions and machines), keyed by a CRC of the tensor name, so the build
container (where the golden fixtures are generated from the real reference) and the GPU box produce
bit-identical tensors from nothing but ``(seed, name, shape)``.

Input contract follows the reference's data pipeline (SURVEY.md section 8a row a0;
``dataset/utils.py:30-71,117-160``): rgb fp32 CLIP-normalised, depth/normal/edge fp32 in [-1,1],
seg/obj/ocr 64-channel piece-wise-constant CLIP-PCA feature maps, obj_detection = {label, instance}.
"""
from __future__ import annotations

import zlib
from collections import OrderedDict
from typing import Dict, Iterable

import numpy as np
import torch

EXPERT_CHANNELS = {"rgb": 3, "depth": 1, "normal": 3, "edge": 1, "seg_coco": 64, "seg_ade": 64,
                   "obj_detection": 64, "ocr_detection": 64}
DEFAULT_EXPERTS = ["depth", "normal", "seg_coco", "edge", "obj_detection", "ocr_detection"]  # configs/caption.yaml:5


def _rs(seed: int, name: str) -> np.random.RandomState:
    return np.random.RandomState((seed * 1000003 + zlib.crc32(name.encode())) % (2 ** 32))


def synth_tensor(name: str, shape, seed: int) -> torch.Tensor:
    """Value distribution chosen per parameter kind so activations stay O(1) through the net."""
    shape = tuple(shape)
    rs = _rs(seed, name)
    n = lambda: rs.standard_normal(shape).astype(np.float32)
    if name.endswith("num_batches_tracked"):
        return torch.zeros(shape, dtype=torch.long)
    if name.endswith("position_ids"):
        return torch.arange(shape[-1]).expand(shape).clone()
    if name.endswith("running_mean"):
        a = 0.1 * n()
    elif name.endswith("running_var"):
        a = (0.5 + rs.uniform(0, 1, shape)).astype(np.float32)
    elif len(shape) == 1 and name.endswith("weight"):  # LayerNorm / BatchNorm gains
        a = 1.0 + 0.1 * n()
    elif len(shape) == 1:  # biases (incl. in_proj_bias, lm_head.bias)
        a = 0.02 * n()
    elif len(shape) == 4:  # conv kernels: He init
        fan_in = shape[1] * shape[2] * shape[3]
        a = n() * np.float32(np.sqrt(2.0 / fan_in))
    elif "word_embeddings" in name or name.endswith("lm_head.decoder.weight"):
        a = 0.06 * n()
    elif "embeddings" in name:  # position / token-type tables
        a = 0.02 * n()
    elif name.endswith("latents") or name.endswith("positional_embedding") or name.endswith("instance_embedding"):
        a = n() * np.float32(shape[-1] ** -0.5)
    else:  # linear weights [out, in]
        a = n() * np.float32(0.7 * shape[-1] ** -0.5)
    return torch.from_numpy(np.ascontiguousarray(a))


def synth_state_dict(template: "OrderedDict[str, torch.Tensor]", seed: int = 0) -> "OrderedDict[str, torch.Tensor]":
    """Fill a state_dict *template* (names/shapes from ``module.state_dict()``) deterministically.
    The tied LM-head keys copy the embedding / bias they are tied to (roberta.py:352-356,417-419)."""
    out = OrderedDict()
    for k, v in template.items():
        if v.dtype.is_floating_point or k.endswith("num_batches_tracked") or k.endswith("position_ids"):
            out[k] = synth_tensor(k, v.shape, seed)
        else:
            out[k] = v.clone()
    for k in list(out):
        if k.endswith("lm_head.decoder.weight"):
            out[k] = out[k.replace("lm_head.decoder.weight", "roberta.embeddings.word_embeddings.weight")]
        if k.endswith("lm_head.decoder.bias"):
            out[k] = out[k.replace("lm_head.decoder.bias", "lm_head.bias")]
    for k in list(out):  # nn.Embedding(padding_idx=1) rows are zero at init (roberta.py:51,64)
        if k.endswith("word_embeddings.weight") or k.endswith("position_embeddings.weight"):
            out[k] = out[k].clone()
            out[k][1].zero_()
    for k in list(out):
        if k.endswith("lm_head.decoder.weight"):
            out[k] = out[k.replace("lm_head.decoder.weight", "roberta.embeddings.word_embeddings.weight")]
    return out


def _blocky(rs, B, size, cells, n_ids) -> np.ndarray:
    """[B, size, size] int map, constant on a cells x cells grid (nearest-upsampled)."""
    low = rs.randint(0, n_ids, size=(B, cells, cells))
    idx = (np.arange(size) * cells) // size
    return low[:, idx][:, :, idx]


def synth_experts(batch: int, image_resolution: int = 224, experts: Iterable[str] = DEFAULT_EXPERTS,
                  label_size: int = 224, seed: int = 0, n_classes: int = 133) -> Dict:
    """The ``experts`` dict the model receives.  Dict order = rgb, then ``experts`` order
    (``dataset/utils.py:69``).  Label maps are always ``label_size`` x ``label_size`` (224 in the reference,
    ``dataset/utils.py:43``)."""
    out = OrderedDict()
    rs = _rs(seed, "rgb")
    out["rgb"] = torch.from_numpy(rs.standard_normal((batch, 3, image_resolution, image_resolution)).astype(np.float32))
    for e in experts:
        rs = _rs(seed, e)
        c = EXPERT_CHANNELS["seg_coco" if "seg" in e else e]
        if c < 64:
            out[e] = torch.from_numpy(rs.uniform(-1, 1, (batch, c, label_size, label_size)).astype(np.float32))
            continue
        # piece-wise constant CLIP-PCA-like features (dataset/*_features.pt: mean -0.06, std 0.75)
        table = (-0.06 + 0.75 * rs.standard_normal((n_classes, 64))).astype(np.float32)
        cls = _blocky(rs, batch, label_size, 7, n_classes)
        lab = torch.from_numpy(np.ascontiguousarray(table[cls].transpose(0, 3, 1, 2)))
        if e == "obj_detection":
            inst = _blocky(rs, batch, label_size, 5, 5).astype(np.int64)
            inst[inst == 4] = 255  # background id (dataset/utils.py:141-160)
            out[e] = {"label": lab, "instance": torch.from_numpy(inst[:, None].copy())}
        else:
            out[e] = lab
    return out


def synth_tokens(batch: int, length: int, vocab: int = 50265, seed: int = 0, ragged: bool = False):
    """Fixed-length ids ``randint(3, V)`` with ``ids[:,0] = <s> = 0``;
    ``ragged`` pads a random tail with ``<pad> = 1`` like ``padding='longest'`` does."""
    rs = _rs(seed, "tokens")
    ids = rs.randint(3, vocab, size=(batch, length)).astype(np.int64)
    ids[:, 0] = 0
    mask = np.ones((batch, length), dtype=np.int64)
    if ragged:
        lens = rs.randint(max(2, length // 2), length + 1, size=batch)
        lens[0] = length
        for b in range(batch):
            ids[b, lens[b] - 1] = 2
            ids[b, lens[b]:] = 1
            mask[b, lens[b]:] = 0
    else:
        ids[:, -1] = 2
    return torch.from_numpy(ids), torch.from_numpy(mask)


def experts_to(experts: Dict, device, non_blocking: bool = False, dtype=None) -> Dict:
    out = OrderedDict()
    for k, v in experts.items():
        if isinstance(v, dict):
            out[k] = {kk: vv.to(device, non_blocking=non_blocking) for kk, vv in v.items()}
        else:
            out[k] = v.to(device, non_blocking=non_blocking)
    return out


# ------------------------------------------------------------------------------------------------ compact expert inputs
def synth_features(seed: int = 0) -> Dict[str, torch.Tensor]:
    """Stand-ins for dataset/{coco,ade,detection,background}_features.pt (CLIP-PCA rows: mean -0.06, std 0.75)."""
    rs = _rs(seed, "features")
    f = lambda n: torch.from_numpy((-0.06 + 0.75 * rs.standard_normal((n, 64))).astype(np.float32))
    return {"coco": f(133), "ade": f(150), "detection": f(722), "background": f(1)[0]}


def synth_compact_experts(batch: int, image_resolution: int = 224, experts: Iterable[str] = DEFAULT_EXPERTS,
                          label_size: int = 224, seed: int = 0) -> Dict:
    """The same pipeline as the reference's workers, in the compact format: per-sample uint8 maps ->
    ``data.compact_label_process`` (drop-in for dataset/utils.py:117-160) -> ``data.collate_experts``."""
    from . import data
    feats = synth_features(seed)
    samples = []
    for b in range(batch):
        rs = _rs(seed, f"compact.{b}")
        s = OrderedDict(rgb=torch.from_numpy(rs.standard_normal((3, image_resolution, image_resolution)).astype(np.float32)))
        info = {}
        for e in experts:
            if e in ("depth", "edge", "normal"):
                s[e] = torch.from_numpy(rs.randint(0, 256, (EXPERT_CHANNELS[e], label_size, label_size)).astype(np.uint8))
                continue
            n = {"seg_coco": 133, "seg_ade": 150, "obj_detection": 5, "ocr_detection": 3}[e]
            m = _blocky(rs, 1, label_size, 7, n)
            m = np.where(_blocky(rs, 1, label_size, 5, 4) == 0, 255, m).astype(np.uint8)
            s[e] = torch.from_numpy(m)
            if e == "obj_detection":
                info[e] = {str(i): int(rs.randint(0, 722)) for i in range(n)}
            elif e == "ocr_detection":
                info[e] = {i: {"features": torch.from_numpy((0.75 * rs.standard_normal(64)).astype(np.float32))} for i in range(n)}
        samples.append(data.compact_label_process(s, info, feats))
    return data.collate_experts(samples)


def expand_compact_on_host(experts: Dict) -> Dict:
    """The float ``experts`` dict the reference's workers would have produced for a compact one (host tensors)."""
    from . import data
    out = OrderedDict()
    for k, v in experts.items():
        if isinstance(v, data.CompactMap):
            out[k] = v.expand_on_host()
        elif isinstance(v, dict):
            out[k] = {kk: (vv.expand_on_host() if isinstance(vv, data.CompactMap) else vv) for kk, vv in v.items()}
        else:
            out[k] = v
    return out


class SyntheticCaptionDataset(torch.utils.data.Dataset):
    """Stands in for ``dataset/caption_dataset.py:Caption`` (train split): ``__getitem__ -> (experts, caption)`` with per-sample
    tensors in the reference's format, or -- ``compact=True`` -- uint8 maps pushed through ``data.compact_label_process`` the way a
    patched ``Caption.__getitem__`` would.  Use ``collate`` as the loader's ``collate_fn``."""

    WORDS = "a an the dog cat man woman street table plate bus train sitting standing holding red blue green two three".split()

    def __init__(self, n: int, experts: Iterable[str] = DEFAULT_EXPERTS, image_resolution: int = 224, label_size: int = 224,
                 compact: bool = False, prefix: str = "A picture of", seed: int = 0):
        self.n, self.experts, self.res, self.label, self.compact, self.prefix, self.seed = (n, list(experts), image_resolution,
                                                                                           label_size, compact, prefix, seed)

    def __len__(self):
        return self.n

    def __getitem__(self, i):
        if self.compact:
            batch = synth_compact_experts(1, self.res, self.experts, self.label, self.seed * 100003 + i)
        else:
            batch = synth_experts(1, self.res, self.experts, self.label, self.seed * 100003 + i)
        def first(v):
            if isinstance(v, dict):
                return {k: first(x) for k, x in v.items()}
            if hasattr(v, "u8"):
                return type(v)(v.u8[0], v.table[0] if v.table.dim() == 3 else v.table)
            return v[0]
        rs = _rs(self.seed, f"caption.{i}")
        caption = self.prefix + " " + " ".join(self.WORDS[j] for j in rs.randint(0, len(self.WORDS), rs.randint(4, 11)))
        return {k: first(v) for k, v in batch.items()}, caption

    @staticmethod
    def collate(samples):
        from . import data
        return data.collate_experts([s[0] for s in samples]), [s[1] for s in samples]
