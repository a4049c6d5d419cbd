"""Static launch manifest of one Prismer-BASE fine-tune step (no GPU needed): the engine's host code is dry-run against a recorder
that captures every C-ABI call and, for the GEMMs / attention calls, their argument blocks.  Prints the GEMM shape histogram, the
FLOPs the step actually issues per image and how that compares with the analytic figure bench.py reports MFU against
(263.1 GFLOP/img) -- i.e. that no work is skipped.

    python tools/launch_manifest.py [batch]        (default 8; M scales linearly with the batch)
"""
import collections
import ctypes
import os
import random
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from prismer_b200 import _C, engine, ops, synthetic  # noqa: E402
from prismer_b200.prismer_caption import PrismerCaption  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
gemms, attn, calls = [], [], collections.Counter()


class Rec:
    def __getattr__(self, name):
        def fn(*a):
            calls[name] += 1
            if name == "prismer_gemm_bf16":
                g = ctypes.cast(a[0], ctypes.POINTER(_C.GemmArgs)).contents
                gemms.append((g.M, g.N, g.K, g.transA, g.transB, g.accumulate))
            elif name in ("prismer_attention_fwd", "prismer_attention_bwd"):
                g = ctypes.cast(a[0], ctypes.POINTER(_C.AttnArgs)).contents
                attn.append((name[-3:], g.B, g.H, g.Lq, g.Lk, g.d, g.causal))
            return 0
        return fn


rec = Rec()
_C.lib = lambda: rec
ops._stream = lambda: 0
ops._req_cuda = lambda *t: None
engine._experts_check = lambda e: None
engine.SIDE_STREAM = False

torch.manual_seed(0)
model = PrismerCaption({"experts": synthetic.DEFAULT_EXPERTS, "prismer_model": "prismer_base", "image_resolution": 224, "freeze": "freeze_vision"})
engine.prepare(model, torch.device("cpu"))
model.train()
ex = synthetic.synth_experts(B, 224, synthetic.DEFAULT_EXPERTS, 224, 1)
ids, mask = synthetic.synth_tokens(B, 30, 50265, 1)
labels = ids.masked_fill(ids == 1, -100)
labels[:, :4] = -100
random.seed(0)
engine.train_loss(model, ex, ids, mask, labels).backward()

g_flop = sum(2.0 * M * N * K for M, N, K, *_ in gemms)
a_flop = sum((4.0 if kind == "fwd" else 10.0) * b * h * lq * lk * d for kind, b, h, lq, lk, d, _ in attn)
print(f"batch {B}: {sum(calls.values())} C-ABI calls; {len(gemms)} GEMMs, {len(attn)} attention calls")
print(f"GEMM FLOPs {g_flop / B / 1e9:.1f} GFLOP/img + attention (QK^T, PV and their backward, counted dense) {a_flop / B / 1e9:.1f} GFLOP/img "
      f"= {(g_flop + a_flop) / B / 1e9:.1f} GFLOP/img   (analytic: 263.1)")
hist = collections.Counter((M, N, K, ta, tb, acc) for M, N, K, ta, tb, acc in gemms)
print("\nGEMM shapes (M scales with the batch):   count        M      N      K  tA tB acc   GFLOP total")
for (M, N, K, ta, tb, acc), n in sorted(hist.items(), key=lambda kv: -kv[1] * kv[0][0] * kv[0][1] * kv[0][2])[:40]:
    print(f"                                        {n:6d} {M:8d} {N:6d} {K:6d}   {ta}  {tb}  {acc}   {2.0 * M * N * K * n / 1e9:10.2f}")
print("\nattention calls:", dict(collections.Counter(attn)))
print("\nother entry points:", {k: v for k, v in calls.most_common() if "gemm" not in k and "attention" not in k})
