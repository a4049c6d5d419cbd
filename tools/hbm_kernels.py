"""One launch of every HBM-bound kernel of the hot path at its Prismer-BASE training shape (B = 32), for
    ncu --set full --clock-control none -o gpurun_out/hbm_rN python tools/hbm_kernels.py
(achieved DRAM GB/s / dram__throughput per kernel: BASELINE.json north_star asks for these on the LayerNorm / softmax-less / embedding /
channel-stack paths).  Also prints CUDA-event timings and the algorithmic bytes of each call (after 2 warm-up calls, L2 flushed)."""
import sys

import torch

sys.path.insert(0, ".")
from prismer_b200 import ops  # noqa: E402

dev = "cuda"
B, S, D, T, Hd, V = 32, 260, 768, 30, 768, 50265
g = torch.Generator(device=dev).manual_seed(0)
rn = lambda *s: torch.randn(*s, device=dev, generator=g)
bf = lambda *s: rn(*s).to(torch.bfloat16)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
rows = S * B
x, dy, dres = bf(rows, D), bf(rows, D), bf(rows, D)
gamma, beta = rn(D), rn(D)
dg, db = torch.zeros(D, device=dev), torch.zeros(D, device=dev)
y, mean, rstd = ops.layernorm_fwd(x, gamma, beta)
xd, dyd = bf(B * T, Hd), bf(B * T, Hd)
_, md, rd = ops.layernorm_fwd(xd, gamma, beta)
seed = torch.tensor([7], dtype=torch.int64, device=dev)
ids = torch.randint(3, V, (B, T), device=dev, generator=g)
word, pos, typ = bf(V, Hd), bf(514, Hd), bf(1, Hd)
lab = rn(B, 64, 224, 224)
colg = torch.zeros(3072, device=dev)
z4 = bf(rows, 3072)
logits = rn(B * T, 50272)
labels = torch.randint(3, V, (B, T), device=dev, generator=g)
act56 = bf(B * 56 * 56, 96)
# conv-stem layer 1 of a depth / normal / edge stem (vit.py:105-119): y1 [B,112,112,96], consumed by the 3x3 stride-2 conv of layer 2
y112 = bf(B * 112 * 112, 96)
dA112 = bf(B * 56 * 56, 9 * 96)
ch = lambda: (rn(96).abs() + 0.5, rn(96), rn(96), rn(96).abs() + 0.5)
sc1, sh1, mu1, rs1 = ch()
gam1 = rn(96)
# KV-cached decode step kernels (csrc/decode.cu) at the BASE caption-inference shape: weight-streaming skinny GEMMs and single-query attention
from prismer_b200 import kv_decode  # noqa: E402
xq = bf(B, Hd)
w_fc, w_head = bf(3072, Hd) * 0.03, word
b_fc = rn(3072)
kv_all = bf(S * B, 2 * Hd)                       # one layer's projected visual K | V, seq-first rows (s*B + b) like engine.cross_kv
lnm = torch.nn.LayerNorm(Hd).to(dev)
CALLS = [
    ("ln_fwd  [8320,768]", 2 * rows * D * 2, lambda: ops.layernorm_fwd(x, gamma, beta)),
    ("ln_bwd  [8320,768] frozen (dx only, + residual grad)", 4 * rows * D * 2, lambda: ops.layernorm_bwd(dy, x, mean, rstd, gamma, dres=dres)),
    ("ln_bwd  [8320,768] trainable (dgamma/dbeta)", 4 * rows * D * 2,
     lambda: ops.layernorm_bwd(dy, x, mean, rstd, gamma, dres=dres, dgamma=dg, dbeta=db)),
    ("ln_bwd  [960,768] decoder (dz with dropout)", 4 * B * T * Hd * 2,
     lambda: ops.layernorm_bwd(dyd, xd, md, rd, gamma, dgamma=dg, dbeta=db, dz=True, drop_p=0.1, seed=seed, rng_stream=3)),
    ("colsum  [8320,3072]", rows * 3072 * 2, lambda: ops.colsum(z4, colg)),
    ("embed_fwd [32,30]x768", 4 * B * T * Hd * 2, lambda: ops.embed_fwd(ids, word, pos, typ, 1)),
    ("resample_bilinear fp32 [32,64,224,224] -> bf16 NHWC 56x56 (reads every other row pair)", B * 64 * 224 * 224 * 4 // 2 + B * 56 * 56 * 64 * 2,
     lambda: ops.resample_bilinear(lab, 56, 56)),
    ("im2col_nhwc 56x56x96 k3 s2", B * 56 * 56 * 96 * 2 + B * 28 * 28 * 9 * 96 * 2, lambda: ops.im2col_nhwc(act56, B, 56, 56, 96, 3, 2)),
    ("im2col_nhwc 112x112x96 k3 s2 + BN affine + ReLU on load", B * 112 * 112 * 96 * 2 + B * 56 * 56 * 9 * 96 * 2,
     lambda: ops.im2col_nhwc(y112, B, 112, 112, 96, 3, 2, sc1, sh1)),
    ("bn_relu_bwd 112x112x96 under a k3 s2 conv: gather (dAcol + y -> dn) and apply (dn + y -> dy), two launches",
     (B * 56 * 56 * 9 * 96 + 5 * B * 112 * 112 * 96) * 2,
     lambda: ops.bn_relu_bwd(dA112, y112, sc1, sh1, mu1, rs1, gam1, None, None, B, 112, 112, 96, 3, 2, 56, 56)),
    ("ce_loss_fwd [960,50265] fp32", B * T * V * 4, lambda: ops.ce_loss_fwd(logits, labels, V)),
    ("skinny_linear [32,768]x[3072,768] gelu (decode MLP fc: weights streamed once)", 3072 * Hd * 2, lambda: kv_decode.skinny_linear(xq, w_fc, b_fc, act="gelu")),
    ("skinny_linear [32,768]x[768,768] + residual, then ln_fwd on the 32 rows", Hd * Hd * 2,
     lambda: kv_decode.skinny_linear(xq, w_fc[:Hd], b_fc[:Hd], residual=xq, ln=lnm)),
    ("skinny_linear [32,768]x[50265,768] fp32 logits (tied LM head)", V * Hd * 2, lambda: kv_decode.skinny_linear(xq, w_head, None, out_dtype=torch.float32)),
    ("decode_attention cross: 32x12 queries over 260 visual keys (K, V read once)", 2 * B * S * Hd * 2,
     lambda: kv_decode.decode_attention(xq, kv_all, kv_all[:, Hd:], 2 * Hd, B * 2 * Hd, S, 12)),
]
for name, nbytes, fn in CALLS:
    for _ in range(2):
        fn()
    ts = []
    for _ in range(5):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    t = sorted(ts)[len(ts) // 2]
    print(f"{name:90s} {t:8.1f} us  {nbytes / 1e6:8.1f} MB algorithmic  {nbytes / t / 1e3:7.0f} GB/s")
