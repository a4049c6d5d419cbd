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
    ("ce_loss_fwd [960,50265] fp32", B * T * V * 4, lambda: ops.ce_loss_fwd(logits, labels, V)),
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
