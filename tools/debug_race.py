import random, sys
import torch
sys.path.insert(0, ".")
from prismer_b200 import engine, synthetic
from prismer_b200.prismer_caption import PrismerCaption
from tests.test_surface_gpu import _model, _experts, TINY_DEC

m0 = _model(PrismerCaption)
ex = _experts(2)
m0.train()
opt = torch.optim.AdamW([p for p in m0.parameters() if p.requires_grad], lr=1e-3)
for _ in range(2):
    loss = m0(ex, ["A picture of a dog", "A picture of two"], prefix="A picture of"); opt.zero_grad(); loss.backward(); opt.step()
torch.cuda.synchronize()
print("m0 trained", flush=True)

m = _model(PrismerCaption)
m.expert_encoder.train(); m.text_decoder.eval()
ids, mask = synthetic.synth_tokens(2, 8, TINY_DEC["vocab_size"], 5, ragged=True)
ids, mask = ids.cuda(), mask.cuda()
labels = ids.masked_fill(ids == 1, -100); labels[:, :3] = -100
st = engine.prepare(m)
grads = []
for i in range(3):
    random.seed(1)
    loss = engine.train_loss(m, ex, ids, mask, labels); loss.backward(); torch.cuda.synchronize()
    grads.append(st.grad_t.clone())
    print("eager", i, flush=True)
engine.SIDE_STREAM = False
random.seed(1)
loss = engine.train_loss(m, ex, ids, mask, labels); loss.backward(); torch.cuda.synchronize()
ref = st.grad_t.clone()
print("eager no-side done", flush=True)
engine.SIDE_STREAM = True
random.seed(1)
g = engine.GraphedTrainStep(m, ex, ids, mask, labels, warmup=1)
print("captured", flush=True)
for i in range(2):
    random.seed(1); g(); torch.cuda.synchronize(); grads.append(st.grad_t.clone())
names = {id(p): n for n, p in m.named_parameters()}
for gi, gr in enumerate(grads):
    tot = float((gr - ref).norm() / ref.norm())
    bad = []
    for p in st.train_params:
        tr, o = st._offset[id(p)]
        a, b = gr[o:o + p.numel()], ref[o:o + p.numel()]
        e = float((a - b).norm() / b.norm().clamp_min(1e-20))
        if e > 1e-3:
            bad.append((names[id(p)], round(e, 4)))
    print(("eager" if gi < 3 else "graph"), gi, f"total {tot:.2e}", bad[:8])

dbg = m.text_decoder._dbg
def cmp(a, b):
    return float((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-20))
for i in range(len(dbg)):
    d, r = dbg[i], dbg[3]
    print(i, "dkv", f"{cmp(d['dkv'], r['dkv']):.2e}", "enc", f"{cmp(d['enc'], r['enc']):.2e}", "kv", f"{cmp(d['kv'], r['kv']):.2e}",
          "q", [f"{cmp(a, b):.1e}" for a, b in zip(d['q'], r['q'])], "o", [f"{cmp(a, b):.1e}" for a, b in zip(d['o'], r['o'])],
          "lse", [f"{cmp(a, b):.1e}" for a, b in zip(d['lse'], r['lse'])])
    if i != 3:
        diff = (d['dkv'].float() - r['dkv'].float()).abs()
        rows = (diff.amax(1) > 0).nonzero().flatten().tolist()
        cols = (diff.amax(0) > 0).nonzero().flatten().tolist()
        print("   rows differing:", rows[:40], "n=", len(rows), " cols differing: n=", len(cols), cols[:8], cols[-4:])
