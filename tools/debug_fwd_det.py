import random, sys
import torch
sys.path.insert(0, ".")
from prismer_b200 import engine, synthetic
from prismer_b200.prismer_caption import PrismerCaption
from tests.test_surface_gpu import _model, _experts

m = _model(PrismerCaption)
m.expert_encoder.train()
ex = engine._canon_experts(_experts(2))
engine.prepare(m)
vit = m.expert_encoder


def cmp(a, b):
    return float((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-20))


def snap():
    random.seed(1)
    out, S, B, sv = engine.encoder_forward(vit, ex, save=True)
    torch.cuda.synchronize()
    d = {"out": out.clone(), "xf": sv.xf.clone(), "x0": sv.pre[0].clone()}
    for mo in sv.mods:
        for i, L in enumerate(mo.stem.layers):
            d[f"{mo.e}.y{i}"] = L.y.clone(); d[f"{mo.e}.scale{i}"] = L.scale.clone(); d[f"{mo.e}.A{i}"] = L.A.clone()
        d[f"{mo.e}.A5"] = mo.stem.A5.clone()
        if mo.table is not None:
            d[f"{mo.e}.table"] = mo.table.clone().float()
    for l, r in enumerate(sv.res):
        d[f"res{l}.kvin"] = r.kvin.clone(); d[f"res{l}.q"] = r.q.clone(); d[f"res{l}.kv"] = r.kv.clone(); d[f"res{l}.o"] = r.o.clone(); d[f"res{l}.lat1"] = r.lat1.clone()
    for l, b in enumerate(sv.blocks):
        d[f"blk{l}.qkv"] = b.qkv.clone(); d[f"blk{l}.o"] = b.o.clone(); d[f"blk{l}.x1"] = b.x1.clone(); d[f"blk{l}.x2"] = b.x2.clone()
    return d


ref = snap()
for it in range(5):
    cur = snap()
    bad = [(k, f"{cmp(cur[k], ref[k]):.1e}") for k in ref if cmp(cur[k], ref[k]) > 0]
    print(it, "first differing:", bad[:6], flush=True)
