"""Kernel-time breakdown of one eager training step (torch.profiler / CUPTI; quick look -- the committed evidence is ncu's)."""
import collections
import re
import sys

import torch

sys.path.insert(0, ".")
import bench  # noqa: E402
from prismer_b200 import engine, synthetic  # noqa: E402
from prismer_b200.optim import FusedAdamW  # noqa: E402
from prismer_b200.prismer_caption import PrismerCaption  # noqa: E402

dev = torch.device("cuda")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
cfg = {"experts": bench.EXPERTS, "prismer_model": "prismer_base", "image_resolution": 224, "freeze": "freeze_vision"}
model = PrismerCaption(cfg).to(dev)
engine.prepare(model, dev)
opt = FusedAdamW(model, lr=5e-5, weight_decay=0.05)
ex, ids, mask = bench.build_inputs(B, 1)
ex = synthetic.experts_to(ex, dev); ids, mask = ids.to(dev), mask.to(dev)
model.train()


def step():
    loss = model(ex, input_ids=ids, attention_mask=mask, prompt_length=4)
    loss.backward()
    opt.step()


for _ in range(2):
    step()
torch.cuda.synchronize()
with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
    step()
    torch.cuda.synchronize()
tot = collections.defaultdict(lambda: [0, 0.0])
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA:
        name = e.name
        m = re.search(r"gemm_bf16_kernel<(\d+), *\(bool\)(\d), *\(bool\)(\d)>", name) or re.search(r"gemm_bf16_kernel<(\d+), *(\w+), *(\w+)>", name)
        key = f"gemm<{m.group(1)},{m.group(2)},{m.group(3)}>" if m else (re.search(r"(\w+_kernel)", name).group(1) if re.search(r"(\w+_kernel)", name) else name[:40])
        tot[key][0] += 1
        tot[key][1] += e.device_time_total if hasattr(e, "device_time_total") else e.cuda_time_total
S = sum(v[1] for v in tot.values())
print(f"total kernel time {S/1e3:.2f} ms")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1])[:32]:
    print(f"{v[1]/1e3:8.3f} ms {100*v[1]/S:5.1f}%  n={v[0]:4d}  {k}")

for kname in ("attn_fwd_kernel", "attn_bwd_dq_kernel", "attn_bwd_dkv_kernel", "ln_bwd_kernel"):
    ds = [round((e.device_time_total if hasattr(e, "device_time_total") else e.cuda_time_total), 1) for e in prof.events()
          if e.device_type == torch.autograd.DeviceType.CUDA and kname in e.name]
    print(kname, ds)
