import sys
import torch
sys.path.insert(0, ".")
from prismer_b200 import ops
B, H, Lq, Lk, d = (int(x) for x in sys.argv[1:6])
qkv = torch.randn(Lq, B, 3 * H * d, device="cuda").to(torch.bfloat16)
q3 = qkv.transpose(0, 1)
q, k, v = q3[..., :H * d], q3[..., H * d:2 * H * d], q3[..., 2 * H * d:]
if Lq != Lk:
    kv = torch.randn(Lk, B, 2 * H * d, device="cuda").to(torch.bfloat16).transpose(0, 1)
    k, v = kv[..., :H * d], kv[..., H * d:]
    q = torch.randn(Lq, B, H * d, device="cuda").to(torch.bfloat16).transpose(0, 1)
for _ in range(3):
    o, lse = ops.attention_fwd(q, k, v, H)
    do = torch.randn_like(o)
    ops.attention_bwd(do, q, k, v, o, lse, H)
torch.cuda.synchronize()
e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
e0.record()
for _ in range(10):
    o, lse = ops.attention_fwd(q, k, v, H)
e1.record()
for _ in range(10):
    ops.attention_bwd(do, q, k, v, o, lse, H)
e2.record()
torch.cuda.synchronize()
fl = 4.0 * B * H * Lq * Lk * d
print(f"attn B{B} H{H} Lq{Lq} Lk{Lk} d{d}: fwd {e0.elapsed_time(e1)*100:.1f} us ({fl/e0.elapsed_time(e1)/1e8:.1f} TF/s)  bwd {e1.elapsed_time(e2)*100:.1f} us ({2.5*fl/e1.elapsed_time(e2)/1e8:.1f} TF/s)")
