import sys
import torch
sys.path.insert(0, ".")
from prismer_b200 import ops
M, N, K = (int(x) for x in sys.argv[1:4])
bn = int(sys.argv[4]) if len(sys.argv) > 4 else 0
a = torch.randn(M, K, device="cuda").to(torch.bfloat16)
b = torch.randn(N, K, device="cuda").to(torch.bfloat16)
for _ in range(4):
    c = ops.gemm(a, b, force_bn=bn)
torch.cuda.synchronize()
