"""Profiling driver: a few launches of the token-level self-attention block (ViT-L/16 shape) for ncu."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from garbage_classification_rca_b200 import functional as F
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
g = torch.Generator().manual_seed(0)
def lin(o, i):
    k = 1.0 / i ** 0.5
    return [((torch.rand(o, i, generator=g) * 2 - 1) * k).cuda(), ((torch.rand(o, generator=g) * 2 - 1) * k).cuda()]
params = lin(128, 1024) + lin(128, 1024) + lin(96, 1024) + [torch.ones(96).cuda(), torch.zeros(96).cuda()]
x = [torch.randn(B, 197, 1024, generator=g).bfloat16().cuda() for _ in range(2)]
blk = F.TokenAttention(params, B, 197)
for i in range(6):
    blk(x[i % 2])
torch.cuda.synchronize()
print("done")
