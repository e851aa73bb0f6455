"""Profiling driver: forward + backward of the token-level self block (ViT-L/16 shape) and of a cross block, for ncu."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from garbage_classification_rca_b200 import functional as F
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
g = torch.Generator().manual_seed(0)
def lin(o, i):
    k = 1.0 / i ** 0.5
    return [((torch.rand(o, i, generator=g) * 2 - 1) * k).cuda(), ((torch.rand(o, generator=g) * 2 - 1) * k).cuda()]
for (L, K, dkq, dv, cross) in [(197, 1024, 128, 96, False), (197, 96, 64, 48, True)]:
    params = lin(dkq, K) + lin(dkq, K) + lin(dv, K) + [torch.ones(dv).cuda(), torch.zeros(dv).cuda()]
    x = torch.randn(B, L, K, generator=g).bfloat16().cuda()
    x2 = torch.roll(x, 1, 0) if cross else None
    d_out = (torch.randn(B, L, dv, generator=g) / (B * L)).cuda()
    grads = [torch.zeros_like(t) for t in params]
    blk = F.TokenAttention(params, B, L, reverse=cross, training=True)
    for i in range(3):
        blk.refresh_weights()
        blk(x, x2)
        blk.backward(d_out, grads, cross, cross)
    torch.cuda.synchronize()
print("done")
