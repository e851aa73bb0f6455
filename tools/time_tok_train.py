"""Development: per-launch times of one token-level fwd+bwd step (B samples), in launch order."""
import sys
sys.path.insert(0, ".")
import torch
from garbage_classification_rca_b200 import _native as N, functional as F

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
dev = torch.device("cuda", 0)
g = torch.Generator().manual_seed(7)


def block(d_q, d_kv, d_kq, d_v):
    def lin(o, i):
        k = 1.0 / i ** 0.5
        return [((torch.rand(o, i, generator=g) * 2 - 1) * k).to(dev), ((torch.rand(o, generator=g) * 2 - 1) * k).to(dev)]
    return lin(d_kq, d_q) + lin(d_kq, d_kv) + lin(d_v, d_kv) + [torch.ones(d_v, device=dev), torch.zeros(d_v, device=dev)]


for (L, K, dkq, dv, cross) in [(197, 1024, 128, 96, False), (256, 768, 128, 96, False), (197, 96, 64, 48, True), (256, 96, 64, 48, True)]:
    params = block(K, K, dkq, dv)
    blk = F.TokenAttention(params, B, L, reverse=cross, training=True)
    x = torch.randn(B, L, K, generator=g).bfloat16().to(dev)
    x2 = torch.roll(x, 1, 0) if cross else None
    d_out = (torch.randn(B, L, dv, generator=g) / (B * L)).to(dev)
    grads = [torch.zeros_like(t) for t in params]
    for _ in range(3):
        blk(x, x2)
        blk.backward(d_out, grads, cross, cross)
    torch.cuda.synchronize()
    N.timing_begin(256)
    for _ in range(3):
        blk.refresh_weights()
        blk(x, x2)
        blk.backward(d_out, grads, cross, cross)
    recs = N.timing_end(256)
    n = len(recs) // 3
    print(f"L={L} K={K} cross={cross}: " + "  ".join(f"{name}={1e3 * t:.1f}" for name, t in recs[2 * n:]) + f"   total={1e3 * sum(t for _, t in recs[2 * n:]):.1f} us")
