"""Development: tok_proj time by output width (self: N = 352 in two accumulators; cross (128, 96): N = 128 and N = 224)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from garbage_classification_rca_b200 import _native as N, functional as F
B, L, K = 256, 197, int(sys.argv[1]) if len(sys.argv) > 1 else 1024
g = torch.Generator().manual_seed(0)
def lin(o, i):
    k = 1.0 / i ** 0.5
    return [((torch.rand(o, i, generator=g) * 2 - 1) * k).cuda(), ((torch.rand(o, generator=g) * 2 - 1) * k).cuda()]
params = lin(128, K) + lin(128, K) + lin(96, K) + [torch.ones(96).cuda(), torch.zeros(96).cuda()]
x = [torch.randn(B, L, K, generator=g).bfloat16().cuda() for _ in range(2)]
sa = F.TokenAttention(params, B, L)
ca = F.TokenAttention(params, B, L, reverse=True)
for i in range(3):
    sa(x[i % 2]); ca(x[i % 2], x[(i + 1) % 2])
torch.cuda.synchronize()
N.timing_begin(256)
for i in range(5):
    sa(x[i % 2]); ca(x[i % 2], x[(i + 1) % 2])
recs = [r for r in N.timing_end(256) if r[0] == "tok_proj"]
for j, lab in enumerate(("N=352 (2 x 176)", "N=128", "N=224")):
    ts = [t for i, (n, t) in enumerate(recs) if i % 3 == j]
    nn = (352, 128, 224)[j]
    print(f"{lab}: {1e3 * sum(ts) / len(ts):.1f} us  -> {2.0 * B * L * K * nn / (sum(ts) / len(ts) * 1e-3) / 1e12:.0f} TFLOP/s")
