"""Profiling driver: a few launches of one attention block (forward, optionally backward) at batch B."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import garbage_classification_rca_b200 as g
from garbage_classification_rca_b200 import _native as N

kind = sys.argv[1] if len(sys.argv) > 1 else "sa80"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
compute = N.COMPUTE_BF16 if (len(sys.argv) <= 3 or sys.argv[3] == "bf16") else N.COMPUTE_FP32
if kind.startswith("sa"):
    d_in, dkq, dv, self_ = int(kind[2:]), 128, 96, True
else:
    d_in, dkq, dv, self_ = 96, 64, 48, False
torch.manual_seed(0)
p = [torch.randn(dkq, d_in) * 0.1, torch.zeros(dkq), torch.randn(dkq, d_in) * 0.1, torch.zeros(dkq),
     torch.randn(dv, d_in) * 0.1, torch.zeros(dv), torch.ones(dv), torch.zeros(dv)]
p = [t.cuda() for t in p]
xq = torch.randn(B, 16, d_in, device="cuda")
xkv = None if self_ else torch.randn(B, 16, d_in, device="cuda")
for _ in range(3):
    out = g.attention_block(xq, xkv, p, reverse=(kind == "rca"), compute=compute)
torch.cuda.synchronize()
print("ok", float(out.sum()))
