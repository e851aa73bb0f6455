"""Profiling driver: a few bf16 training steps of the head at batch B (for ncu -k <kernel>)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import garbage_classification_rca_b200 as g
from garbage_classification_rca_b200 import _native as N, functional as F
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
params = F.init_head_parameters("cuda", seed=0)
step = g.HeadTrainStep(params, B, 1280, 768, reverse=True, compute=N.COMPUTE_BF16, drop_p=0.6)
img = torch.randn(B, 1280, device="cuda"); txt = torch.randn(B, 768, device="cuda")
lab = torch.randint(0, 4, (B,), device="cuda")
for i in range(3):
    step.zero_grad(); step(img, txt, lab, drop_seed=i)
torch.cuda.synchronize()
print("loss", float(step.loss))
