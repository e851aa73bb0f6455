"""Diagnostic: per-tensor errors of the bf16 path vs the fp32 oracle (run on the GPU box)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import garbage_classification_rca_b200 as g
from garbage_classification_rca_b200 import _native as N
from garbage_classification_rca_b200.training import CrossEntropyLoss
from oracle import mmrca_oracle as orc
from tests._util import make_inputs

for qk in (1.0, 40.0):
    for compute in (N.COMPUTE_FP32, N.COMPUTE_BF16):
        B = 200
        p = orc.init_head_params(seed=31, qk_gain=qk)
        img, txt, labels = make_inputs(B, 31)
        ref = orc.np_head_forward_backward(p, img.numpy(), txt.numpy(), True, False, False, labels=labels.numpy())
        names = g.head_param_names()
        params = [p[n].cuda().requires_grad_(True) for n in names]
        logits = g.mmrca_head(img.cuda(), txt.cuda(), params, reverse=True, compute=compute)
        loss = CrossEntropyLoss()(logits, labels.cuda())
        loss.backward()
        torch.cuda.synchronize()
        print(f"== qk_gain {qk} compute {compute}: logits max abs err {np.abs(logits.detach().cpu().numpy()-ref['logits']).max():.3e} loss err {abs(loss.item()-ref['loss']):.2e}")
        scale = max(np.abs(v).max() for v in ref['grads'].values())
        for n, t in zip(names, params):
            r = ref['grads'][n]; o = t.grad.cpu().numpy()
            print(f"   {n:42s} max|ref| {np.abs(r).max():.2e} ({np.abs(r).max()/scale:.1e} of global)  relerr {np.abs(o-r).max()/max(np.abs(r).max(),1e-30):.2e}")
