"""Diagnostic: per-tensor gradient errors of a compute mode vs the fp32 oracle (run on the GPU box)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import garbage_classification_rca_b200 as g
from garbage_classification_rca_b200 import _native as N
from garbage_classification_rca_b200.training import CrossEntropyLoss
from oracle import mmrca_oracle as orc
from tests._util import make_inputs

compute = int(sys.argv[1]) if len(sys.argv) > 1 else N.COMPUTE_BF16_FUSED
for qk in (1.0,):
    for flags in ((True, False, True),):
        rev, fo, co = flags
        B = 200
        p = orc.init_head_params(features_only=fo, cross_attention_only=co, seed=31, qk_gain=qk)
        img, txt, labels = make_inputs(B, 31)
        ref = orc.np_head_forward_backward(p, img.numpy(), txt.numpy(), rev, fo, co, labels=labels.numpy())
        names = g.head_param_names(fo, co)
        params = [p[n].cuda().requires_grad_(True) for n in names]
        logits = g.mmrca_head(img.cuda(), txt.cuda(), params, reverse=rev, features_only=fo, cross_attention_only=co,
                              compute=compute)
        loss = CrossEntropyLoss()(logits, labels.cuda())
        loss.backward()
        torch.cuda.synchronize()
        print(f"== qk_gain {qk} flags {flags} compute {compute}: logits max abs err "
              f"{np.abs(logits.detach().cpu().numpy()-ref['logits']).max():.3e} loss err {abs(loss.item()-ref['loss']):.2e}")
        scale = max(np.abs(v).max() for v in ref['grads'].values())
        worst_g = 0.0
        for n, t in zip(names, params):
            r = ref['grads'][n]; o = t.grad.cpu().numpy()
            worst_g = max(worst_g, np.abs(o - r).max() / scale)
            print(f"   {n:42s} max|ref|/global {np.abs(r).max()/scale:.1e}  maxerr/max {np.abs(o-r).max()/max(np.abs(r).max(),1e-30):.2e}"
                  f"  l2err/l2 {np.linalg.norm(o-r)/max(np.linalg.norm(r),1e-30):.2e}  maxerr/global {np.abs(o-r).max()/scale:.1e}")
        print(f"   worst maxerr/global = {worst_g:.2e}")
