"""Diagnostic (run on the GPU box): logits / loss / gradient errors of a compute mode against the fp32 oracle,
next to what torch's own bf16 autocast of the same algorithm gets on the CPU (the "ordinary bf16" yardstick).

  python tools/diag_bf16.py [compute=1] [B=200] [drop_p=0.0] [verbose=0]
"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import garbage_classification_rca_b200 as g
from garbage_classification_rca_b200 import _native as N, functional as F
from garbage_classification_rca_b200.training import CrossEntropyLoss
from oracle import mmrca_oracle as orc
from tests._util import make_inputs, grad_summary

compute = int(sys.argv[1]) if len(sys.argv) > 1 else N.COMPUTE_BF16
B = int(sys.argv[2]) if len(sys.argv) > 2 else 200
drop_p = float(sys.argv[3]) if len(sys.argv) > 3 else 0.0
verbose = int(sys.argv[4]) if len(sys.argv) > 4 else 0


def autocast_grads(p, img, txt, labels, rev, fo, co, mask, scale):
    q = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    with torch.autocast("cpu", dtype=torch.bfloat16):
        logits = orc.head_forward(q, img, txt, rev, fo, co, drop_mask=mask, drop_scale=scale)
    orc.cross_entropy(logits.float(), labels).backward()
    return logits.detach().float().numpy(), {k: v.grad.float().numpy() for k, v in q.items() if v.grad is not None}


for qk in (1.0, 40.0):
    for flags in ((True, False, False), (False, False, False), (True, False, True)):
        rev, fo, co = flags
        p = orc.init_head_params(features_only=fo, cross_attention_only=co, seed=31, qk_gain=qk)
        img, txt, labels = make_inputs(B, 31)
        D = F.concat_width(1280, 768, fo, co)
        mask, scale = None, 1.0
        if drop_p > 0:
            mask, scale = F.dropout_mask(1234, drop_p, B, D, "cuda").cpu(), 1.0 / (1.0 - drop_p)
        ref = orc.np_head_forward_backward(p, img.numpy(), txt.numpy(), rev, fo, co, labels=labels.numpy(),
                                           drop_mask=None if mask is None else mask.numpy(), drop_scale=scale)
        names = g.head_param_names(fo, co)
        params = [p[n].cuda().requires_grad_(True) for n in names]
        logits = g.mmrca_head(img.cuda(), txt.cuda(), params, reverse=rev, features_only=fo, cross_attention_only=co,
                              compute=compute, drop_p=drop_p, drop_seed=1234)
        loss = CrossEntropyLoss()(logits, labels.cuda())
        loss.backward()
        torch.cuda.synchronize()
        ours = {n: t.grad.cpu().numpy() for n, t in zip(names, params)}
        al, ag = autocast_grads(p, img, txt, labels, rev, fo, co, mask, scale)
        so, sa = grad_summary(ours, ref["grads"]), grad_summary(ag, ref["grads"])
        print(f"== qk_gain {qk} flags {flags} compute {compute} drop {drop_p}: logits max abs err "
              f"{np.abs(logits.detach().cpu().numpy() - ref['logits']).max():.3e} (autocast {np.abs(al - ref['logits']).max():.3e}) "
              f"loss err {abs(loss.item() - ref['loss']):.2e}")
        print(f"   ours    : flat l2 rel {so['flat_l2_rel']:.3e} cos {so['cos']:.5f} worst tensor l2 rel {so['worst_l2_rel']:.3e} ({so['worst_name']}) "
              f"max err/global {so['max_err_over_global']:.3e}")
        print(f"   autocast: flat l2 rel {sa['flat_l2_rel']:.3e} cos {sa['cos']:.5f} worst tensor l2 rel {sa['worst_l2_rel']:.3e} ({sa['worst_name']}) "
              f"max err/global {sa['max_err_over_global']:.3e}")
        if verbose:
            scale_g = max(np.abs(v).max() for v in ref["grads"].values())
            for n in names:
                r, o = ref["grads"][n], ours[n]
                print(f"   {n:42s} max|ref|/global {np.abs(r).max() / scale_g:.1e}  l2err/l2 "
                      f"{np.linalg.norm(o - r) / max(np.linalg.norm(r), 1e-30):.2e}  (autocast "
                      f"{np.linalg.norm(ag[n] - r) / max(np.linalg.norm(r), 1e-30):.2e})")
