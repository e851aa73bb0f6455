"""Development aid: per-tensor errors of the hierarchical head against the float64 oracle."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import garbage_classification_rca_b200 as g
from oracle import mmrca_oracle as orc
B = int(sys.argv[1]) if len(sys.argv) > 1 else 5
gen = torch.Generator().manual_seed(1)
feats = [torch.randn(B, w, generator=gen) for w in (1280, 2560, 2048, 768, 768, 768)]
labels = torch.randint(0, 4, (B,), generator=gen)
p = orc.init_hier_params(seed=0)
params = [p[n].cuda() for n in g.functional.HIER_PARAM_NAMES]
step = g.HierTrainStep(params, B)
step.zero_grad()
loss, logits = step([f.cuda() for f in feats], labels.cuda())
rl, rloss, rg = orc.hier_loss_and_grads(p, feats[:3], feats[3:], labels)
print("logits err", (logits.cpu() - rl.float()).abs().max().item(), "loss", loss.item(), rloss.item())
for n, v in zip(g.functional.HIER_PARAM_NAMES, step.grads.views):
    a, r = v.cpu().double().numpy(), rg[n].numpy()
    err = np.abs(a - r)
    i = np.unravel_index(err.argmax(), err.shape)
    print(f"{n:36s} max|ref| {np.abs(r).max():.3e}  max err {err.max():.3e} at {i}  got {a[i]:.4e} ref {r[i]:.4e}  corr {np.corrcoef(a.ravel(), r.ravel())[0,1]:.5f}")
if B <= 128:
    a, r = step.grads.views[0].cpu().double().numpy(), rg[g.functional.HIER_PARAM_NAMES[0]].numpy()
    rowerr = np.abs(a - r).max(1); colerr = np.abs(a - r).max(0)
    print("dW_img row err (first 16 / by 128-blocks):", rowerr[:16].round(6), [rowerr[i:i+128].max().round(6) for i in range(0, 512, 128)])
    print("dW_img col err by 256-blocks:", [colerr[i:i+256].max().round(6) for i in range(0, 5888, 256)])
    print("col err within first 32:", colerr[:32].round(6))
