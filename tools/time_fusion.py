"""Development: time of one classic / normalized head training step (forward + CE + backward) at batch B."""
import sys
sys.path.insert(0, ".")
import torch
from garbage_classification_rca_b200 import _native as N, functional as F
from oracle import mmrca_oracle as orc

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
for normalized, compute in ((False, "fp32"), (True, "fp32"), (False, "bf16"), (True, "bf16")):
    p = orc.init_fusion_params(seed=1)
    params = [p[n].cuda() for n in F.FUSION_PARAM_NAMES]
    step = F.FusionTrainStep(params, B, normalized=normalized, drop_p=0.6, compute=compute)
    g = torch.Generator().manual_seed(0)
    img, txt = torch.randn(B, 1280, generator=g).cuda(), torch.randn(B, 768, generator=g).cuda()
    labels = torch.randint(0, 4, (B,), generator=g).cuda()
    for i in range(3):
        step.zero_grad(); step(img, txt, labels, drop_seed=i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(10):
        step.zero_grad(); step(img, txt, labels, drop_seed=i)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    N.timing_begin(128)
    step.zero_grad(); step(img, txt, labels, drop_seed=0)
    recs = N.timing_end(128)
    print(f"normalized={normalized} compute={compute} B={B}: {ms * 1e3:.1f} us/step = {B / ms / 1e3:.2f} M samples/s; " + " ".join(f"{n}={t * 1e3:.0f}" for n, t in recs))
