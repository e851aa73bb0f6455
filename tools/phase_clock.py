"""Development aid: per-phase clock64 stamps of CTA 0 of the instrumented tile kernel (mmrca_dev_set_debug)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import garbage_classification_rca_b200 as g
from garbage_classification_rca_b200 import _native as N, functional as F
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
kern = int(sys.argv[2]) if len(sys.argv) > 2 else 0      # 0: SA backward, 1: CA backward
per = 16 if kern == 1 else 12
rounds = 4 if kern >= 2 else 7
params = F.init_head_parameters("cuda", seed=0)
step = g.HeadTrainStep(params, B, 1280, 768, reverse=True, compute=N.COMPUTE_BF16, drop_p=float(os.environ.get("DROP_P", "0.6")))
img = torch.randn(B, 1280, device="cuda"); txt = torch.randn(B, 768, device="cuda")
lab = torch.randint(0, 4, (B,), device="cuda")
dbg = torch.zeros(1024, dtype=torch.int64, device="cuda")
for i in range(3):
    step.zero_grad(); step(img, txt, lab, drop_seed=i)
N.lib().mmrca_dev_set_debug(dbg.data_ptr(), kern)
step.zero_grad(); step(img, txt, lab, drop_seed=9)
torch.cuda.synchronize()
N.lib().mmrca_dev_set_debug(None, 0)
d = dbg.cpu().tolist()
t0 = d[0] if kern < 2 else d[1]
print("stamps (cycles since kernel start of CTA 0):")
vals = [v - t0 for v in d if v]
print(vals[:1 + rounds * per + 2])
for t in range(rounds):
    row = d[1 + t * per: 1 + (t + 1) * per]
    if not row[0]: break
    print(f"tile {t}: start {row[0] - t0:7d}  deltas", [row[i + 1] - row[i] for i in range(per - 1) if row[i + 1]])
ns = d[251] - d[250]
cyc = max(v for v in d[:250]) - d[0]
ncta = 148
print(f"CTA 0: {cyc} cycles in {ns} ns -> SM clock {cyc / ns * 1e3:.0f} MHz")
ent = [d[300 + 2 * i] for i in range(148)]; ex = [d[301 + 2 * i] for i in range(148)]
t0 = min(ent)
print("CTA entry ns (min..max):", 0, max(ent) - t0, " main-loop start of CTA 0:", d[250] - t0, " exits (min..max):", min(ex) - t0, max(ex) - t0)
print("per-CTA duration ns: ", sorted(e - s for s, e in zip(ent, ex))[::12])

