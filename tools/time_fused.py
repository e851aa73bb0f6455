"""Per-kernel CUDA-event timing of the fused bf16 forward at batch B (development aid)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import garbage_classification_rca_b200 as g
from garbage_classification_rca_b200 import _native as N, functional as F

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
compute = int(sys.argv[2]) if len(sys.argv) > 2 else N.COMPUTE_BF16_FUSED
params = F.init_head_parameters("cuda", seed=0)
desc = N.HeadDesc(B, 1280, 768, 4, F.make_flags(True, False, False), compute)
L = N.lib()
ws = torch.zeros(int(L.mmrca_head_workspace_bytes(C.byref(desc), 0)), dtype=torch.uint8, device="cuda")
logits = torch.empty(B, 4, device="cuda")
hp = F._head_struct(params)
NB = 8
img = torch.randn(NB, B, 1280, device="cuda"); txt = torch.randn(NB, B, 768, device="cuda")
st = torch.cuda.current_stream().cuda_stream
def fwd(i):
    N.check(L.mmrca_head_forward(C.byref(desc), C.byref(hp), img[i % NB].data_ptr(), txt[i % NB].data_ptr(), None, 1.0,
                                 logits.data_ptr(), ws.data_ptr(), ws.numel(), st), "fwd")
for i in range(5): fwd(i)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(20): fwd(i)
e1.record(); torch.cuda.synchronize()
print(f"forward B={B}: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us/call")
N.timing_begin(512)
for i in range(10): fwd(i)
acc = {}
for name, ms in N.timing_end(512): acc.setdefault(name, []).append(ms)
for k, v in acc.items(): print(f"  {k:28s} {sum(v) / len(v) * 1e3:8.1f} us  x{len(v) // 10}/call")
