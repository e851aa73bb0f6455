"""torchrun --nproc-per-node N tools/check_peer_allreduce.py : the peer-memory all-reduce against NCCL, and their times."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from garbage_classification_rca_b200.training import PeerAllReduce, allreduce_mean_
rank, lr = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
n = 94820 // 4 * 4 + 4
ar = PeerAllReduce(n, dev)
worst = 0.0
for it in range(40):
    g = torch.Generator(device=dev).manual_seed(1000 * it + rank)
    x = torch.randn(n, device=dev, generator=g)
    ref = x.clone(); allreduce_mean_(ref)
    out = ar(x.clone())
    worst = max(worst, (out - ref).abs().max().item())
    chk = out.clone(); dist.broadcast(chk, 0)
    assert torch.equal(chk, out), "ranks disagree bitwise"
x = torch.randn(n, device=dev)
for name, fn in (("peer", lambda: ar(x)), ("nccl", lambda: allreduce_mean_(x))):
    for _ in range(10): fn()
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(200): fn()
    e1.record(); torch.cuda.synchronize()
    if rank == 0: print(f"{name}: {e0.elapsed_time(e1) / 200 * 1e3:.1f} us per all-reduce of {n * 4 / 1e3:.0f} KB over {dist.get_world_size()} GPUs", file=sys.stderr)
if rank == 0: print(f"max |peer - nccl| = {worst:.3e}", file=sys.stderr)
dist.destroy_process_group()
