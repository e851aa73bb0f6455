"""CPU, world_size 2, gloo: the host-side logic of the data-parallel head step — contiguous batch
sharding plus ONE averaged all-reduce of the flat gradient bucket reproduces the single-process
full-batch gradient.  The per-rank compute is stood in for by the oracle (test infrastructure); on the
GPU box the same plumbing runs over NCCL with the CUDA train step (bench.py --gpus N)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import mmrca_oracle as orc
from tests._util import make_inputs


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, B, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from garbage_classification_rca_b200.training import allreduce_mean_, shard_range
    torch.set_num_threads(2)
    p = orc.init_head_params(seed=4, qk_gain=40.0)
    img, txt, labels = make_inputs(B, 4)
    lo, hi = shard_range(B, rank, world)
    _, loss, grads, _, _ = orc.head_loss_and_grads(p, img[lo:hi], txt[lo:hi], labels[lo:hi], True, False, False)
    names = orc.head_param_names()
    flat = torch.cat([grads[n].reshape(-1) for n in names])
    allreduce_mean_(flat)
    if rank == 0:
        q.put((flat.numpy(), (lo, hi)))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_flat_allreduce_equals_full_batch_gradient():
    B, world = 16, 2   # equal shards: per-rank mean loss averaged over ranks == global mean loss
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, B, q)) for r in range(world)]
    for pr in procs:
        pr.start()
    flat, (lo, hi) = q.get(timeout=240)
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    assert (lo, hi) == (0, 8)
    p = orc.init_head_params(seed=4, qk_gain=40.0)
    img, txt, labels = make_inputs(B, 4)
    _, _, grads, _, _ = orc.head_loss_and_grads(p, img, txt, labels, True, False, False)
    ref = torch.cat([grads[n].reshape(-1) for n in orc.head_param_names()]).numpy()
    assert np.abs(flat - ref).max() / np.abs(ref).max() < 1e-5


def _worker_weighted(rank, world, port, B, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from garbage_classification_rca_b200.training import allreduce_global_mean_, shard_range
    torch.set_num_threads(2)
    p = orc.init_head_params(seed=4, qk_gain=40.0)
    img, txt, labels = make_inputs(B, 4)
    cw = torch.tensor([0.5, 2.0, 1.0, 0.7])
    lo, hi = shard_range(B, rank, world)          # B = 13, world 2: shards of 7 and 6 samples
    _, _, grads, _, _ = orc.head_loss_and_grads(p, img[lo:hi], txt[lo:hi], labels[lo:hi], True, False, False,
                                                class_weight=cw, label_smoothing=0.1)
    names = orc.head_param_names()
    g = torch.cat([grads[n].reshape(-1) for n in names])
    flat = torch.cat([g, torch.zeros(4)])
    allreduce_global_mean_(flat, g.numel(), cw[labels[lo:hi]].sum())
    if rank == 0:
        q.put(flat[:g.numel()].numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_class_weighted_uneven_shards_give_the_global_weighted_mean_gradient():
    """Per-rank CrossEntropyLoss(weight=) normalises by the rank's own sum of w[y]; with unequal shards on top, a plain
    1/world average is not the single-process gradient.  The denominators ride in the bucket's spare slot."""
    B, world = 13, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_weighted, args=(r, world, port, B, q)) for r in range(world)]
    for pr in procs:
        pr.start()
    flat = q.get(timeout=240)
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    p = orc.init_head_params(seed=4, qk_gain=40.0)
    img, txt, labels = make_inputs(B, 4)
    cw = torch.tensor([0.5, 2.0, 1.0, 0.7])
    _, _, grads, _, _ = orc.head_loss_and_grads(p, img, txt, labels, True, False, False, class_weight=cw,
                                                label_smoothing=0.1)
    ref = torch.cat([grads[n].reshape(-1) for n in orc.head_param_names()]).numpy()
    assert np.abs(flat - ref).max() / np.abs(ref).max() < 1e-5


def test_shard_range_covers_batch_without_overlap():
    from garbage_classification_rca_b200.training import shard_range
    for n in (0, 1, 7, 16, 4096, 4099):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
