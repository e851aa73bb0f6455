"""Training-loop pieces around the hot path: the reference's loss / backward / optimizer-step semantics
(main_both.py:81-134), model dispatch (main_both.py:272-343), checkpoint save (main_both.py:201-226)
and the data-parallel plumbing the reference never had (it only wraps nn.DataParallel and that path
crashes): one process per GPU, batch-sharded, ONE all-reduce of the flat head-gradient
bucket per optimizer step.
"""
from __future__ import annotations

import math
import os
from typing import Iterable, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist
from torch import nn

from . import _native as N
from . import functional as F


# ---- loss -----------------------------------------------------------------------------------------------
class _CrossEntropyFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, labels, weight, label_smoothing):
        loss, dlogits = F.cross_entropy(logits, labels, weight, label_smoothing, want_dlogits=True)
        ctx.save_for_backward(dlogits)
        return loss.reshape(())

    @staticmethod
    def backward(ctx, grad_out):
        (dlogits,) = ctx.saved_tensors
        return dlogits * grad_out, None, None, None


class CrossEntropyLoss(nn.Module):
    """torch.nn.CrossEntropyLoss(weight=, label_smoothing=) with mean reduction (what run_one_epoch builds,
    reference main_both.py:87-93), as one fused forward+dlogits kernel."""

    def __init__(self, weight: Optional[torch.Tensor] = None, label_smoothing: float = 0.0):
        super().__init__()
        self.register_buffer("weight", weight)
        self.label_smoothing = label_smoothing

    def forward(self, logits, labels):
        return _CrossEntropyFunction.apply(logits, labels, self.weight, self.label_smoothing)


# ---- model dispatch -------------------------------------------------------------------------------------
def build_model(args, n_classes: int = 4, pretrained: bool = True):
    """--late_fusion dispatch (reference main_both.py:272-343): MM_RCA and hierarchical are rebuilt here."""
    from . import multimodal_model as mm
    compute = N.COMPUTE_BF16 if getattr(args, "compute", "fp32") == "bf16" else N.COMPUTE_FP32
    common = (n_classes, args.model_dropout, args.image_text_dropout, args.image_prob_dropout,
              args.num_neurons_FC, args.text_model, args.batch_size, args.reverse)
    if args.late_fusion == "MM_RCA":
        return mm.MM_RCA(*common, args.features_only, args.cross_attention_only, pretrained=pretrained,
                         compute=compute)
    if args.late_fusion == "hierarchical":      # reference main_both.py:318-330
        return mm.Hierarchical(*common, args.features_only, args.cross_attention_only, pretrained=pretrained)
    raise SystemExit(f"late fusion strategy {args.late_fusion!r} is outside the B200-native hot path "
                     "(MM_RCA and hierarchical are; see SURVEY.md §8)")


# ---- data-parallel plumbing -----------------------------------------------------------------------------
def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) slice of n samples owned by `rank` (balanced to within one sample)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def allreduce_mean_(flat: torch.Tensor, group=None) -> torch.Tensor:
    """The single collective of a data-parallel head step: sum the flat gradient bucket over ranks and
    divide by the world size (equal shards + per-rank mean loss == global mean loss; with class weights
    the per-rank normaliser differs, SURVEY.md §8 e)."""
    if dist.is_available() and dist.is_initialized():
        world = dist.get_world_size(group)
        if world > 1:
            dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
            flat.div_(world)
    return flat


def allreduce_global_mean_(flat: torch.Tensor, n: int, local_den, group=None, reduce=None) -> torch.Tensor:
    """Exact global-mean gradient when the per-rank normalisers differ (class-weighted CrossEntropyLoss: each rank divided
    by its own sum of w[y_b]; or unequal shards: by its own batch).  flat[:n] holds this rank's gradient normalised by
    `local_den`; flat[n] is a spare slot.  The bucket is rescaled to the un-normalised sum, the denominator rides in the
    spare slot through the SAME single collective, and the result is sum_r den_r g_r / sum_r den_r = the gradient a single
    process gets on the concatenated batch (reference main_both.py:87-93 on the global batch).
    `reduce`: callable(flat) doing the collective (sum or mean over ranks: the ratio is the same); default all-reduce."""
    if flat.numel() <= n:
        raise ValueError("the bucket needs one spare float after its n gradient entries")
    flat[:n].mul_(local_den)
    flat[n] = local_den
    if reduce is not None:
        reduce(flat)
    elif dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    flat[:n].div_(flat[n])
    return flat


class PeerAllReduce:
    """The step's one collective as a one-shot all-reduce over NVLink peer memory (csrc/mmrca_peer.cuh): every rank
    reads the W staged copies of the 379 KB bucket itself.  torch's symmetric memory does the plumbing (allocation, IPC
    exchange, mapping); the reduction is libmmrca.so's kernel.  Construction is a collective over `group`; it raises
    if symmetric memory is unavailable (the caller then keeps NCCL)."""

    def __init__(self, numel: int, device, group=None):
        import torch.distributed._symmetric_memory as symm_mem
        group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.n = numel
        self.n_pad = (numel + 63) // 64 * 64
        pad_words = N.lib().mmrca_peer_allreduce_pad_bytes(self.world) // 4
        self.pad_off = 2 * self.n_pad                                   # [staging half 0 | half 1 | flag pad] in one buffer
        self.buf = symm_mem.empty(self.pad_off + pad_words, dtype=torch.float32, device=device)
        self.buf.zero_()
        self.handle = symm_mem.rendezvous(self.buf, group)
        ptrs = [int(p) for p in self.handle.buffer_ptrs]
        self.staging = (N._fp * self.world)(*ptrs)
        self.pads = (N._fp * self.world)(*[p + 4 * self.pad_off for p in ptrs])
        self.step = 0
        torch.cuda.synchronize(device)
        dist.barrier(group)                                             # every pad is zero before anyone publishes a flag

    def __call__(self, flat: torch.Tensor) -> torch.Tensor:
        if flat.numel() != self.n or flat.dtype != torch.float32 or not flat.is_contiguous():
            raise ValueError("bucket does not match the one this all-reduce was built for")
        self.step += 1
        with torch.cuda.device(flat.device):
            N.check(N.lib().mmrca_peer_allreduce_mean(flat.data_ptr(), self.n, self.n_pad, self.staging, self.pads,
                                                      self.rank, self.world, self.step,
                                                      torch.cuda.current_stream(flat.device).cuda_stream),
                    "mmrca_peer_allreduce_mean")
        return flat

    def status(self) -> int:
        """0 = every peer arrived so far; 1 + r = rank r never published its flags within the kernel's bounded wait (the
        bucket was left unreduced); synchronises the current stream."""
        dev = self.buf.device
        with torch.cuda.device(dev):
            rc = N.lib().mmrca_peer_allreduce_status(self.pads[self.rank], self.world,
                                                     torch.cuda.current_stream(dev).cuda_stream)
        if rc < 0:
            N.check(-rc, "mmrca_peer_allreduce_status")
        return rc


class HeadDataParallel:
    """Batch-sharded training of the fusion head on precomputed features: each rank runs the one-call
    train step on its shard, then ONE all-reduce of `step.grads.flat` (94 820 floats, 379 KB): over NVLink peer
    memory (PeerAllReduce) when `peer=True` and symmetric memory is available, through NCCL / gloo otherwise."""

    def __init__(self, step, group=None, peer: bool = False):
        self.step = step
        self.group = group
        self.peer = None
        self.collective = "torch.distributed all_reduce"
        # Averaging per-rank gradients by 1/world is the global mean only when every rank normalised by the same
        # denominator.  With class weights (sum of w[y_b] differs per rank) or unequal shards the denominators ride along
        # in the bucket's spare slot (allreduce_global_mean_): still ONE collective per step.
        self.exact_mean = getattr(step, "cw", None) is not None
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            sizes = [None] * dist.get_world_size(group)
            dist.all_gather_object(sizes, int(step.desc.batch), group=group)
            self.exact_mean = self.exact_mean or len(set(sizes)) > 1
        if peer and dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1 \
                and step.grads.flat.is_cuda:
            ok = torch.zeros(1, device=step.grads.flat.device)
            try:
                self.peer = PeerAllReduce(step.grads.flat.numel(), step.grads.flat.device, group)
                ok += 1
            except Exception as e:      # noqa: BLE001  (no symmetric memory on this box: NCCL does the job)
                self.peer_error = repr(e)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)     # all ranks or none
            if ok.item() < 1:
                self.peer = None
            else:
                self.collective = "one-shot all-reduce over NVLink peer memory (mmrca_peer_allreduce_mean)"

    def __call__(self, img, txt, labels, drop_mask=None, drop_scale=1.0, sync: bool = True, drop_seed: int = 0):
        loss, logits = self.step(img, txt, labels, drop_mask, drop_scale, drop_seed)
        if sync:   # sync=False == DDP.no_sync() while accumulating (reference steps every acc_steps batches)
            flat = self.step.grads.flat
            if self.exact_mean:
                cw = self.step.cw
                den = cw[labels].sum() if cw is not None else float(labels.shape[0])
                allreduce_global_mean_(flat, self.step.grads.n, den, self.group, reduce=self.peer)
            elif self.peer is not None:
                self.peer(flat)
            else:
                allreduce_mean_(flat, self.group)
        return loss, logits


def attach_flat_grads(params: Sequence[nn.Parameter], grads: F.FlatGrads) -> None:
    """Point every head parameter's .grad at its view of the flat bucket so a stock torch optimizer
    consumes the kernels' output without copies."""
    for p, v in zip(params, grads.views):
        p.grad = v


# ---- one epoch, reference semantics ---------------------------------------------------------------------
def run_one_epoch(epoch_num, model, data_loader, len_train_data, hw_device, batch_size, train_optimizer, weights,
                  use_class_weights, acc_steps, smoothing, verbose: bool = False):
    """Mirror of run_one_epoch (reference main_both.py:81-134) including its quirks: gradients are NOT
    scaled by 1/acc_steps (the division happens after backward, :112-114), the optimizer steps every
    acc_steps batches and on the last batch, and the loss is read back every batch (:128-130)."""
    n_batches = math.ceil(len_train_data / batch_size)
    cw = torch.tensor(weights, dtype=torch.float32, device=hw_device) if use_class_weights else None
    criterion = CrossEntropyLoss(weight=cw, label_smoothing=smoothing)
    batch_loss: List[torch.Tensor] = []
    n_total = len(data_loader)
    for batch_idx, (data, labels) in enumerate(data_loader):
        ids = data["text"]["tokens"].to(hw_device)
        mask = data["text"]["attention_mask"].to(hw_device)
        images = data["image"]["raw_image"].to(hw_device)
        labels = labels.to(hw_device)
        out = model(_input_ids=ids, _attention_mask=mask, _images=images)
        loss = criterion(out, labels)
        loss.backward()
        if acc_steps != 0:
            loss = loss / acc_steps
            if (batch_idx + 1) % acc_steps == 0 or batch_idx + 1 == n_total:
                train_optimizer.step()
                train_optimizer.zero_grad()
        else:
            train_optimizer.step()
            train_optimizer.zero_grad()
        if verbose:
            print("Batch {}/{} on epoch {}".format(batch_idx, n_batches, epoch_num))
        batch_loss.append(loss.detach().cpu())
    return n_batches, batch_loss


def save_model_weights(model: nn.Module, path: str, device) -> str:
    """state_dict-only checkpoint written from the CPU copy (reference main_both.py:201-226)."""
    os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
    model.to("cpu")
    torch.save(model.state_dict(), path)
    model.to(device)
    return path
