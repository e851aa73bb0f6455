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
    if args.late_fusion == "classic":           # reference main_both.py:284-292
        return mm.EffV2MediumAndDistilbertClassic(*common, pretrained=pretrained)
    if args.late_fusion == "normalized":        # reference main_both.py:294-302
        return mm.EffV2MediumAndDistilbertNormalized(*common, pretrained=pretrained)
    raise SystemExit(f"late fusion strategy {args.late_fusion!r} is outside the B200-native hot path "
                     "(MM_RCA, hierarchical, classic and normalized are; see SURVEY.md §8)")


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

    def __call__(self, img, txt, labels, drop_mask=None, drop_scale=1.0, sync: bool = True, drop_seed: int = 0,
                 zero_grad: bool = False):
        loss, logits = self.step(img, txt, labels, drop_mask, drop_scale, drop_seed, zero_grad=zero_grad)
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


class FlatParams:
    """One contiguous fp32 buffer holding the head parameters themselves, laid out exactly like FlatGrads lays out their
    gradients; every nn.Parameter's .data becomes a view of it (state_dict, .to(), strict loading are unaffected: a view
    serialises as its own tensor).  Together with FlatGrads it lets ONE kernel update all 34 tensors (FusedSGD /
    FusedAdamW) and one collective reduce all their gradients."""

    def __init__(self, params: Sequence[nn.Parameter], layout: F.FlatGrads):
        self.flat = torch.zeros_like(layout.flat)
        self.views = []
        for p, off in zip(params, layout.offsets):
            v = self.flat[off:off + p.numel()].view(p.shape)
            v.copy_(p.detach())
            p.data = v
            self.views.append(v)
        self.n = layout.n


class _FusedOptimizer:
    """Base of the fused optimizers over (FlatParams, FlatGrads).  param_groups / zero_grad mirror what the reference's
    loop touches (main_both.py:113-121 step + zero_grad, :700-703 lr rescale, ReduceLROnPlateau writes group['lr'])."""

    def __init__(self, flat_params: FlatParams, flat_grads: F.FlatGrads, **defaults):
        if flat_params.flat.numel() != flat_grads.flat.numel():
            raise ValueError("parameter and gradient buckets differ in layout")
        self.p, self.g = flat_params, flat_grads
        self.param_groups = [dict(defaults, params=[])]
        self.defaults = dict(defaults)
        self.state: dict = {}
        self.steps = 0

    def zero_grad(self, set_to_none: bool = False):
        self.g.zero_()

    def _stream(self):
        return torch.cuda.current_stream(self.p.flat.device).cuda_stream


class FusedSGD(_FusedOptimizer):
    """torch.optim.SGD(lr, momentum, dampening, weight_decay, nesterov) on the flat head bucket: one launch."""

    def __init__(self, flat_params, flat_grads, lr, momentum=0.0, dampening=0.0, weight_decay=0.0, nesterov=False):
        super().__init__(flat_params, flat_grads, lr=lr, momentum=momentum, dampening=dampening,
                         weight_decay=weight_decay, nesterov=nesterov)
        self.buf = torch.zeros_like(flat_params.flat) if momentum != 0.0 else None

    @torch.no_grad()
    def step(self):
        g = self.param_groups[0]
        with torch.cuda.device(self.p.flat.device):
            N.check(N.lib().mmrca_sgd_step(self.p.flat.data_ptr(), self.g.flat.data_ptr(),
                                           self.buf.data_ptr() if self.buf is not None else None, self.p.flat.numel(),
                                           float(g["lr"]), float(g["momentum"]), float(g["dampening"]),
                                           float(g["weight_decay"]), int(bool(g["nesterov"])), int(self.steps == 0),
                                           self._stream()), "mmrca_sgd_step")
        self.steps += 1


class FusedAdamW(_FusedOptimizer):
    """torch.optim.AdamW(lr, betas, eps, weight_decay) on the flat head bucket: one launch."""

    def __init__(self, flat_params, flat_grads, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2):
        super().__init__(flat_params, flat_grads, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
        self.m = torch.zeros_like(flat_params.flat)
        self.v = torch.zeros_like(flat_params.flat)

    @torch.no_grad()
    def step(self):
        g = self.param_groups[0]
        self.steps += 1
        with torch.cuda.device(self.p.flat.device):
            N.check(N.lib().mmrca_adamw_step(self.p.flat.data_ptr(), self.g.flat.data_ptr(), self.m.data_ptr(),
                                             self.v.data_ptr(), self.p.flat.numel(), float(g["lr"]), float(g["betas"][0]),
                                             float(g["betas"][1]), float(g["eps"]), float(g["weight_decay"]), self.steps,
                                             self._stream()), "mmrca_adamw_step")


class FeatureCache:
    """Pooled backbone features of the frozen (transfer-learning) phase, kept on the device by dataset index, so that
    epochs after the first skip the two backbones (> 99.99 % of a step's FLOPs).  The reference recomputes them every
    step (main_both.py:562-577).  Only valid while the backbones are frozen AND the inputs of a sample do not change
    between epochs (no random augmentation / modality dropout): opt-in, off by default.  bf16 storage is what the bf16
    head takes directly (MMRCA_FLAG_FEATURES_BF16); 4 KB per sample."""

    def __init__(self, n_samples: int, d_img: int, d_txt: int, device, dtype: torch.dtype = torch.bfloat16):
        self.img = torch.empty(n_samples, d_img, dtype=dtype, device=device)
        self.txt = torch.empty(n_samples, d_txt, dtype=dtype, device=device)
        self.valid = torch.zeros(n_samples, dtype=torch.bool, device=device)
        self.hits = self.misses = 0

    def lookup(self, ids: torch.Tensor):
        """(img, txt) for `ids` if every one of them is cached, else None (one host sync on the validity bits)."""
        ids = ids.to(self.valid.device)
        if bool(self.valid[ids].all()):
            self.hits += int(ids.numel())
            return self.img[ids], self.txt[ids]
        self.misses += int(ids.numel())
        return None

    def store(self, ids: torch.Tensor, img: torch.Tensor, txt: torch.Tensor) -> None:
        ids = ids.to(self.valid.device)
        self.img[ids] = img.detach().to(self.img.dtype)
        self.txt[ids] = txt.detach().to(self.txt.dtype)
        self.valid[ids] = True

    def invalidate(self) -> None:
        self.valid.zero_()


# ---- evaluation, reference semantics --------------------------------------------------------------------
# main_both.py:43-47: the three evaluation modes of calculate_set_accuracy
mode_config_dict = {
    "image_only": {"remove_text": True, "remove_image": False},
    "text_only": {"remove_text": False, "remove_image": True},
    "both": {"remove_text": False, "remove_image": False},
}


def calculate_set_accuracy(model, data_loader, len_data, device, batch_size, mode, eval_mode, verbose: bool = False):
    """Mirror of calculate_set_accuracy (reference main_both.py:141-198): accuracy of `model` over a loader in one of the
    modes of mode_config_dict (`mode` is the dict, like in the reference's calls :596-652), predictions by argmax.  Returns
    (accuracy in percent, per-class report dict).  The report is computed here (precision / recall / f1 / support per
    class + accuracy) instead of importing sklearn."""
    n_batches = math.ceil(len_data / batch_size)
    all_labels, all_predictions = [], []
    correct = 0
    with torch.no_grad():
        for batch_idx, (data, labels) in enumerate(data_loader):
            ids = data["text"]["tokens"].to(device)
            mask = data["text"]["attention_mask"].to(device)
            images = data["image"]["raw_image"].to(device)
            labels = labels.to(device)
            outputs = model(_input_ids=ids, _attention_mask=mask, _images=images, eval=eval_mode,
                            remove_text=mode["remove_text"], remove_image=mode["remove_image"])
            pred = torch.max(outputs, 1)[1].view(-1)
            correct += torch.sum(torch.eq(pred, labels)).item()
            if verbose:
                print("Batches {}/{} ".format(batch_idx, n_batches))
            all_labels.append(labels.cpu())
            all_predictions.append(pred.cpu())
    y = torch.cat(all_labels) if all_labels else torch.zeros(0, dtype=torch.long)
    yhat = torch.cat(all_predictions) if all_predictions else torch.zeros(0, dtype=torch.long)
    report = {}
    for c, name in enumerate(["black", "blue", "green", "ttr"]):
        tp = int(((yhat == c) & (y == c)).sum())
        fp = int(((yhat == c) & (y != c)).sum())
        fn = int(((yhat != c) & (y == c)).sum())
        prec = tp / (tp + fp) if tp + fp else 0.0
        rec = tp / (tp + fn) if tp + fn else 0.0
        report[name] = {"precision": prec, "recall": rec,
                        "f1-score": 2 * prec * rec / (prec + rec) if prec + rec else 0.0, "support": tp + fn}
    report["accuracy"] = correct / max(1, int(y.numel()))
    acc = 100 * (correct / len_data)
    return acc, report


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
