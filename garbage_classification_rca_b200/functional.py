"""Host-side entry points of the MM-RCA fusion head: torch tensors in, libmmrca.so (C ABI) underneath.

PyTorch is plumbing here (device memory, current stream, autograd bookkeeping); all arithmetic of
the hot path — CVPR_code/multimodal_model.py:661-728 of the reference, its CrossEntropyLoss
(main_both.py:87-93) and the backward through both (main_both.py:112) — runs in the sm_100a kernels.
Nothing here falls back to torch ops or the CPU: non-CUDA inputs raise.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import _native as N

NUM_PATCHES = 16
SA_DKQ, SA_DV, CA_DKQ, CA_DV = 128, 96, 64, 48

_ATTN_LEAVES = ("W_query.weight", "W_query.bias", "W_key.weight", "W_key.bias",
                "W_value.weight", "W_value.bias", "norm.weight", "norm.bias")
_ATTN_FIELDS = ("wq", "bq", "wk", "bk", "wv", "bv", "ln_g", "ln_b")
# order of the attention blocks inside MmrcaHeadParams
_BLOCKS = (("sa_img", "self_attention_image"), ("sa_txt", "self_attention_text"),
           ("ca1", "cross_attention_1"), ("ca2", "cross_attention_2"))


def final_linear_name(features_only: bool, cross_attention_only: bool) -> str:
    """Classifier used by the forward — reference multimodal_model.py:721-726."""
    if features_only:
        return "final_features_only_linear"
    if cross_attention_only:
        return "cross_attention_only_linear"
    return "final_with_everything"


def head_param_names(features_only: bool = False, cross_attention_only: bool = False) -> List[str]:
    """state_dict names of the 34 tensors MM_RCA.forward reads, in the order of MmrcaHeadParams."""
    names = [f"{mod}.{leaf}" for _, mod in _BLOCKS for leaf in _ATTN_LEAVES]
    fin = final_linear_name(features_only, cross_attention_only)
    return names + [f"{fin}.weight", f"{fin}.bias"]


def head_param_shapes(d_img: int = 1280, d_txt: int = 768, n_classes: int = 4, features_only: bool = False,
                      cross_attention_only: bool = False) -> List[Tuple[str, Tuple[int, ...]]]:
    """(name, shape) of the head tensors in head_param_names order (reference multimodal_model.py:249-292)."""
    out = []
    dims = {"self_attention_image": (d_img // NUM_PATCHES, SA_DKQ, SA_DV),
            "self_attention_text": (d_txt // NUM_PATCHES, SA_DKQ, SA_DV),
            "cross_attention_1": (SA_DV, CA_DKQ, CA_DV), "cross_attention_2": (SA_DV, CA_DKQ, CA_DV)}
    for _, mod in _BLOCKS:
        d_in, d_kq, d_v = dims[mod]
        out += [(f"{mod}.W_query.weight", (d_kq, d_in)), (f"{mod}.W_query.bias", (d_kq,)),
                (f"{mod}.W_key.weight", (d_kq, d_in)), (f"{mod}.W_key.bias", (d_kq,)),
                (f"{mod}.W_value.weight", (d_v, d_in)), (f"{mod}.W_value.bias", (d_v,)),
                (f"{mod}.norm.weight", (d_v,)), (f"{mod}.norm.bias", (d_v,))]
    fin = final_linear_name(features_only, cross_attention_only)
    D = concat_width(d_img, d_txt, features_only, cross_attention_only)
    return out + [(f"{fin}.weight", (n_classes, D)), (f"{fin}.bias", (n_classes,))]


def init_head_parameters(device, d_img: int = 1280, d_txt: int = 768, n_classes: int = 4,
                         features_only: bool = False, cross_attention_only: bool = False,
                         seed: int = 0) -> List[torch.Tensor]:
    """Random-init head tensors with the distributions torch.nn.Linear / LayerNorm use in the reference
    ctor (uniform(-1/sqrt(fan_in), 1/sqrt(fan_in)); LayerNorm weight 1, bias 0)."""
    g = torch.Generator().manual_seed(seed)
    ps = []
    for name, shape in head_param_shapes(d_img, d_txt, n_classes, features_only, cross_attention_only):
        if name.endswith("norm.weight"):
            t = torch.ones(shape)
        elif name.endswith("norm.bias"):
            t = torch.zeros(shape)
        else:
            fan_in = shape[1] if len(shape) == 2 else dict(head_param_shapes(
                d_img, d_txt, n_classes, features_only, cross_attention_only))[name[:-4] + "weight"][1]
            k = 1.0 / (fan_in ** 0.5)
            t = (torch.rand(shape, generator=g) * 2 - 1) * k
        ps.append(t.to(device))
    return ps


def concat_width(d_img: int, d_txt: int, features_only: bool, cross_attention_only: bool) -> int:
    ca = 2 * NUM_PATCHES * CA_DV
    if features_only:
        return d_img + d_txt
    if cross_attention_only:
        return ca
    return ca + d_img + d_txt


def make_flags(reverse: bool, features_only: bool, cross_attention_only: bool) -> int:
    return ((N.FLAG_REVERSE if reverse else 0) | (N.FLAG_FEATURES_ONLY if features_only else 0)
            | (N.FLAG_CROSS_ATTENTION_ONLY if cross_attention_only else 0))


def _stream_ptr(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def _check_dev(t: torch.Tensor, what: str, dtype=torch.float32) -> torch.Tensor:
    if not t.is_cuda:
        raise RuntimeError(f"{what} must be a CUDA tensor: the MM-RCA head has no CPU fallback")
    if t.dtype != dtype:
        raise TypeError(f"{what} must be {dtype}, got {t.dtype}")
    return t if t.is_contiguous() else t.contiguous()


def _head_struct(tensors: Sequence[Optional[torch.Tensor]]) -> N.HeadParams:
    """Pack 34 tensors (order of head_param_names) into the MmrcaHeadParams / MmrcaHeadGrads layout."""
    hp = N.HeadParams()
    it = iter(tensors)
    for field, _ in _BLOCKS:
        blk = getattr(hp, field)
        for leaf in _ATTN_FIELDS:
            t = next(it)
            setattr(blk, leaf, t.data_ptr() if t is not None else None)
    for leaf in ("wf", "bf"):
        t = next(it)
        setattr(hp, leaf, t.data_ptr() if t is not None else None)
    return hp


def _desc(batch: int, d_img: int, d_txt: int, n_classes: int, flags: int, compute: int,
          drop_p: float = 0.0, drop_seed: int = 0) -> N.HeadDesc:
    return N.HeadDesc(batch, d_img, d_txt, n_classes, flags, compute, float(drop_p), int(drop_seed) & (2 ** 64 - 1))


def dropout_mask(seed: int, p: float, batch: int, width: int, device) -> torch.Tensor:
    """uint8 keep-mask [batch, width] the seeded dropout of the head draws for (seed, p) — what the kernels
    regenerate on chip (mmrca_dropout.cuh); for the fp32 kernels, parity tests and debugging."""
    out = torch.empty(batch, width, dtype=torch.uint8, device=device)
    with torch.cuda.device(device):
        N.check(N.lib().mmrca_dropout_mask(int(seed) & (2 ** 64 - 1), float(p), batch, width, out.data_ptr(),
                                           _stream_ptr(device)), "mmrca_dropout_mask")
    return out


def workspace_bytes(batch: int, d_img: int, d_txt: int, n_classes: int, flags: int, compute: int,
                    training: bool, drop_p: float = 0.0) -> int:
    d = _desc(batch, d_img, d_txt, n_classes, flags, compute, drop_p)
    return int(N.lib().mmrca_head_workspace_bytes(C.byref(d), 1 if training else 0))


class _WorkspacePool:
    """Device scratch of the autograd paths, recycled instead of allocated on every forward.  A forward checks a buffer
    out, its backward hands it back (a forward whose backward never runs simply drops its buffer)."""

    def __init__(self):
        self._free: Dict[Tuple[int, int], List[torch.Tensor]] = {}

    def take(self, nbytes: int, device) -> torch.Tensor:
        key = (torch.device(device).index or 0, int(nbytes))
        lst = self._free.get(key)
        if lst:
            return lst.pop()
        return torch.empty(max(1, nbytes), dtype=torch.uint8, device=device)

    def give(self, t: torch.Tensor) -> None:
        lst = self._free.setdefault((t.device.index or 0, t.numel()), [])
        if len(lst) < 2:
            lst.append(t)


_pool = _WorkspacePool()


class _Lease:
    """Hands a pooled workspace back when the autograd node that owns it dies (after backward, or when the graph is
    dropped) - never earlier, so backward(retain_graph=True) can run again on intact saved buffers."""

    def __init__(self, buf: torch.Tensor):
        self.buf = buf

    def __del__(self):
        try:
            _pool.give(self.buf)
        except Exception:      # noqa: BLE001  (interpreter shutdown)
            pass


def _check_mask(drop_mask: Optional[torch.Tensor], batch: int, width: int, what: str) -> Optional[torch.Tensor]:
    if drop_mask is None:
        return None
    drop_mask = _check_dev(drop_mask, "dropout mask", torch.uint8)
    if tuple(drop_mask.shape) != (batch, width):
        raise ValueError(f"{what} dropout mask has shape {tuple(drop_mask.shape)}, expected ({batch}, {width}): "
                         "[batch, concat width of the selected late-fusion variant]")
    return drop_mask


class FlatGrads:
    """One contiguous fp32 buffer holding the gradients of all head tensors (the single NCCL
    all-reduce bucket, SURVEY.md §8 e) plus per-tensor views in head_param_names order."""

    def __init__(self, params: Sequence[torch.Tensor]):
        dev = params[0].device
        # 4-float alignment per tensor keeps every view 16-byte aligned
        self.offsets, off = [], 0
        for p in params:
            self.offsets.append(off)
            off += (p.numel() + 3) // 4 * 4
        self.n = off            # gradient entries; flat[n : n + 4] is a spare slot (the loss denominator of a
        off += 4                # data-parallel step rides there, training.allreduce_global_mean_)
        self.flat = torch.zeros(off, dtype=torch.float32, device=dev)
        self.views = [self.flat[o:o + p.numel()].view(p.shape) for o, p in zip(self.offsets, params)]

    def zero_(self):
        self.flat.zero_()
        return self


class _HeadFunction(torch.autograd.Function):
    """logits = MM_RCA head(img_feat, txt_feat); backward recomputes attention internals on chip."""

    @staticmethod
    def forward(ctx, img, txt, drop_mask, drop_scale, flags, n_classes, compute, drop_p, drop_seed, grad_sink, *params):
        fdt = torch.float32
        if img.dtype == torch.bfloat16 and txt.dtype == torch.bfloat16:      # MMRCA_FLAG_FEATURES_BF16
            if compute == N.COMPUTE_FP32 or ctx.needs_input_grad[0] or ctx.needs_input_grad[1] or drop_mask is not None:
                raise TypeError("bf16 features are taken by the bf16 pipeline with frozen backbones and seeded dropout only; "
                                "pass fp32 features otherwise")
            fdt = torch.bfloat16
            flags |= N.FLAG_FEATURES_BF16
        img = _check_dev(img, "image features", fdt)
        txt = _check_dev(txt, "text features", fdt)
        params = [_check_dev(p, "head parameter") for p in params]
        ctx.grad_sink = grad_sink
        B, d_img, d_txt = img.shape[0], img.shape[1], txt.shape[1]
        if txt.shape[0] != B:
            raise ValueError("image and text feature batches differ")
        drop_mask = _check_mask(drop_mask, B, concat_width(d_img, d_txt, bool(flags & N.FLAG_FEATURES_ONLY),
                                                           bool(flags & N.FLAG_CROSS_ATTENTION_ONLY)), "head")
        if drop_mask is not None:
            compute = N.COMPUTE_FP32           # a caller-drawn mask (parity with torch's Philox stream): fp32 kernels
        needs_bwd = any(ctx.needs_input_grad)
        if ctx.needs_input_grad[0] or ctx.needs_input_grad[1]:
            flags |= N.FLAG_FEATURE_GRADS      # fine-tune phase: the backward must return feature gradients
        if needs_bwd:
            flags |= N.FLAG_TRAINING           # the forward keeps what the backward reloads
        desc = _desc(B, d_img, d_txt, n_classes, flags, compute, drop_p, drop_seed)
        L = N.lib()
        ws = _pool.take(L.mmrca_head_workspace_bytes(C.byref(desc), 1 if needs_bwd else 0), img.device)
        ctx.lease = _Lease(ws)
        logits = torch.empty(B, n_classes, dtype=torch.float32, device=img.device)
        hp = _head_struct(params)
        ctx.save_for_backward(img, txt, drop_mask, ws, *params)
        ctx.cfg = (drop_scale, flags, n_classes, compute, drop_p, drop_seed)
        if B == 0:      # empty shard: nothing to launch (a zero-size tensor has a null data pointer)
            return logits
        with torch.cuda.device(img.device):
            N.check(L.mmrca_head_forward(C.byref(desc), C.byref(hp), img.data_ptr(), txt.data_ptr(),
                                         drop_mask.data_ptr() if drop_mask is not None else None,
                                         float(drop_scale), logits.data_ptr(), ws.data_ptr(), ws.numel(),
                                         _stream_ptr(img.device)), "mmrca_head_forward")
        return logits

    @staticmethod
    def backward(ctx, dlogits):
        img, txt, drop_mask, ws, *params = ctx.saved_tensors
        drop_scale, flags, n_classes, compute, drop_p, drop_seed = ctx.cfg
        dlogits = _check_dev(dlogits, "dlogits")
        B = img.shape[0]
        desc = _desc(B, img.shape[1], txt.shape[1], n_classes, flags, compute, drop_p, drop_seed)
        # gradients accumulate (+=) straight into the module's persistent bucket when one is attached
        # (MM_RCA.attach_flat_grads: p.grad IS a view of it), else into a fresh bucket returned to autograd
        sink = ctx.grad_sink
        fg = sink if sink is not None else FlatGrads(params)
        want_feat = ctx.needs_input_grad[0] or ctx.needs_input_grad[1]
        d_img = torch.empty_like(img) if want_feat else None
        d_txt = torch.empty_like(txt) if want_feat else None
        hp, hg = _head_struct(params), _head_struct(fg.views)
        L = N.lib()
        with torch.cuda.device(img.device):
            if B > 0:
                N.check(L.mmrca_head_backward(C.byref(desc), C.byref(hp), img.data_ptr(), txt.data_ptr(),
                    drop_mask.data_ptr() if drop_mask is not None else None, float(drop_scale),
                    dlogits.data_ptr(), C.byref(hg), d_img.data_ptr() if want_feat else None,
                    d_txt.data_ptr() if want_feat else None, ws.data_ptr(), ws.numel(),
                    _stream_ptr(img.device)), "mmrca_head_backward")
        features_only = bool(flags & N.FLAG_FEATURES_ONLY)
        pg = []
        for i, v in enumerate(fg.views):
            need = ctx.needs_input_grad[10 + i] and sink is None
            # features_only: the attention blocks are outside the graph in the reference (their result is
            # discarded, multimodal_model.py:676-699) -> grad None, like autograd there.
            pg.append(v if need and not (features_only and i < 32) else None)
        return (d_img if ctx.needs_input_grad[0] else None, d_txt if ctx.needs_input_grad[1] else None,
                None, None, None, None, None, None, None, None, *pg)


def mmrca_head(img_feat: torch.Tensor, txt_feat: torch.Tensor, params: Sequence[torch.Tensor], *,
               reverse: bool, features_only: bool = False, cross_attention_only: bool = False,
               n_classes: int = 4, drop_mask: Optional[torch.Tensor] = None, drop_scale: float = 1.0,
               drop_p: float = 0.0, drop_seed: int = 0, compute: int = N.COMPUTE_FP32,
               grad_sink: Optional["FlatGrads"] = None) -> torch.Tensor:
    """Fusion head of MM_RCA.forward (reference multimodal_model.py:661-728) on pooled features.

    grad_sink: a persistent FlatGrads over `params` (same order): the backward accumulates into it and returns no
    parameter gradients to autograd (the caller has pointed every p.grad at its view, training.attach_flat_grads).

    params: the 34 tensors in head_param_names(features_only, cross_attention_only) order.
    self.drop (:719), two ways:
      drop_p > 0 (+ drop_seed): the kernels draw the keep mask themselves (dropout_mask() returns the same mask);
      drop_mask: a caller-drawn uint8 keep-mask [B, D], kept values scaled by drop_scale (fp32 kernels).
    """
    flags = make_flags(reverse, features_only, cross_attention_only)
    if drop_mask is not None:
        drop_p, drop_seed = 0.0, 0
    return _HeadFunction.apply(img_feat, txt_feat, drop_mask, drop_scale, flags, n_classes, compute,
                               float(drop_p), int(drop_seed), grad_sink, *params)


def cross_entropy(logits: torch.Tensor, labels: torch.Tensor, class_weight: Optional[torch.Tensor] = None,
                  label_smoothing: float = 0.0, want_dlogits: bool = True
                  ) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    """torch.nn.CrossEntropyLoss(weight, label_smoothing) forward + dL/dlogits (reference
    main_both.py:87-93,110) in one kernel.  Returns (loss[1], dlogits or None)."""
    logits = _check_dev(logits, "logits")
    labels = _check_dev(labels, "labels", torch.int64)
    cw = _check_dev(class_weight, "class weights") if class_weight is not None else None
    B, nc = logits.shape
    loss = torch.empty(1, dtype=torch.float32, device=logits.device)
    dl = torch.empty_like(logits) if want_dlogits else None
    ce = N.CeDesc(cw.data_ptr() if cw is not None else None, float(label_smoothing))
    with torch.cuda.device(logits.device):
        N.check(N.lib().mmrca_cross_entropy(logits.data_ptr(), labels.data_ptr(), C.byref(ce), B, nc,
                                            loss.data_ptr(), dl.data_ptr() if dl is not None else None,
                                            _stream_ptr(logits.device)), "mmrca_cross_entropy")
    return loss, dl


def feature_handoff(hidden: torch.Tensor, fmap: torch.Tensor, out_dtype: torch.dtype = torch.bfloat16
                    ) -> Tuple[torch.Tensor, torch.Tensor]:
    """(image features [B, C], text features [B, H]) from the stock backbones' raw outputs in ONE kernel: the CLS row of
    the text backbone's last hidden state `hidden` [B, T, H] (reference multimodal_model.py:651-658) and the global average
    pool + flatten of the image backbone's final feature map `fmap` [B, C, h, w] (:25-36), cast to `out_dtype` (bf16: what
    the head takes with MMRCA_FLAG_FEATURES_BF16).  fp32 or bf16 inputs, NCHW or channels_last maps."""
    for t, what in ((hidden, "hidden state"), (fmap, "feature map")):
        if not t.is_cuda:
            raise RuntimeError(f"{what} must be a CUDA tensor: the hand-off has no CPU fallback")
        if t.dtype not in (torch.float32, torch.bfloat16):
            raise TypeError(f"{what} must be fp32 or bf16, got {t.dtype}")
    if out_dtype not in (torch.float32, torch.bfloat16):
        raise TypeError("out_dtype must be fp32 or bf16")
    B, T, H = hidden.shape
    if hidden.stride(2) != 1 or hidden.stride(1) != H:
        hidden = hidden.contiguous()
    Bf, Cc, h, w = fmap.shape
    if Bf != B:
        raise ValueError("hidden state and feature map batches differ")
    if fmap.is_contiguous():
        channels_last = 0
    elif fmap.is_contiguous(memory_format=torch.channels_last):
        channels_last = 1
    else:
        fmap, channels_last = fmap.contiguous(), 0
    txt = torch.empty(B, H, dtype=out_dtype, device=hidden.device)
    img = torch.empty(B, Cc, dtype=out_dtype, device=hidden.device)
    if B > 0:
        with torch.cuda.device(hidden.device):
            N.check(N.lib().mmrca_feature_handoff(
                hidden.data_ptr(), int(hidden.dtype == torch.bfloat16), hidden.stride(0), H, fmap.data_ptr(),
                int(fmap.dtype == torch.bfloat16), Cc, h * w, channels_last, B, txt.data_ptr(), img.data_ptr(),
                int(out_dtype == torch.bfloat16), _stream_ptr(hidden.device)), "mmrca_feature_handoff")
    return img, txt


class HeadTrainStep:
    """Persistent buffers for the one-call training step of the head (forward + CrossEntropyLoss +
    backward = the body of run_one_epoch, reference main_both.py:106-112, restricted to the head).

    Gradients are ACCUMULATED into `grads.flat` (like loss.backward()); call zero_grad() between
    optimizer steps.  `grads.flat` is the bucket a data-parallel job all-reduces once per step."""

    def __init__(self, params: Sequence[torch.Tensor], batch: int, d_img: int, d_txt: int, *, reverse: bool,
                 features_only: bool = False, cross_attention_only: bool = False, n_classes: int = 4,
                 class_weight: Optional[torch.Tensor] = None, label_smoothing: float = 0.0,
                 compute: int = N.COMPUTE_FP32, feature_grads: bool = False, drop_p: float = 0.0):
        self.params = [_check_dev(p.detach(), "head parameter") for p in params]
        dev = self.params[0].device
        self.flags = make_flags(reverse, features_only, cross_attention_only) | (N.FLAG_FEATURE_GRADS if feature_grads else 0)
        self.desc = _desc(batch, d_img, d_txt, n_classes, self.flags, compute, drop_p, 0)
        self.concat_width = concat_width(d_img, d_txt, features_only, cross_attention_only)
        self.grads = FlatGrads(self.params)
        self.hp, self.hg = _head_struct(self.params), _head_struct(self.grads.views)
        L = N.lib()
        self.ws = torch.empty(max(1, L.mmrca_head_workspace_bytes(C.byref(self.desc), 1)), dtype=torch.uint8,
                              device=dev)
        self.logits = torch.empty(batch, n_classes, dtype=torch.float32, device=dev)
        self.loss = torch.empty(1, dtype=torch.float32, device=dev)
        self.cw = _check_dev(class_weight, "class weights") if class_weight is not None else None
        self.ce = N.CeDesc(self.cw.data_ptr() if self.cw is not None else None, float(label_smoothing))
        self.d_img = torch.empty(batch, d_img, dtype=torch.float32, device=dev) if feature_grads else None
        self.d_txt = torch.empty(batch, d_txt, dtype=torch.float32, device=dev) if feature_grads else None
        self.device = dev

    def zero_grad(self):
        self.grads.zero_()

    def __call__(self, img: torch.Tensor, txt: torch.Tensor, labels: torch.Tensor,
                 drop_mask: Optional[torch.Tensor] = None, drop_scale: float = 1.0, drop_seed: int = 0,
                 zero_grad: bool = False):
        """drop_seed: seed of this step's dropout mask when the step was built with drop_p > 0.
        zero_grad: clear the gradient bucket inside the step's first kernel (== zero_grad() right before the call,
        without the extra fill launch)."""
        self.desc.drop_seed = int(drop_seed) & (2 ** 64 - 1)
        # bf16 features (a backbone under bf16 autocast, or a host hand-off that ships half the bytes): bf16 pipeline only
        fdt = torch.bfloat16 if (img.dtype == torch.bfloat16 and txt.dtype == torch.bfloat16) else torch.float32
        if fdt == torch.bfloat16 and self.desc.compute == N.COMPUTE_FP32:
            raise TypeError("bf16 features need a step built with compute=COMPUTE_BF16")
        self.desc.flags = ((self.flags | N.FLAG_FEATURES_BF16) if fdt == torch.bfloat16 else self.flags) | \
            (N.FLAG_ZERO_GRADS if zero_grad else 0)
        img, txt = _check_dev(img, "image features", fdt), _check_dev(txt, "text features", fdt)
        labels = _check_dev(labels, "labels", torch.int64)
        if img.shape != (self.desc.batch, self.desc.d_img) or txt.shape != (self.desc.batch, self.desc.d_txt):
            raise ValueError("feature shapes do not match the shapes this step was built for")
        drop_mask = _check_mask(drop_mask, self.desc.batch, self.concat_width, "head")
        if drop_mask is not None and self.desc.compute != N.COMPUTE_FP32:
            raise ValueError("a caller-drawn dropout mask needs a step built with compute=COMPUTE_FP32 (the bf16 pipeline "
                             "draws its seeded mask on chip: drop_p / drop_seed)")
        with torch.cuda.device(self.device):
            N.check(N.lib().mmrca_head_train_step(
                C.byref(self.desc), C.byref(self.hp), img.data_ptr(), txt.data_ptr(),
                drop_mask.data_ptr() if drop_mask is not None else None, float(drop_scale), labels.data_ptr(),
                C.byref(self.ce), self.logits.data_ptr(), self.loss.data_ptr(), C.byref(self.hg),
                self.d_img.data_ptr() if self.d_img is not None else None,
                self.d_txt.data_ptr() if self.d_txt is not None else None,
                self.ws.data_ptr(), self.ws.numel(), _stream_ptr(self.device)), "mmrca_head_train_step")
        return self.loss, self.logits


def _attn_struct(tensors: Sequence[torch.Tensor]) -> N.AttnParams:
    ap = N.AttnParams()
    for f, t in zip(_ATTN_FIELDS, tensors):
        setattr(ap, f, t.data_ptr())
    return ap


class _AttentionFunction(torch.autograd.Function):
    """One SelfAttention (x_kv is x_q) / ReverseCrossAttention block, reference multimodal_model.py:39-108."""

    @staticmethod
    def forward(ctx, x_q, x_kv, reverse, is_self, compute, *params):
        x_q = _check_dev(x_q, "x_q")
        x_kv = x_q if is_self else _check_dev(x_kv, "x_kv")
        params = [_check_dev(p, "attention parameter") for p in params]
        B, Lr, d_in = x_q.shape
        if Lr != NUM_PATCHES or x_kv.shape != x_q.shape:
            raise ValueError("attention block expects [B, 16, d_in] inputs of equal shape (square attention, "
                             "reference multimodal_model.py:93)")
        d_kq, d_v = params[0].shape[0], params[4].shape[0]
        out = torch.empty(B, Lr, d_v, dtype=torch.float32, device=x_q.device)
        ap = _attn_struct(params)
        L = N.lib()
        sb = int(L.mmrca_attention_forward_scratch_bytes(d_in, d_kq, d_v, compute))
        scratch = torch.empty(sb, dtype=torch.uint8, device=x_q.device) if sb else None
        with torch.cuda.device(x_q.device):
            N.check(L.mmrca_attention_forward(C.byref(ap), x_q.data_ptr(), x_kv.data_ptr(), B, d_in, d_kq,
                                              d_v, 1 if reverse else 0, 0, None, out.data_ptr(),
                                              scratch.data_ptr() if sb else None, sb, compute,
                                              _stream_ptr(x_q.device)), "mmrca_attention_forward")
        ctx.save_for_backward(x_q, x_kv, *params)
        ctx.cfg = (reverse, is_self, compute, d_in, d_kq, d_v)
        return out

    @staticmethod
    def backward(ctx, d_out):
        x_q, x_kv, *params = ctx.saved_tensors
        reverse, is_self, compute, d_in, d_kq, d_v = ctx.cfg
        d_out = _check_dev(d_out, "d_out")
        B = x_q.shape[0]
        L = N.lib()
        fg = FlatGrads(params)
        scratch = torch.empty(max(1, L.mmrca_attention_backward_scratch_bytes(B, d_kq, d_v)), dtype=torch.uint8,
                              device=x_q.device)
        dxq = torch.empty_like(x_q)
        dxkv = None if is_self else torch.empty_like(x_kv)
        ap, ag = _attn_struct(params), _attn_struct(fg.views)
        with torch.cuda.device(x_q.device):
            N.check(L.mmrca_attention_backward(C.byref(ap), x_q.data_ptr(), x_kv.data_ptr(), d_out.data_ptr(), B,
                                               d_in, d_kq, d_v, 1 if reverse else 0, C.byref(ag), dxq.data_ptr(),
                                               dxkv.data_ptr() if dxkv is not None else None, scratch.data_ptr(),
                                               scratch.numel(), compute, _stream_ptr(x_q.device)),
                    "mmrca_attention_backward")
        return (dxq, dxkv, None, None, None, *fg.views)


def attention_block(x_q: torch.Tensor, x_kv: Optional[torch.Tensor], params: Sequence[torch.Tensor], *,
                    reverse: bool = False, compute: int = N.COMPUTE_FP32) -> torch.Tensor:
    """params: (Wq, bq, Wk, bk, Wv, bv, ln_weight, ln_bias).  x_kv=None -> self attention."""
    is_self = x_kv is None or x_kv is x_q
    return _AttentionFunction.apply(x_q, x_q if is_self else x_kv, reverse, is_self, compute, *params)


# ---------------------------------------------------------------------------------------------------------
# Hierarchical late-fusion head (reference multimodal_model.py:729-818)
# ---------------------------------------------------------------------------------------------------------
HIER_PARAM_NAMES = ("final_hierarchical_image.weight", "final_hierarchical_image.bias",
                    "final_hierarchical_text.weight", "final_hierarchical_text.bias",
                    "final_hierarchical_all.weight", "final_hierarchical_all.bias")
HIER_SEGMENTS = (1280, 2560, 2048, 768, 768, 768)      # pooled, stage 3, stage 6 | CLS last, layer 2, layer 4
HIER_CONCAT = sum(HIER_SEGMENTS)                        # 8192: [image 5888 | text 2304], the dropout mask's columns


def _hier_feats(feats: Sequence[torch.Tensor]):
    if len(feats) != 6:
        raise ValueError("the hierarchical head takes six feature tensors (3 image, 3 text)")
    feats = [_check_dev(f.float(), "hierarchical feature") for f in feats]
    B = feats[0].shape[0]
    for f, w in zip(feats, HIER_SEGMENTS):
        if f.shape != (B, w):
            raise ValueError(f"hierarchical feature of shape {tuple(f.shape)}, expected ({B}, {w})")
    arr = (N._fp * 6)(*[f.data_ptr() for f in feats])
    return feats, arr, B


def _hier_struct(tensors: Sequence[Optional[torch.Tensor]]) -> N.HierParams:
    return N.HierParams(*[t.data_ptr() if t is not None else None for t in tensors])


class _HierFunction(torch.autograd.Function):
    """logits = hierarchical head(six pooled feature tensors); frozen backbones (no feature gradients)."""

    @staticmethod
    def forward(ctx, drop_mask, drop_scale, drop_p, drop_seed, f0, f1, f2, f3, f4, f5, *params):
        feats, arr, B = _hier_feats((f0, f1, f2, f3, f4, f5))
        params = [_check_dev(p, "hierarchical parameter") for p in params]
        if drop_mask is not None:
            drop_mask = _check_dev(drop_mask, "dropout mask", torch.uint8)
            if drop_mask.shape != (B, HIER_CONCAT):
                raise ValueError("hierarchical dropout mask must be [B, 8192]")
        dev = feats[0].device
        want_feat = any(ctx.needs_input_grad[4:10])
        desc = N.HierDesc(B, params[4].shape[0], float(drop_p), N.HIER_FEATURE_GRADS if want_feat else 0,
                          int(drop_seed) & (2 ** 64 - 1))
        L = N.lib()
        ws = torch.empty(max(1, L.mmrca_hier_workspace_bytes(C.byref(desc))), dtype=torch.uint8, device=dev)
        logits = torch.empty(B, params[4].shape[0], dtype=torch.float32, device=dev)
        ctx.save_for_backward(ws, *params)
        ctx.desc = desc
        ctx.feat_state = (feats, drop_mask, float(drop_scale)) if want_feat else None
        if B > 0:
            with torch.cuda.device(dev):
                N.check(L.mmrca_hier_forward(C.byref(desc), C.byref(_hier_struct(params)), arr,
                                             drop_mask.data_ptr() if drop_mask is not None else None,
                                             float(drop_scale), logits.data_ptr(), ws.data_ptr(), ws.numel(),
                                             _stream_ptr(dev)), "mmrca_hier_forward")
        return logits

    @staticmethod
    def backward(ctx, dlogits):
        ws, *params = ctx.saved_tensors
        dlogits = _check_dev(dlogits, "dlogits")
        fg = FlatGrads(params)
        dfe = [None] * 6
        if ctx.desc.batch > 0:
            with torch.cuda.device(ws.device):
                N.check(N.lib().mmrca_hier_backward(C.byref(ctx.desc), C.byref(_hier_struct(params)), dlogits.data_ptr(),
                                                    C.byref(_hier_struct(fg.views)), ws.data_ptr(), ws.numel(),
                                                    _stream_ptr(ws.device)), "mmrca_hier_backward")
                if ctx.feat_state is not None:      # fine-tune phase (reference main_both.py:687-694)
                    feats, drop_mask, drop_scale = ctx.feat_state
                    dfe = [torch.empty_like(f) for f in feats]
                    N.check(N.lib().mmrca_hier_backward_features(
                        C.byref(ctx.desc), C.byref(_hier_struct(params)), (N._fp * 6)(*[f.data_ptr() for f in feats]),
                        drop_mask.data_ptr() if drop_mask is not None else None, drop_scale,
                        (N._fp * 6)(*[t.data_ptr() for t in dfe]), ws.data_ptr(), ws.numel(), _stream_ptr(ws.device)),
                        "mmrca_hier_backward_features")
        return (None,) * 4 + tuple(g if ctx.needs_input_grad[4 + i] else None for i, g in enumerate(dfe)) + \
            tuple(v if ctx.needs_input_grad[10 + i] else None for i, v in enumerate(fg.views))


def hierarchical_head(feats: Sequence[torch.Tensor], params: Sequence[torch.Tensor], *,
                      drop_mask: Optional[torch.Tensor] = None, drop_scale: float = 1.0, drop_p: float = 0.0,
                      drop_seed: int = 0) -> torch.Tensor:
    """Hierarchical.forward after the backbones and the two AvgPool2d (reference multimodal_model.py:777-816).

    feats: (pooled image [B,1280], stage-3 pooled+flattened [B,2560], stage-6 pooled+flattened [B,2048],
            text CLS last layer, hidden_states[2], hidden_states[4] — [B,768] each);
    params: the six tensors of HIER_PARAM_NAMES.  Dropout as in mmrca_head (seeded, or a caller mask [B, 8192])."""
    if drop_mask is not None:
        drop_p, drop_seed = 0.0, 0
    return _HierFunction.apply(drop_mask, float(drop_scale), float(drop_p), int(drop_seed), *feats, *params)


class HierTrainStep:
    """One-call training step of the hierarchical head (forward + CrossEntropyLoss + backward, reference
    main_both.py:106-112 restricted to the head); gradients accumulate into `grads.flat`."""

    def __init__(self, params: Sequence[torch.Tensor], batch: int, *, class_weight: Optional[torch.Tensor] = None,
                 label_smoothing: float = 0.0, drop_p: float = 0.0):
        self.params = [_check_dev(p.detach(), "hierarchical parameter") for p in params]
        dev = self.params[0].device
        self.desc = N.HierDesc(batch, self.params[4].shape[0], float(drop_p), 0, 0)
        self.grads = FlatGrads(self.params)
        self.hp, self.hg = _hier_struct(self.params), _hier_struct(self.grads.views)
        self.ws = torch.empty(max(1, N.lib().mmrca_hier_workspace_bytes(C.byref(self.desc))), dtype=torch.uint8, device=dev)
        self.logits = torch.empty(batch, self.params[4].shape[0], dtype=torch.float32, device=dev)
        self.loss = torch.empty(1, dtype=torch.float32, device=dev)
        self.cw = _check_dev(class_weight, "class weights") if class_weight is not None else None
        self.ce = N.CeDesc(self.cw.data_ptr() if self.cw is not None else None, float(label_smoothing))
        self.device = dev

    def zero_grad(self):
        self.grads.zero_()

    def __call__(self, feats: Sequence[torch.Tensor], labels: torch.Tensor, drop_mask: Optional[torch.Tensor] = None,
                 drop_scale: float = 1.0, drop_seed: int = 0):
        self.desc.drop_seed = int(drop_seed) & (2 ** 64 - 1)
        feats, arr, B = _hier_feats(feats)
        if B != self.desc.batch:
            raise ValueError("feature batch does not match the batch this step was built for")
        labels = _check_dev(labels, "labels", torch.int64)
        drop_mask = _check_mask(drop_mask, B, HIER_CONCAT, "hierarchical")
        with torch.cuda.device(self.device):
            N.check(N.lib().mmrca_hier_train_step(
                C.byref(self.desc), C.byref(self.hp), arr, drop_mask.data_ptr() if drop_mask is not None else None,
                float(drop_scale), labels.data_ptr(), C.byref(self.ce), self.logits.data_ptr(), self.loss.data_ptr(),
                C.byref(self.hg), self.ws.data_ptr(), self.ws.numel(), _stream_ptr(self.device)), "mmrca_hier_train_step")
        return self.loss, self.logits


# ---------------------------------------------------------------------------------------------------------
# Classic / Normalized late-fusion heads (reference multimodal_model.py:489-579)
# ---------------------------------------------------------------------------------------------------------
FUSION_PARAM_NAMES = ("image_to_hidden_size.weight", "image_to_hidden_size.bias", "text_to_hidden_size.weight",
                      "text_to_hidden_size.bias", "concat_layer.weight", "concat_layer.bias", "fc_layer.weight",
                      "fc_layer.bias")


def _fusion_struct(tensors: Sequence[Optional[torch.Tensor]]) -> N.FusionParams:
    return N.FusionParams(*[t.data_ptr() if t is not None else None for t in tensors])


def _fusion_desc(B, params, normalized, drop_p, drop_seed, compute: str = "fp32") -> N.FusionDesc:
    H, d_img = params[0].shape
    d_txt = params[2].shape[1]
    if tuple(params[4].shape) != (H, 2 * H) or params[6].shape[1] != H:
        raise ValueError("fusion head: concat_layer must be [H, 2H] and fc_layer [n_classes, H]")
    if compute not in ("fp32", "bf16"):
        raise ValueError("compute must be 'fp32' or 'bf16'")
    flags = (N.FUSION_NORMALIZED if normalized else 0) | (N.FUSION_BF16 if compute == "bf16" else 0)
    return N.FusionDesc(B, d_img, d_txt, H, params[6].shape[0], flags, float(drop_p), int(drop_seed) & (2 ** 64 - 1))


class _FusionFunction(torch.autograd.Function):
    """logits = classic / normalized fusion head(pooled image features, text CLS features)."""

    @staticmethod
    def forward(ctx, img, txt, drop_mask, drop_scale, normalized, drop_p, drop_seed, compute, *params):
        img, txt = _check_dev(img, "image features"), _check_dev(txt, "text features")
        params = [_check_dev(p, "fusion parameter") for p in params]
        B = img.shape[0]
        desc = _fusion_desc(B, params, normalized, drop_p, drop_seed, compute)
        if img.shape[1] != desc.d_img or txt.shape != (B, desc.d_txt):
            raise ValueError("feature shapes do not match the projection weights")
        drop_mask = _check_mask(drop_mask, B, desc.hidden, "fusion head")
        L = N.lib()
        nbytes = L.mmrca_fusion_workspace_bytes(C.byref(desc))
        if nbytes == 0:
            raise ValueError("unsupported fusion head shape: " + N.last_error())
        ws = _pool.take(nbytes, img.device)
        ctx.lease = _Lease(ws)
        logits = torch.empty(B, desc.n_classes, dtype=torch.float32, device=img.device)
        ctx.save_for_backward(img, txt, drop_mask, ws, *params)
        ctx.cfg = (desc, float(drop_scale))
        if B > 0:
            with torch.cuda.device(img.device):
                N.check(L.mmrca_fusion_forward(C.byref(desc), C.byref(_fusion_struct(params)), img.data_ptr(), txt.data_ptr(),
                                               drop_mask.data_ptr() if drop_mask is not None else None, float(drop_scale),
                                               logits.data_ptr(), ws.data_ptr(), ws.numel(), _stream_ptr(img.device)),
                        "mmrca_fusion_forward")
        return logits

    @staticmethod
    def backward(ctx, dlogits):
        img, txt, drop_mask, ws, *params = ctx.saved_tensors
        desc, drop_scale = ctx.cfg
        dlogits = _check_dev(dlogits, "dlogits")
        fg = FlatGrads(params)
        want_feat = ctx.needs_input_grad[0] or ctx.needs_input_grad[1]
        d_img = torch.empty_like(img) if want_feat else None
        d_txt = torch.empty_like(txt) if want_feat else None
        if desc.batch > 0:
            with torch.cuda.device(img.device):
                N.check(N.lib().mmrca_fusion_backward(
                    C.byref(desc), C.byref(_fusion_struct(params)), img.data_ptr(), txt.data_ptr(),
                    drop_mask.data_ptr() if drop_mask is not None else None, drop_scale, dlogits.data_ptr(),
                    C.byref(_fusion_struct(fg.views)), d_img.data_ptr() if want_feat else None,
                    d_txt.data_ptr() if want_feat else None, ws.data_ptr(), ws.numel(), _stream_ptr(img.device)),
                    "mmrca_fusion_backward")
        elif want_feat:
            d_img.zero_(); d_txt.zero_()
        return (d_img if ctx.needs_input_grad[0] else None, d_txt if ctx.needs_input_grad[1] else None, None, None, None,
                None, None, None, *[v if ctx.needs_input_grad[8 + i] else None for i, v in enumerate(fg.views)])


def fusion_head(img_feat: torch.Tensor, txt_feat: torch.Tensor, params: Sequence[torch.Tensor], *, normalized: bool,
                drop_mask: Optional[torch.Tensor] = None, drop_scale: float = 1.0, drop_p: float = 0.0,
                drop_seed: int = 0, compute: str = "fp32") -> torch.Tensor:
    """Classic (`normalized=False`) / Normalized late-fusion head after the backbones (reference
    multimodal_model.py:521-529 / :566-577).  params: the eight tensors of FUSION_PARAM_NAMES.  Dropout on the
    concat_layer output [B, H]: seeded (drop_p, drop_seed; dropout_mask(seed, p, B, H) returns the mask) or caller-drawn.
    compute="bf16": the three Linear layers' GEMMs and their weight / hidden gradients on the tensor cores (bf16 operands,
    fp32 accumulate; 2e-2-absolute logits contract); "fp32": the 1e-4-relative contract."""
    if drop_mask is not None:
        drop_p, drop_seed = 0.0, 0
    return _FusionFunction.apply(img_feat, txt_feat, drop_mask, float(drop_scale), bool(normalized), float(drop_p),
                                 int(drop_seed), str(compute), *params)


class FusionTrainStep:
    """One-call training step of the classic / normalized head (forward + CrossEntropyLoss + backward, reference
    main_both.py:106-112 restricted to the head); gradients accumulate into `grads.flat`."""

    def __init__(self, params: Sequence[torch.Tensor], batch: int, *, normalized: bool,
                 class_weight: Optional[torch.Tensor] = None, label_smoothing: float = 0.0, drop_p: float = 0.0,
                 feature_grads: bool = False, compute: str = "fp32"):
        """compute="bf16": the three Linear layers' GEMMs (projections, concat_layer) and their weight / hidden gradients
        run as TMA-fed bf16 tcgen05 GEMMs (fp32 accumulate; the 2e-2-absolute logits contract instead of 1e-4 relative);
        normalisation, dropout, fc_layer and the loss stay fp32."""
        self.params = [_check_dev(p.detach(), "fusion parameter") for p in params]
        dev = self.params[0].device
        self.desc = _fusion_desc(batch, self.params, normalized, drop_p, 0, compute)
        if N.lib().mmrca_fusion_workspace_bytes(C.byref(self.desc)) == 0:
            raise ValueError("unsupported fusion head shape: " + N.last_error())
        self.grads = FlatGrads(self.params)
        self.hp, self.hg = _fusion_struct(self.params), _fusion_struct(self.grads.views)
        self.ws = torch.empty(max(1, N.lib().mmrca_fusion_workspace_bytes(C.byref(self.desc))), dtype=torch.uint8, device=dev)
        self.logits = torch.empty(batch, self.desc.n_classes, dtype=torch.float32, device=dev)
        self.loss = torch.empty(1, dtype=torch.float32, device=dev)
        self.cw = _check_dev(class_weight, "class weights") if class_weight is not None else None
        self.ce = N.CeDesc(self.cw.data_ptr() if self.cw is not None else None, float(label_smoothing))
        self.d_img = torch.empty(batch, self.desc.d_img, dtype=torch.float32, device=dev) if feature_grads else None
        self.d_txt = torch.empty(batch, self.desc.d_txt, dtype=torch.float32, device=dev) if feature_grads else None
        self.device = dev

    def zero_grad(self):
        self.grads.zero_()

    def __call__(self, img: torch.Tensor, txt: torch.Tensor, labels: torch.Tensor, drop_mask: Optional[torch.Tensor] = None,
                 drop_scale: float = 1.0, drop_seed: int = 0):
        self.desc.drop_seed = int(drop_seed) & (2 ** 64 - 1)
        img, txt = _check_dev(img, "image features"), _check_dev(txt, "text features")
        labels = _check_dev(labels, "labels", torch.int64)
        if img.shape != (self.desc.batch, self.desc.d_img) or txt.shape != (self.desc.batch, self.desc.d_txt):
            raise ValueError("feature shapes do not match the shapes this step was built for")
        drop_mask = _check_mask(drop_mask, self.desc.batch, self.desc.hidden, "fusion head")
        with torch.cuda.device(self.device):
            N.check(N.lib().mmrca_fusion_train_step(
                C.byref(self.desc), C.byref(self.hp), img.data_ptr(), txt.data_ptr(),
                drop_mask.data_ptr() if drop_mask is not None else None, float(drop_scale), labels.data_ptr(),
                C.byref(self.ce), self.logits.data_ptr(), self.loss.data_ptr(), C.byref(self.hg),
                self.d_img.data_ptr() if self.d_img is not None else None,
                self.d_txt.data_ptr() if self.d_txt is not None else None, self.ws.data_ptr(), self.ws.numel(),
                _stream_ptr(self.device)), "mmrca_fusion_train_step")
        return self.loss, self.logits


# ---------------------------------------------------------------------------------------------------------
# Token-level attention blocks (BASELINE.json configs[4]): SelfAttention / ReverseCrossAttention on [B, L, d_in], L <= 256
# ---------------------------------------------------------------------------------------------------------
class _TokenFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, block, x_q, x_kv, *params):
        ctx.block = block
        ctx.dtypes = (x_q.dtype, x_kv.dtype if x_kv is not None else None)
        out = block(x_q, x_kv)
        ctx.serial = block._serial = getattr(block, "_serial", 0) + 1
        return out.clone()

    @staticmethod
    def backward(ctx, d_out):
        block = ctx.block
        if block._serial != ctx.serial:
            raise RuntimeError("TokenAttention: the workspace was reused by a later forward before this backward ran")
        grads = [torch.zeros_like(p) for p in block.params]
        need_q, need_kv = ctx.needs_input_grad[1], ctx.needs_input_grad[2]
        dxq, dxkv = block.backward(d_out, grads, need_q, need_kv)
        if dxq is not None and not need_q:
            dxq = None
        if dxq is not None and ctx.dtypes[0] != torch.float32:
            dxq = dxq.to(ctx.dtypes[0])
        if dxkv is not None and ctx.dtypes[1] != torch.float32:
            dxkv = dxkv.to(ctx.dtypes[1])
        pg = [g if ctx.needs_input_grad[3 + i] else None for i, g in enumerate(grads)]
        return (None, dxq, dxkv, *pg)


class TokenAttention:
    """SelfAttention.forward (reference multimodal_model.py:51-68) / ReverseCrossAttention.forward (:82-108) on real token
    sequences (ViT-L/16: [B, 197, 1024]; RoBERTa: [B, 256, 768]; the blocks' own outputs [B, L, 96]) through the bf16
    tensor-core path: a TMA-fed tcgen05 projection GEMM + one attention kernel per (sample, 128-query tile).
    Activations are bf16 (fp32 inputs are cast at the hand-off); parameters fp32 (W_query.weight, W_query.bias,
    W_key.weight, W_key.bias, W_value.weight, W_value.bias, norm.weight, norm.bias); output fp32 [B, L, d_v].
    Persistent workspace: build once per (batch, L, widths), call many times.  training=True keeps the attention weights
    for backward() (the five-GEMM attention backward + weight / input gradient GEMMs, all tcgen05), and apply() runs the
    block under torch.autograd."""

    def __init__(self, params: Sequence[torch.Tensor], batch: int, seq_len: int, *, d_in_kv: Optional[int] = None,
                 reverse: bool = False, training: bool = False, out_dtype: torch.dtype = torch.float32):
        if out_dtype not in (torch.float32, torch.bfloat16):
            raise ValueError("out_dtype must be torch.float32 or torch.bfloat16")
        self.param_tensors = list(params)
        self.training = training
        self._saved = None
        self.params = [_check_dev(p.detach(), "attention parameter") for p in params]
        dev = self.params[0].device
        d_kq, d_in_q = self.params[0].shape
        d_v = self.params[4].shape[0]
        kkv = self.params[2].shape[1]
        if d_in_kv is not None and d_in_kv != kkv:
            raise ValueError("d_in_kv does not match W_key")
        self.desc = N.TokenDesc(batch, seq_len, d_in_q, kkv, d_kq, d_v, 1 if reverse else 0,
                                (N.TOKEN_TRAINING if training else 0) | (N.TOKEN_OUT_BF16 if out_dtype == torch.bfloat16 else 0))
        self.ap = _attn_struct(self.params)
        nbytes = N.lib().mmrca_token_attention_workspace_bytes(C.byref(self.desc))
        if nbytes == 0:
            raise ValueError("unsupported token attention shape: " + N.last_error())
        self.ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        self.out = torch.empty(batch, seq_len, d_v, dtype=out_dtype, device=dev)
        self.device = dev

    def __call__(self, x_q: torch.Tensor, x_kv: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """out: optional destination (contiguous, [B, L, d_v], the block's out_dtype) instead of the block's own buffer."""
        d = self.desc
        if out is None:
            out = self.out
        elif (tuple(out.shape) != tuple(self.out.shape) or out.dtype != self.out.dtype or not out.is_contiguous()
              or out.device != self.device):
            raise ValueError("out must be a contiguous tensor shaped and typed like the block's output")
        xs = []
        for x, k, what in ((x_q, d.d_in_q, "x_q"), (x_kv, d.d_in_kv, "x_kv")):
            if x is None:
                xs.append(None)
                continue
            if not x.is_cuda:
                raise RuntimeError(f"{what} must be a CUDA tensor: the token attention has no CPU fallback")
            if tuple(x.shape) != (d.batch, d.seq_len, k):
                raise ValueError(f"{what} has shape {tuple(x.shape)}, expected {(d.batch, d.seq_len, k)} "
                                 "(square attention, reference multimodal_model.py:93)")
            xs.append((x if x.dtype == torch.bfloat16 else x.to(torch.bfloat16)).contiguous())
        with torch.cuda.device(self.device):
            N.check(N.lib().mmrca_token_attention_forward(
                C.byref(self.desc), C.byref(self.ap), xs[0].data_ptr(), xs[1].data_ptr() if xs[1] is not None else None,
                out.data_ptr(), self.ws.data_ptr(), self.ws.numel(), _stream_ptr(self.device)),
                "mmrca_token_attention_forward")
        self.desc.flags |= N.TOKEN_WEIGHTS_READY      # the bf16 weights now sit in the workspace
        self._saved = (xs[0], xs[1])
        return out

    def backward(self, d_out: torch.Tensor, grads: Sequence[torch.Tensor], need_dx_q: bool = False,
                 need_dx_kv: bool = False) -> Tuple[Optional[torch.Tensor], Optional[torch.Tensor]]:
        """Accumulates the parameter gradients into `grads` (eight fp32 tensors shaped like the parameters) and returns
        (d_x_q, d_x_kv) (fp32, None where not asked for; self attention: d_x_q is the whole input gradient).  Must
        follow the forward call it differentiates: it reads the workspace that call left."""
        if not self.training or self._saved is None:
            raise RuntimeError("TokenAttention.backward needs training=True and a preceding forward call")
        d = self.desc
        xq, xkv = self._saved
        self_attn = xkv is None
        d_out = _check_dev(d_out, "d_out").to(torch.float32).contiguous()
        if tuple(d_out.shape) != (d.batch, d.seq_len, d.d_v):
            raise ValueError(f"d_out has shape {tuple(d_out.shape)}, expected {(d.batch, d.seq_len, d.d_v)}")
        for g, p_ in zip(grads, self.params):
            if g.dtype != torch.float32 or g.shape != p_.shape or not g.is_contiguous() or g.device != self.device:
                raise ValueError("gradient buffers must be contiguous fp32 tensors shaped like the parameters")
        ag = _attn_struct(list(grads))
        dxq = torch.empty(d.batch, d.seq_len, d.d_in_q, dtype=torch.float32, device=self.device) \
            if (need_dx_q or (self_attn and need_dx_kv)) else None
        dxkv = torch.empty(d.batch, d.seq_len, d.d_in_kv, dtype=torch.float32, device=self.device) \
            if (need_dx_kv and not self_attn) else None
        with torch.cuda.device(self.device):
            N.check(N.lib().mmrca_token_attention_backward(
                C.byref(self.desc), C.byref(self.ap), xq.data_ptr(), xkv.data_ptr() if xkv is not None else None,
                d_out.data_ptr(), C.byref(ag), dxq.data_ptr() if dxq is not None else None,
                dxkv.data_ptr() if dxkv is not None else None, self.ws.data_ptr(), self.ws.numel(),
                _stream_ptr(self.device)), "mmrca_token_attention_backward")
        return dxq, dxkv

    def apply(self, x_q: torch.Tensor, x_kv: Optional[torch.Tensor] = None) -> torch.Tensor:
        """The block under torch.autograd: gradients flow to the parameter tensors this object was built from and to
        x_q / x_kv.  (One live graph per object: the backward reads the workspace of the latest forward.)"""
        return _TokenFunction.apply(self, x_q, x_kv, *self.param_tensors)

    def refresh_weights(self) -> None:
        """Call after the parameters changed (optimizer step, load_state_dict): the next call converts them again."""
        self.desc.flags &= ~N.TOKEN_WEIGHTS_READY
