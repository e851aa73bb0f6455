"""TEST INFRASTRUCTURE ONLY — CPU oracle for the MM-RCA fusion head.

A functional restatement (torch on CPU, plus an independent float64 numpy
forward/backward with hand-derived gradients) of the hot path of
espiriki/Garbage_Classification_RCA.  Every function cites the reference
file:line it follows (paths relative to the reference checkout root).

Parity pin: the reference ships no tests / golden vectors, so
this oracle is pinned against OUTPUTS OF THE REFERENCE ITSELF: the fixtures in
tests/golden/*.npz were produced by importing the unmodified reference module
(tests/golden/make_golden.py, run in the build container where
/root/reference exists) and tests/test_oracle_vs_golden.py checks this file
against them.

Parameters are passed as a dict keyed by the reference's state_dict names
(e.g. "self_attention_text.W_query.weight"), so the same dict drives the
reference module, this oracle and the CUDA path.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import numpy as np
import torch

NUM_PATCHES = 16          # multimodal_model.py:249
LN_EPS = 1e-5             # torch.nn.LayerNorm default, multimodal_model.py:48,78

SA_TEXT = "self_attention_text"
SA_IMAGE = "self_attention_image"
CA_1 = "cross_attention_1"   # Q from text SA, K/V from image SA (multimodal_model.py:683-684)
CA_2 = "cross_attention_2"   # Q from image SA, K/V from text SA (multimodal_model.py:685-686)
_ATTN_BLOCKS = (SA_IMAGE, SA_TEXT, CA_1, CA_2)
_ATTN_LEAVES = ("W_query.weight", "W_query.bias", "W_key.weight", "W_key.bias",
                "W_value.weight", "W_value.bias", "norm.weight", "norm.bias")


def final_linear_name(features_only: bool, cross_attention_only: bool) -> str:
    """Which classifier the forward uses — multimodal_model.py:721-726."""
    if features_only:
        return "final_features_only_linear"
    if cross_attention_only:
        return "cross_attention_only_linear"
    return "final_with_everything"


def head_param_names(features_only: bool = False, cross_attention_only: bool = False):
    """Names of all head tensors that are *read* by MM_RCA.forward (34 in the full
    variant; the attention blocks run unconditionally, multimodal_model.py:676-692)."""
    names = [f"{blk}.{leaf}" for blk in _ATTN_BLOCKS for leaf in _ATTN_LEAVES]
    fin = final_linear_name(features_only, cross_attention_only)
    names += [f"{fin}.weight", f"{fin}.bias"]
    return names


def init_head_params(d_img: int = 1280, d_txt: int = 768, n_classes: int = 4,
                     features_only: bool = False, cross_attention_only: bool = False,
                     seed: int = 0, dtype=torch.float32, qk_gain: float = 1.0) -> Dict[str, torch.Tensor]:
    """Random head parameters with torch.nn.Linear-like scale; LayerNorm affine is
    perturbed away from (1, 0) so that gamma/beta paths are exercised.  Shapes follow
    multimodal_model.py:249-292.

    qk_gain multiplies the self-attention W_query/W_key weights.  At init scale the
    L2-normalised 48/80-wide chunks give near-zero scores, i.e. uniform attention,
    identical SA rows and analytically ~0 query/key gradients; qk_gain ~ 40 gives
    "trained-like" O(1) scores so that the softmax paths are really exercised."""
    g = torch.Generator().manual_seed(seed)
    p_img, p_txt = d_img // NUM_PATCHES, d_txt // NUM_PATCHES

    def lin(out_f, in_f):
        k = 1.0 / math.sqrt(in_f)
        w = (torch.rand(out_f, in_f, generator=g, dtype=torch.float64) * 2 - 1) * k
        b = (torch.rand(out_f, generator=g, dtype=torch.float64) * 2 - 1) * k
        return w.to(dtype), b.to(dtype)

    p: Dict[str, torch.Tensor] = {}

    def attn(prefix, d_q, d_kv, d_kq, d_v):
        p[f"{prefix}.W_query.weight"], p[f"{prefix}.W_query.bias"] = lin(d_kq, d_q)
        p[f"{prefix}.W_key.weight"], p[f"{prefix}.W_key.bias"] = lin(d_kq, d_kv)
        p[f"{prefix}.W_value.weight"], p[f"{prefix}.W_value.bias"] = lin(d_v, d_kv)
        p[f"{prefix}.norm.weight"] = (1.0 + 0.2 * torch.randn(d_v, generator=g, dtype=torch.float64)).to(dtype)
        p[f"{prefix}.norm.bias"] = (0.2 * torch.randn(d_v, generator=g, dtype=torch.float64)).to(dtype)

    attn(SA_IMAGE, p_img, p_img, 128, 96)     # multimodal_model.py:266-267
    attn(SA_TEXT, p_txt, p_txt, 128, 96)      # :268-269
    attn(CA_1, 96, 96, 64, 48)                # :271-273
    attn(CA_2, 96, 96, 64, 48)                # :275-277
    fin = final_linear_name(features_only, cross_attention_only)
    d_cat = concat_width(d_img, d_txt, features_only, cross_attention_only)
    p[f"{fin}.weight"], p[f"{fin}.bias"] = lin(n_classes, d_cat)
    if qk_gain != 1.0:
        for blk in (SA_IMAGE, SA_TEXT):
            for leaf in ("W_query.weight", "W_key.weight"):
                p[f"{blk}.{leaf}"] = p[f"{blk}.{leaf}"] * qk_gain
    return p


def concat_width(d_img: int, d_txt: int, features_only: bool, cross_attention_only: bool) -> int:
    """multimodal_model.py:282-292."""
    ca = 48 * NUM_PATCHES * 2
    if features_only:
        return d_img + d_txt
    if cross_attention_only:
        return ca
    return ca + d_img + d_txt


# ----------------------------------------------------------------------------
# torch functional restatement (autograd supplies gradients)
# ----------------------------------------------------------------------------

def l2_normalise(x: torch.Tensor) -> torch.Tensor:
    """multimodal_model.py:662-665 — x / ||x||_2 per row, NO epsilon."""
    return x / x.norm(dim=1, keepdim=True)


def self_attention(x: torch.Tensor, p: Dict[str, torch.Tensor], prefix: str) -> torch.Tensor:
    """SelfAttention.forward — multimodal_model.py:51-68."""
    wq, bq = p[f"{prefix}.W_query.weight"], p[f"{prefix}.W_query.bias"]
    wk, bk = p[f"{prefix}.W_key.weight"], p[f"{prefix}.W_key.bias"]
    wv, bv = p[f"{prefix}.W_value.weight"], p[f"{prefix}.W_value.bias"]
    keys = torch.nn.functional.linear(x, wk, bk)            # :52
    queries = torch.nn.functional.linear(x, wq, bq)         # :53
    values = torch.nn.functional.linear(x, wv, bv)          # :54
    scores = queries @ keys.transpose(-1, -2)               # :56
    attn = torch.softmax(scores / wq.shape[0] ** 0.5, dim=-1)   # :58-60
    ctx = attn @ values                                     # :62
    out = torch.nn.functional.layer_norm(ctx, (wv.shape[0],), p[f"{prefix}.norm.weight"],
                                         p[f"{prefix}.norm.bias"], LN_EPS)   # :65
    return torch.relu(out)                                  # :66


def reverse_cross_attention(x1: torch.Tensor, x2: torch.Tensor, p: Dict[str, torch.Tensor],
                            prefix: str, reverse: bool) -> torch.Tensor:
    """ReverseCrossAttention.forward — multimodal_model.py:82-108."""
    wq, bq = p[f"{prefix}.W_query.weight"], p[f"{prefix}.W_query.bias"]
    wk, bk = p[f"{prefix}.W_key.weight"], p[f"{prefix}.W_key.bias"]
    wv, bv = p[f"{prefix}.W_value.weight"], p[f"{prefix}.W_value.bias"]
    q = torch.nn.functional.linear(x1, wq, bq)              # :83
    k = torch.nn.functional.linear(x2, wk, bk)              # :84
    v = torch.nn.functional.linear(x2, wv, bv)              # :85
    scores = q @ k.transpose(-1, -2)                        # :87
    attn = torch.softmax(scores / wq.shape[0] ** 0.5, dim=-1)   # :89-91
    assert attn.shape[1] == attn.shape[2]                   # :93
    if reverse:
        dim = attn.shape[1]
        ctx = ((1.0 - attn) / (dim - 1)) @ v                # :97-99
    else:
        ctx = attn @ v                                      # :102
    out = torch.nn.functional.layer_norm(ctx, (wv.shape[0],), p[f"{prefix}.norm.weight"],
                                         p[f"{prefix}.norm.bias"], LN_EPS)   # :105
    return torch.relu(out)                                  # :106


def head_forward(p: Dict[str, torch.Tensor], img_feat: torch.Tensor, txt_feat: torch.Tensor,
                 reverse: bool = True, features_only: bool = False,
                 cross_attention_only: bool = False,
                 drop_mask: Optional[torch.Tensor] = None, drop_scale: float = 1.0,
                 return_intermediates: bool = False):
    """MM_RCA.forward from the pooled features on — multimodal_model.py:661-728.

    drop_mask (uint8/bool [B, D]) with drop_scale = 1/(1-p) stands in for
    torch.nn.Dropout (:719) so that the mask is shared by oracle and CUDA path;
    None means eval mode / p = 0.
    """
    txt_n = l2_normalise(txt_feat)                          # :662-663
    img_n = l2_normalise(img_feat)                          # :664-665
    bs = txt_n.shape[0]
    txt_r = txt_n.reshape(bs, NUM_PATCHES, -1)              # :669-671
    img_r = img_n.reshape(bs, NUM_PATCHES, -1)              # :672-674
    t_sa = self_attention(txt_r, p, SA_TEXT)                # :677-678
    i_sa = self_attention(img_r, p, SA_IMAGE)               # :679-680
    t_i = reverse_cross_attention(t_sa, i_sa, p, CA_1, reverse)   # :683-684
    i_t = reverse_cross_attention(i_sa, t_sa, p, CA_2, reverse)   # :685-686
    t_i_f = t_i.flatten(1, 2)                               # :689-690
    i_t_f = i_t.flatten(1, 2)                               # :691-692
    if features_only:
        cat = torch.cat((img_n, txt_n), dim=1)              # :694-699
    elif cross_attention_only:
        cat = torch.cat((t_i_f, i_t_f), dim=1)              # :701-706
    else:
        cat = torch.cat((t_i_f, i_t_f, img_n, txt_n), dim=1)   # :708-716
    if drop_mask is not None:
        cat = cat * drop_mask.to(cat.dtype) * drop_scale    # :719
    fin = final_linear_name(features_only, cross_attention_only)
    logits = torch.nn.functional.linear(cat, p[f"{fin}.weight"], p[f"{fin}.bias"])   # :721-726
    if return_intermediates:
        return logits, dict(txt_n=txt_n, img_n=img_n, t_sa=t_sa, i_sa=i_sa, t_i=t_i, i_t=i_t, cat=cat)
    return logits


def cross_entropy(logits: torch.Tensor, labels: torch.Tensor,
                  weight: Optional[torch.Tensor] = None, label_smoothing: float = 0.0) -> torch.Tensor:
    """torch.nn.CrossEntropyLoss(weight=, label_smoothing=) mean reduction, written out —
    main_both.py:87-93,110.  SURVEY.md §8 a8:
        l_i = (1-e) w[y_i] (-log p_i[y_i]) + (e/C) sum_c w[c] (-log p_i[c]);  loss = sum_i l_i / sum_i w[y_i]
    """
    logp = torch.log_softmax(logits, dim=1)
    n, c = logits.shape
    w = torch.ones(c, dtype=logits.dtype, device=logits.device) if weight is None else weight.to(logits.dtype)
    wy = w[labels]
    nll = -(logp[torch.arange(n, device=logits.device), labels]) * wy
    smooth = -(logp * w[None, :]).sum(dim=1)
    li = (1.0 - label_smoothing) * nll + (label_smoothing / c) * smooth
    return li.sum() / wy.sum()


def head_loss_and_grads(p: Dict[str, torch.Tensor], img_feat: torch.Tensor, txt_feat: torch.Tensor,
                        labels: torch.Tensor, reverse: bool = True, features_only: bool = False,
                        cross_attention_only: bool = False, class_weight: Optional[torch.Tensor] = None,
                        label_smoothing: float = 0.0, drop_mask: Optional[torch.Tensor] = None,
                        drop_scale: float = 1.0, feature_grads: bool = False):
    """forward + CrossEntropyLoss + backward (main_both.py:106-112) through autograd.
    Returns (logits, loss, grads{name: tensor}, d_img, d_txt)."""
    pp = {k: v.detach().clone().requires_grad_(True) for k, v in p.items()}
    img = img_feat.detach().clone().requires_grad_(feature_grads)
    txt = txt_feat.detach().clone().requires_grad_(feature_grads)
    logits = head_forward(pp, img, txt, reverse, features_only, cross_attention_only, drop_mask, drop_scale)
    loss = cross_entropy(logits, labels, class_weight, label_smoothing)
    loss.backward()
    grads = {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in pp.items()}
    return logits.detach(), loss.detach(), grads, (img.grad if feature_grads else None), \
        (txt.grad if feature_grads else None)


# ----------------------------------------------------------------------------
# Independent numpy float64 forward/backward with hand-derived gradients.  This is
# the formula sheet the CUDA backward kernels follow; test_oracle_vs_golden.py checks
# it against autograd through the restatement above and against reference goldens.
# ----------------------------------------------------------------------------

def _np(p):
    return {k: np.asarray(v.detach().cpu().double().numpy() if isinstance(v, torch.Tensor) else v,
                          dtype=np.float64) for k, v in p.items()}


def _np_attn_fwd(xq, xkv, p, prefix, reverse):
    wq, bq = p[f"{prefix}.W_query.weight"], p[f"{prefix}.W_query.bias"]
    wk, bk = p[f"{prefix}.W_key.weight"], p[f"{prefix}.W_key.bias"]
    wv, bv = p[f"{prefix}.W_value.weight"], p[f"{prefix}.W_value.bias"]
    g, b = p[f"{prefix}.norm.weight"], p[f"{prefix}.norm.bias"]
    q = xq @ wq.T + bq
    k = xkv @ wk.T + bk
    v = xkv @ wv.T + bv
    scale = 1.0 / math.sqrt(wq.shape[0])
    s = (q @ k.transpose(0, 2, 1)) * scale
    s = s - s.max(axis=-1, keepdims=True)
    e = np.exp(s)
    a = e / e.sum(axis=-1, keepdims=True)
    L = a.shape[1]
    pm = (1.0 - a) / (L - 1) if reverse else a
    ctx = pm @ v
    mu = ctx.mean(axis=-1, keepdims=True)
    var = ((ctx - mu) ** 2).mean(axis=-1, keepdims=True)
    rstd = 1.0 / np.sqrt(var + LN_EPS)
    xhat = (ctx - mu) * rstd
    y = xhat * g + b
    out = np.maximum(y, 0.0)
    cache = dict(xq=xq, xkv=xkv, q=q, k=k, v=v, a=a, pm=pm, xhat=xhat, rstd=rstd, y=y, scale=scale,
                 reverse=reverse, L=L)
    return out, cache


def _np_attn_bwd(dout, c, p, prefix, grads):
    wq = p[f"{prefix}.W_query.weight"]
    wk = p[f"{prefix}.W_key.weight"]
    wv = p[f"{prefix}.W_value.weight"]
    g = p[f"{prefix}.norm.weight"]
    dy = dout * (c["y"] > 0)
    grads[f"{prefix}.norm.weight"] = (dy * c["xhat"]).sum(axis=(0, 1))
    grads[f"{prefix}.norm.bias"] = dy.sum(axis=(0, 1))
    dxhat = dy * g
    dctx = c["rstd"] * (dxhat - dxhat.mean(axis=-1, keepdims=True)
                        - c["xhat"] * (dxhat * c["xhat"]).mean(axis=-1, keepdims=True))
    dpm = dctx @ c["v"].transpose(0, 2, 1)
    dv = c["pm"].transpose(0, 2, 1) @ dctx
    da = -dpm / (c["L"] - 1) if c["reverse"] else dpm
    a = c["a"]
    ds = a * (da - (da * a).sum(axis=-1, keepdims=True)) * c["scale"]
    dq = ds @ c["k"]
    dk = ds.transpose(0, 2, 1) @ c["q"]
    grads[f"{prefix}.W_query.weight"] = np.einsum("brn,brk->nk", dq, c["xq"])
    grads[f"{prefix}.W_query.bias"] = dq.sum(axis=(0, 1))
    grads[f"{prefix}.W_key.weight"] = np.einsum("brn,brk->nk", dk, c["xkv"])
    grads[f"{prefix}.W_key.bias"] = dk.sum(axis=(0, 1))
    grads[f"{prefix}.W_value.weight"] = np.einsum("brn,brk->nk", dv, c["xkv"])
    grads[f"{prefix}.W_value.bias"] = dv.sum(axis=(0, 1))
    dxq = dq @ wq
    dxkv = dk @ wk + dv @ wv
    return dxq, dxkv


def np_cross_entropy_fwd_bwd(logits, labels, weight=None, label_smoothing=0.0):
    """float64 loss and dloss/dlogits for cross_entropy() above."""
    z = np.asarray(logits, dtype=np.float64)
    n, c = z.shape
    w = np.ones(c) if weight is None else np.asarray(weight, dtype=np.float64)
    zs = z - z.max(axis=1, keepdims=True)
    logp = zs - np.log(np.exp(zs).sum(axis=1, keepdims=True))
    prob = np.exp(logp)
    y = np.asarray(labels).astype(np.int64)
    t = np.tile((label_smoothing / c) * w[None, :], (n, 1))
    t[np.arange(n), y] += (1.0 - label_smoothing) * w[y]
    denom = w[y].sum()
    loss = -(t * logp).sum() / denom
    dlogits = (prob * t.sum(axis=1, keepdims=True) - t) / denom
    return loss, dlogits


def np_head_forward_backward(p, img_feat, txt_feat, reverse=True, features_only=False,
                             cross_attention_only=False, labels=None, dlogits=None,
                             class_weight=None, label_smoothing=0.0, drop_mask=None, drop_scale=1.0):
    """float64 forward (+ optional backward when `labels` or `dlogits` is given).
    Returns dict(logits, loss, grads, d_img, d_txt, dlogits)."""
    p = _np(p)
    img = np.asarray(img_feat, dtype=np.float64)
    txt = np.asarray(txt_feat, dtype=np.float64)
    B = img.shape[0]
    n_i = np.sqrt((img ** 2).sum(axis=1, keepdims=True))
    n_t = np.sqrt((txt ** 2).sum(axis=1, keepdims=True))
    img_n, txt_n = img / n_i, txt / n_t
    xt = txt_n.reshape(B, NUM_PATCHES, -1)
    xi = img_n.reshape(B, NUM_PATCHES, -1)
    t_sa, c_t = _np_attn_fwd(xt, xt, p, SA_TEXT, False)
    i_sa, c_i = _np_attn_fwd(xi, xi, p, SA_IMAGE, False)
    t_i, c_1 = _np_attn_fwd(t_sa, i_sa, p, CA_1, reverse)
    i_t, c_2 = _np_attn_fwd(i_sa, t_sa, p, CA_2, reverse)
    t_i_f, i_t_f = t_i.reshape(B, -1), i_t.reshape(B, -1)
    if features_only:
        cat = np.concatenate((img_n, txt_n), axis=1)
    elif cross_attention_only:
        cat = np.concatenate((t_i_f, i_t_f), axis=1)
    else:
        cat = np.concatenate((t_i_f, i_t_f, img_n, txt_n), axis=1)
    m = np.ones_like(cat) if drop_mask is None else np.asarray(drop_mask, dtype=np.float64) * drop_scale
    catd = cat * m
    fin = final_linear_name(features_only, cross_attention_only)
    wf, bf = p[f"{fin}.weight"], p[f"{fin}.bias"]
    logits = catd @ wf.T + bf
    out = dict(logits=logits, loss=None, grads=None, d_img=None, d_txt=None, dlogits=None)
    if labels is not None:
        out["loss"], dlogits = np_cross_entropy_fwd_bwd(logits, labels, class_weight, label_smoothing)
    if dlogits is None:
        return out
    dlogits = np.asarray(dlogits, dtype=np.float64)
    out["dlogits"] = dlogits
    grads = {k: np.zeros_like(v) for k, v in p.items()}
    grads[f"{fin}.weight"] = dlogits.T @ catd
    grads[f"{fin}.bias"] = dlogits.sum(axis=0)
    dcat = (dlogits @ wf) * m
    ca_w = 48 * NUM_PATCHES
    d_img_n = np.zeros_like(img_n)
    d_txt_n = np.zeros_like(txt_n)
    d_ti = np.zeros_like(t_i)
    d_it = np.zeros_like(i_t)
    if features_only:
        d_img_n += dcat[:, :img.shape[1]]
        d_txt_n += dcat[:, img.shape[1]:]
    else:
        d_ti = dcat[:, :ca_w].reshape(t_i.shape)
        d_it = dcat[:, ca_w:2 * ca_w].reshape(i_t.shape)
        if not cross_attention_only:
            d_img_n += dcat[:, 2 * ca_w:2 * ca_w + img.shape[1]]
            d_txt_n += dcat[:, 2 * ca_w + img.shape[1]:]
    if not features_only:
        d_t_sa = np.zeros_like(t_sa)
        d_i_sa = np.zeros_like(i_sa)
        dq1, dkv1 = _np_attn_bwd(d_ti, c_1, p, CA_1, grads)
        d_t_sa += dq1
        d_i_sa += dkv1
        dq2, dkv2 = _np_attn_bwd(d_it, c_2, p, CA_2, grads)
        d_i_sa += dq2
        d_t_sa += dkv2
        dxq, dxkv = _np_attn_bwd(d_t_sa, c_t, p, SA_TEXT, grads)
        d_txt_n += (dxq + dxkv).reshape(B, -1)
        dxq, dxkv = _np_attn_bwd(d_i_sa, c_i, p, SA_IMAGE, grads)
        d_img_n += (dxq + dxkv).reshape(B, -1)
    out["grads"] = grads
    out["d_img"] = (d_img_n - img_n * (img_n * d_img_n).sum(axis=1, keepdims=True)) / n_i
    out["d_txt"] = (d_txt_n - txt_n * (txt_n * d_txt_n).sum(axis=1, keepdims=True)) / n_t
    return out


# =====================================================================================================
# Hierarchical late-fusion head (reference multimodal_model.py:729-818, the second --late_fusion value)
# =====================================================================================================
HIER_IMG_SEGMENTS = (1280, 2560, 2048)     # pooled, AvgPool(7) of the 160-channel stage, AvgPool(6) of the 512-channel stage
HIER_TXT_SEGMENTS = (768, 768, 768)        # CLS of the last layer, of hidden_states[2], of hidden_states[4]
HIER_D_IMG, HIER_D_TXT, HIER_HIDDEN = sum(HIER_IMG_SEGMENTS), sum(HIER_TXT_SEGMENTS), 512
HIER_PARAM_NAMES = ("final_hierarchical_image.weight", "final_hierarchical_image.bias",
                    "final_hierarchical_text.weight", "final_hierarchical_text.bias",
                    "final_hierarchical_all.weight", "final_hierarchical_all.bias")


def init_hier_params(n_classes: int = 4, seed: int = 0, dtype=torch.float32, bias_gap: float = 0.0) -> Dict[str, torch.Tensor]:
    """Random parameters of the three Linear layers of the hierarchical head (multimodal_model.py:294-296),
    torch.nn.Linear-like scale, a pure function of the seed (the 12 MB weight is never stored in a fixture).

    bias_gap pushes the two hidden-layer biases away from zero (b += sign(b) * bias_gap).  At init scale the hidden
    pre-activations are ~N(0, 0.013): a bf16 GEMM flips the ReLU of the few units that sit within its rounding error of
    zero, which changes whole rows of the weight gradient.  With a gap of a few sigma the ReLU pattern is decided by the
    bias and element-wise gradient parity is meaningful (the analogue of qk_gain for the attention head)."""
    g = torch.Generator().manual_seed(10_000 + seed)

    def lin(out_f, in_f):
        k = 1.0 / math.sqrt(in_f)
        w = (torch.rand(out_f, in_f, generator=g, dtype=torch.float64) * 2 - 1) * k
        b = (torch.rand(out_f, generator=g, dtype=torch.float64) * 2 - 1) * k
        return w.to(dtype), b.to(dtype)

    p: Dict[str, torch.Tensor] = {}
    p["final_hierarchical_image.weight"], p["final_hierarchical_image.bias"] = lin(HIER_HIDDEN, HIER_D_IMG)
    p["final_hierarchical_text.weight"], p["final_hierarchical_text.bias"] = lin(HIER_HIDDEN, HIER_D_TXT)
    p["final_hierarchical_all.weight"], p["final_hierarchical_all.bias"] = lin(n_classes, 2 * HIER_HIDDEN)
    if bias_gap:
        for k in ("final_hierarchical_image.bias", "final_hierarchical_text.bias"):
            p[k] = p[k] + torch.sign(p[k]) * bias_gap
    return p


def hier_forward(p: Dict[str, torch.Tensor], img_segments, txt_segments,
                 drop_mask: Optional[torch.Tensor] = None, drop_scale: float = 1.0) -> torch.Tensor:
    """Hierarchical.forward after the backbones and the two AvgPool2d (multimodal_model.py:777-816).

    img_segments: (pooled [B,1280], stage-3 pooled+flattened [B,2560], stage-6 pooled+flattened [B,2048]);
    txt_segments: (CLS last layer, CLS hidden_states[2], CLS hidden_states[4]), each [B,768].
    drop_mask: keep-mask [B, 5888 + 2304] (image concat columns first) of the two self.drop calls (:805-806)."""
    img = torch.cat([l2_normalise(s) for s in img_segments], dim=1)        # :777-782, :791-796 (order: pooled, s3, s6)
    txt = torch.cat([l2_normalise(s) for s in txt_segments], dim=1)        # :784-789, :798-803
    if drop_mask is not None:
        m = drop_mask.to(img.dtype)
        img = img * m[:, :HIER_D_IMG] * drop_scale                         # :805
        txt = txt * m[:, HIER_D_IMG:] * drop_scale                         # :806
    h_img = torch.relu(img @ p["final_hierarchical_image.weight"].T + p["final_hierarchical_image.bias"])   # :808, :811
    h_txt = torch.relu(txt @ p["final_hierarchical_text.weight"].T + p["final_hierarchical_text.bias"])     # :809, :812
    return torch.cat((h_img, h_txt), dim=1) @ p["final_hierarchical_all.weight"].T + p["final_hierarchical_all.bias"]   # :814-816


def hier_loss_and_grads(p: Dict[str, torch.Tensor], img_segments, txt_segments, labels: torch.Tensor,
                        class_weight: Optional[torch.Tensor] = None, label_smoothing: float = 0.0,
                        drop_mask: Optional[torch.Tensor] = None, drop_scale: float = 1.0, dtype=torch.float64,
                        feature_grads: bool = False):
    """forward + CrossEntropyLoss + backward (main_both.py:106-112) of the hierarchical head through autograd,
    in float64 by default.  Returns (logits, loss, grads{name: tensor}) (+ the six feature gradients on request)."""
    pp = {k: v.detach().to(dtype).clone().requires_grad_(True) for k, v in p.items() if k in HIER_PARAM_NAMES}
    xi = [s.detach().to(dtype).clone().requires_grad_(feature_grads) for s in img_segments]
    xt = [s.detach().to(dtype).clone().requires_grad_(feature_grads) for s in txt_segments]
    logits = hier_forward(pp, xi, xt, drop_mask, drop_scale)
    cw = None if class_weight is None else class_weight.to(dtype)
    loss = cross_entropy(logits, labels, cw, label_smoothing)
    loss.backward()
    if feature_grads:      # fine-tune phase: d(loss)/d(each of the six pooled feature tensors), image segments first
        return logits.detach(), loss.detach(), {k: v.grad for k, v in pp.items()}, [t.grad for t in xi + xt]
    return logits.detach(), loss.detach(), {k: v.grad for k, v in pp.items()}


# =====================================================================================================
# Classic / Normalized late-fusion heads (reference multimodal_model.py:489-579)
# =====================================================================================================
FUSION_PARAM_NAMES = ("image_to_hidden_size.weight", "image_to_hidden_size.bias", "text_to_hidden_size.weight",
                      "text_to_hidden_size.bias", "concat_layer.weight", "concat_layer.bias", "fc_layer.weight",
                      "fc_layer.bias")


def init_fusion_params(d_img: int = 1280, d_txt: int = 768, hidden: int = 256, n_classes: int = 4, seed: int = 0,
                       dtype=torch.float32) -> Dict[str, torch.Tensor]:
    """Random parameters of the four Linear layers (multimodal_model.py:199-212), torch.nn.Linear-like scale, a pure
    function of the seed."""
    g = torch.Generator().manual_seed(20_000 + seed)

    def lin(out_f, in_f):
        k = 1.0 / math.sqrt(in_f)
        w = (torch.rand(out_f, in_f, generator=g, dtype=torch.float64) * 2 - 1) * k
        b = (torch.rand(out_f, generator=g, dtype=torch.float64) * 2 - 1) * k
        return w.to(dtype), b.to(dtype)

    p: Dict[str, torch.Tensor] = {}
    p["image_to_hidden_size.weight"], p["image_to_hidden_size.bias"] = lin(hidden, d_img)
    p["text_to_hidden_size.weight"], p["text_to_hidden_size.bias"] = lin(hidden, d_txt)
    p["concat_layer.weight"], p["concat_layer.bias"] = lin(hidden, 2 * hidden)
    p["fc_layer.weight"], p["fc_layer.bias"] = lin(n_classes, hidden)
    return p


def fusion_forward(p: Dict[str, torch.Tensor], img_feat: torch.Tensor, txt_feat: torch.Tensor, normalized: bool,
                   drop_mask: Optional[torch.Tensor] = None, drop_scale: float = 1.0) -> torch.Tensor:
    """EffV2MediumAndDistilbertClassic.forward (:521-531) / ...Normalized.forward (:566-579) after the backbones.
    img_feat is the POOLED image vector (the reference hands the extractor's tuple to the Linear, a TypeError as shipped)."""
    h_i = torch.nn.functional.linear(img_feat, p["image_to_hidden_size.weight"], p["image_to_hidden_size.bias"])   # :521 / :566
    h_t = torch.nn.functional.linear(txt_feat, p["text_to_hidden_size.weight"], p["text_to_hidden_size.bias"])     # :522 / :567
    if normalized:
        h_i = h_i / h_i.norm(dim=1, keepdim=True)                                                                # :569
        h_t = h_t / h_t.norm(dim=1, keepdim=True)                                                                # :570
    cat = torch.cat((h_i, h_t), dim=1)                                                                           # :524-525 / :572-573
    c = torch.nn.functional.linear(cat, p["concat_layer.weight"], p["concat_layer.bias"])                        # :527 / :575
    if drop_mask is not None:
        c = c * drop_mask.to(c.dtype) * drop_scale                                                               # :528 / :576
    return torch.nn.functional.linear(c, p["fc_layer.weight"], p["fc_layer.bias"])                               # :529 / :577


def fusion_loss_and_grads(p: Dict[str, torch.Tensor], img_feat: torch.Tensor, txt_feat: torch.Tensor, labels: torch.Tensor,
                          normalized: bool, class_weight: Optional[torch.Tensor] = None, label_smoothing: float = 0.0,
                          drop_mask: Optional[torch.Tensor] = None, drop_scale: float = 1.0, dtype=torch.float64):
    """forward + CrossEntropyLoss + backward (main_both.py:106-112) through autograd, float64 by default.
    Returns (logits, loss, grads{name: tensor}, d_img, d_txt)."""
    pp = {k: v.detach().to(dtype).clone().requires_grad_(True) for k, v in p.items() if k in FUSION_PARAM_NAMES}
    img = img_feat.detach().to(dtype).clone().requires_grad_(True)
    txt = txt_feat.detach().to(dtype).clone().requires_grad_(True)
    logits = fusion_forward(pp, img, txt, normalized, drop_mask, drop_scale)
    cw = None if class_weight is None else class_weight.to(dtype)
    loss = cross_entropy(logits, labels, cw, label_smoothing)
    loss.backward()
    return logits.detach(), loss.detach(), {k: v.grad for k, v in pp.items()}, img.grad, txt.grad
