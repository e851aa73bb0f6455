"""Shared helpers of the parity tests (test infrastructure; the only place besides bench.py's baseline
legs and smoke() that touches oracle/)."""
import glob
import os

import numpy as np
import torch

from oracle import mmrca_oracle as orc

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
HEAD_CASES = sorted(os.path.basename(f)[5:-4] for f in glob.glob(os.path.join(GOLDEN, "head_*.npz")))

# tolerances of BASELINE.json north_star
LOGITS_REL_FP32 = 1e-4      # fp32 logits within 1e-4 relative
GRAD_REL = 1e-2             # fusion-head gradients within 1e-2 relative
# what the fp32 kernels are actually held to (max-abs error over max-abs reference, per tensor)
GRAD_REL_FP32_TIGHT = 2e-4


def load_case(name):
    d = np.load(os.path.join(GOLDEN, f"head_{name}.npz"))
    rev, fo, co = (bool(x) for x in d["flags"])
    qk = float(d["qk_gain"]) if "qk_gain" in d.files else 1.0
    params = orc.init_head_params(features_only=fo, cross_attention_only=co, seed=int(d["seed"]), qk_gain=qk)
    return d, (rev, fo, co), params


def rel_err(a, b):
    """max |a-b| / max |b| (relative to the tensor's scale, the way the tolerances are stated)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    den = np.abs(b).max()
    if den == 0:
        return float(np.abs(a).max())
    return float(np.abs(a - b).max() / den)


def make_inputs(B, seed, d_img=1280, d_txt=768, n_classes=4):
    g = torch.Generator().manual_seed(seed)
    img = torch.randn(B, d_img, generator=g) * 0.7 + 0.1
    txt = torch.randn(B, d_txt, generator=g) * 1.3 - 0.05
    labels = torch.randint(0, n_classes, (B,), generator=g)
    return img, txt, labels


def is_key_bias(name):
    # d/d(W_key.bias) is analytically zero (softmax is invariant to a per-row constant); the reference
    # only holds rounding noise there.
    return name.endswith("W_key.bias")


def assert_grad_close(name, ours, ref, tol, scale=1.0):
    """Per-tensor gradient check.  Tensors whose reference gradient is pure rounding noise (W_key.bias, or
    the query/key weights under uniform attention) are held to a noise floor relative to `scale` (the
    largest gradient magnitude of the same backward call) instead."""
    ours = np.asarray(ours, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    if is_key_bias(name) or np.abs(ref).max() < 1e-9:
        floor = 2e-6 * max(1.0, scale)
        assert np.abs(ours).max() < floor, f"{name}: expected ~0 (< {floor:.1e}), got {np.abs(ours).max():.3e}"
        return
    e = rel_err(ours, ref)
    assert e < tol, f"{name}: rel err {e:.3e} >= {tol:.1e} (max|ref| = {np.abs(ref).max():.3e})"


def grad_summary(ours, ref, significant=1e-3):
    """Whole-gradient error metrics for the bf16 pipeline (dicts name -> array).  W_key.bias is excluded (its
    reference gradient is rounding noise); the per-tensor figure only looks at tensors whose largest reference
    entry is at least `significant` x the largest entry of the whole gradient."""
    names = [n for n in ref if n in ours and not is_key_bias(n)]
    r = {n: np.asarray(ref[n], dtype=np.float64).ravel() for n in names}
    o = {n: np.asarray(ours[n], dtype=np.float64).ravel() for n in names}
    scale = max(np.abs(v).max() for v in r.values())
    fr, fo = np.concatenate([r[n] for n in names]), np.concatenate([o[n] for n in names])
    worst, worst_name = 0.0, ""
    for n in names:
        if np.abs(r[n]).max() >= significant * scale:
            e = np.linalg.norm(o[n] - r[n]) / np.linalg.norm(r[n])
            if e > worst:
                worst, worst_name = float(e), n
    return {"flat_l2_rel": float(np.linalg.norm(fo - fr) / np.linalg.norm(fr)),
            "cos": float(fo @ fr / (np.linalg.norm(fo) * np.linalg.norm(fr))),
            "worst_l2_rel": worst, "worst_name": worst_name,
            "max_err_over_global": float(max(np.abs(o[n] - r[n]).max() for n in names) / scale)}
