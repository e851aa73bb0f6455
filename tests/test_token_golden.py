"""Token-level attention blocks against fixtures produced by the UNMODIFIED reference classes (SelfAttention,
ReverseCrossAttention; tests/golden/make_golden_token.py, float64 autograd) on BASELINE.json configs[4] shapes:
  CPU: the oracle restatement reproduces the reference's outputs AND gradients (pins the oracle the GPU tests use);
  GPU: the CUDA forward + backward (mmrca_token_attention_forward / _backward) against the same fixtures.
d_out is zeroed where the reference's LayerNorm output sits within 0.1 of the ReLU kink (mask stored in the fixture)."""
import os
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
from make_golden_token import CASES, LEAVES, sample_index, token_case      # noqa: E402

from oracle import mmrca_oracle as orc      # noqa: E402


def _load(name):
    z = np.load(os.path.join(HERE, "golden", f"token_{name}.npz"))
    kind, reverse, p, x_q, x_kv, d_out = token_case(name)
    mask = torch.from_numpy(np.unpackbits(z["dmask"])[:d_out.numel()].reshape(d_out.shape).astype(np.float32))
    return z, kind, reverse, p, x_q, x_kv, d_out * mask


def _compare(z, key, got, limit, what):
    """got: tensor; the fixture holds it whole or as row sums / column sums / sampled entries."""
    a = got.detach().double().cpu().numpy()
    a = a.reshape(-1, a.shape[-1]) if a.ndim > 2 else a
    frob = 0.0
    if key in z.files:
        pairs = [(a, z[key])]
    else:
        pairs = [(a.sum(1), z[key + "#rows"]), (a.sum(0), z[key + "#cols"]), (a.reshape(-1)[sample_index(a.size)], z[key + "#sample"])]
        # Frobenius norm of the whole tensor, estimated from the sampled entries: rounding errors do not cancel in a sum the
        # way the exact values can (the column sums of dW_value are analytically zero: a LayerNorm's input gradient sums
        # to zero over the columns), so a summary is held to `limit` relative to max(its own norm, the tensor's norm)
        frob = np.linalg.norm(z[key + "#sample"].astype(np.float64)) * np.sqrt(a.size / len(z[key + "#sample"]))
    for got_v, ref_v in pairs:
        ref_v = ref_v.astype(np.float64)
        scale = max(np.linalg.norm(ref_v), frob, 1e-30)
        rel = np.linalg.norm(got_v - ref_v) / scale
        assert rel < limit, f"{what} {key}: rel {rel:.3e}"


def test_fixture_set_is_complete():
    for name in CASES:
        assert os.path.exists(os.path.join(HERE, "golden", f"token_{name}.npz")), name


@pytest.mark.parametrize("name", list(CASES))
def test_oracle_matches_reference_token_fixture(name):
    z, kind, reverse, p, x_q, x_kv, d_out = _load(name)
    p64 = {"b." + k: v.double().requires_grad_(True) for k, v in p.items()}
    xq = x_q.double().requires_grad_(True)
    xkv = x_kv.double().requires_grad_(True) if x_kv is not None else None
    out = orc.self_attention(xq, p64, "b") if kind == "self" else orc.reverse_cross_attention(xq, xkv, p64, "b", reverse)
    assert np.abs(out.detach().numpy() - z["out"]).max() < 1e-5
    (out * d_out.double()).sum().backward()
    for leaf in LEAVES:
        if leaf == "W_key.bias":       # analytically zero: absolute
            assert p64["b." + leaf].grad.abs().max().item() < 1e-12 and np.abs(z["g." + leaf]).max() < 1e-9
            continue
        _compare(z, "g." + leaf, p64["b." + leaf].grad, 1e-5, name)
    _compare(z, "g.x_q", xq.grad, 1e-5, name)
    if xkv is not None:
        _compare(z, "g.x_kv", xkv.grad, 1e-5, name)


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(CASES))
def test_cuda_token_blocks_match_reference_fixture(native_lib, name):
    from garbage_classification_rca_b200 import functional as F
    z, kind, reverse, p, x_q, x_kv, d_out = _load(name)
    params = [p[l].cuda() for l in LEAVES]
    B, L, _ = x_q.shape
    blk = F.TokenAttention(params, B, L, reverse=reverse, training=True)
    out = blk(x_q.cuda(), x_kv.cuda() if x_kv is not None else None)
    torch.cuda.synchronize()
    err = np.abs(out.cpu().numpy() - z["out"])
    assert err.max() < 8e-2 and err.mean() < 8e-3, f"{name}: forward max {err.max():.3e} mean {err.mean():.3e}"
    grads = [torch.zeros_like(t) for t in params]
    dxq, dxkv = blk.backward(d_out.cuda(), grads, True, x_kv is not None)
    torch.cuda.synchronize()
    for leaf, g in zip(LEAVES, grads):
        if leaf == "W_key.bias":
            assert g.abs().max().item() < 1e-2 * np.abs(z["g.W_query.bias"]).max() + 1e-7
            continue
        _compare(z, "g." + leaf, g, 3e-2, name)
    _compare(z, "g.x_q", dxq, 3e-2, name)
    if x_kv is not None:
        _compare(z, "g.x_kv", dxkv, 3e-2, name)
