"""Golden fixtures of the token-level attention blocks from the UNMODIFIED reference classes SelfAttention
(CVPR_code/multimodal_model.py:39-68) and ReverseCrossAttention (:71-108) on real token-sequence shapes (BASELINE.json
configs[4]): forward outputs AND the gradients loss.backward() leaves, computed by the reference's own autograd in float64.

Run in the build container only (needs /root/reference):   python tests/golden/make_golden_token.py

The classes are shape-generic (Linear on the last dimension, batched matmul, square-attention assert :93), so they are
instantiated with the token-level widths directly: no patch.  Parameters and inputs are pure functions of a seed
(token_case() below, shared with tests/test_token_golden.py), so a fixture holds outputs and gradients only; the large
weight gradients are stored as their row sums, column sums and 256 sampled entries (float32).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

LEAVES = ("W_query.weight", "W_query.bias", "W_key.weight", "W_key.bias", "W_value.weight", "W_value.bias", "norm.weight",
          "norm.bias")

# name: (kind, B, L, d_in, d_kq, d_v, reverse, seed)
CASES = {
    "sa_vit_l16": ("self", 2, 197, 1024, 128, 96, False, 1),
    "sa_roberta": ("self", 1, 256, 768, 128, 96, False, 2),
    "rca_197": ("cross", 2, 197, 96, 64, 48, True, 3),
    "ca_256": ("cross", 1, 256, 96, 64, 48, False, 4),
    "rca_ragged": ("cross", 3, 77, 96, 64, 48, True, 5),
}


def token_case(name):
    """(kind, reverse, params {leaf: tensor}, x_q, x_kv or None, d_out): everything a pure function of the case's seed.
    Activations are bf16-representable (the CUDA path's activation format), parameters fp32."""
    kind, B, L, d_in, d_kq, d_v, reverse, seed = CASES[name]
    g = torch.Generator().manual_seed(77_000 + seed)

    def lin(o, i, gain=1.0):
        k = 1.0 / i ** 0.5
        return (torch.rand(o, i, generator=g) * 2 - 1) * k * gain, (torch.rand(o, generator=g) * 2 - 1) * k

    p = {}
    p["W_query.weight"], p["W_query.bias"] = lin(d_kq, d_in, 2.0)
    p["W_key.weight"], p["W_key.bias"] = lin(d_kq, d_in, 2.0)
    p["W_value.weight"], p["W_value.bias"] = lin(d_v, d_in)
    p["norm.weight"] = 1.0 + 0.2 * torch.randn(d_v, generator=g)
    p["norm.bias"] = 0.2 * torch.randn(d_v, generator=g)
    if kind == "self":
        x_q, x_kv = torch.randn(B, L, d_in, generator=g).bfloat16().float(), None
    else:      # the cross blocks read SelfAttention outputs: post-ReLU
        x_q = torch.relu(torch.randn(B, L, d_in, generator=g)).bfloat16().float()
        x_kv = torch.relu(torch.randn(B, L, d_in, generator=g)).bfloat16().float()
    d_out = torch.randn(B, L, d_v, generator=g) / (B * L)
    return kind, reverse, p, x_q, x_kv, d_out


KINK_MARGIN = 0.1


def sample_index(size):
    return np.random.default_rng(4321).choice(size, min(256, size), replace=False)


def summarise(name, t):
    """Small stand-in for a large gradient: row sums, column sums and sampled entries (2-D), or the tensor itself."""
    a = t.detach().double().numpy()
    if a.ndim == 2 and a.size > 20_000:
        return {f"{name}#rows": a.sum(1).astype(np.float32), f"{name}#cols": a.sum(0).astype(np.float32),
                f"{name}#sample": a.reshape(-1)[sample_index(a.size)].astype(np.float32)}
    return {name: a.astype(np.float32)}


def main():
    sys.path.insert(0, "/root/reference")
    import CVPR_code.multimodal_model as mm
    for name in CASES:
        kind, reverse, p, x_q, x_kv, d_out = token_case(name)
        d_kq, d_in = p["W_query.weight"].shape
        d_v = p["W_value.weight"].shape[0]
        if kind == "self":
            m = mm.SelfAttention(d_in, d_kq, d_v, name)                     # (:40: d_in, d_out_kq, d_out_v, name)
        else:
            m = mm.ReverseCrossAttention(d_in, d_in, d_kq, d_v, reverse)     # (:72: d_in_x1, d_in_x2, d_out_kq, d_out_v, reverse)
        m = m.double()
        missing, unexpected = m.load_state_dict({k: v.double() for k, v in p.items()}, strict=True)
        xq = x_q.double().requires_grad_(True)
        xkv = x_kv.double().requires_grad_(True) if x_kv is not None else None
        pre = {}
        m.norm.register_forward_hook(lambda mod, inp, y: pre.__setitem__("y", y.detach()))
        out = m(xq) if kind == "self" else m(xq, xkv)
        # d_out is zeroed where the LayerNorm output sits within 0.1 of the ReLU kink: a bf16 forward (error up to a few 1e-2)
        # may put such an element on the other side of the kink, which is a property of comparing across precisions, not of
        # a backward.  The mask travels with the fixture; gates that are clearly open or clearly closed are all exercised.
        mask = (pre["y"].abs() > KINK_MARGIN)
        (out * (d_out.double() * mask)).sum().backward()
        rec = {"out": out.detach().numpy().astype(np.float32), "dmask": np.packbits(mask.numpy().reshape(-1))}
        for leaf in LEAVES:
            mod, attr = leaf.split(".")
            rec.update(summarise("g." + leaf, getattr(getattr(m, mod), attr).grad))
        rec.update(summarise("g.x_q", xq.grad.reshape(-1, xq.shape[-1])))
        if xkv is not None:
            rec.update(summarise("g.x_kv", xkv.grad.reshape(-1, xkv.shape[-1])))
        path = os.path.join(HERE, f"token_{name}.npz")
        np.savez_compressed(path, **rec)
        print(name, {k: v.shape for k, v in rec.items() if k in ("out", "g.W_query.weight#rows", "g.norm.bias")},
              os.path.getsize(path) // 1024, "KB")


if __name__ == "__main__":
    main()
