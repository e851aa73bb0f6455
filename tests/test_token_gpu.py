"""GPU: token-level attention blocks (BASELINE.json configs[4]; SURVEY.md §8 d cfg 5) through the C ABI
(mmrca_token_attention_forward): the TMA-fed tcgen05 projection GEMM + the per-(sample, query tile) attention kernel,
against the shape-generic oracle restatement of the reference classes (multimodal_model.py:39-108; pinned to the
reference's SelfAttention / ReverseCrossAttention by tests/golden/components.npz) on square L = 197 (ViT-L/16 patch
tokens, d 1024) and L = 256 (RoBERTa tokens, d 768), plus ragged small shapes.  bf16 operands, fp32 accumulate:
LayerNorm'd O(1) outputs are held to 8e-2 max / 8e-3 mean absolute error (the same limits as the fused head's SA images)."""
import pytest
import torch

from oracle import mmrca_oracle as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pkg(native_lib):
    import garbage_classification_rca_b200 as g
    return g


def _block_params(prefix, d_q, d_kv, d_kq, d_v, seed, gain=1.0):
    g = torch.Generator().manual_seed(seed)

    def lin(o, i):
        k = 1.0 / i ** 0.5
        return (torch.rand(o, i, generator=g) * 2 - 1) * k, (torch.rand(o, generator=g) * 2 - 1) * k

    p = {}
    p[f"{prefix}.W_query.weight"], p[f"{prefix}.W_query.bias"] = lin(d_kq, d_q)
    p[f"{prefix}.W_key.weight"], p[f"{prefix}.W_key.bias"] = lin(d_kq, d_kv)
    p[f"{prefix}.W_value.weight"], p[f"{prefix}.W_value.bias"] = lin(d_v, d_kv)
    p[f"{prefix}.W_query.weight"] *= gain
    p[f"{prefix}.W_key.weight"] *= gain
    p[f"{prefix}.norm.weight"] = 1.0 + 0.2 * torch.randn(d_v, generator=g)
    p[f"{prefix}.norm.bias"] = 0.2 * torch.randn(d_v, generator=g)
    return p


LEAVES = ("W_query.weight", "W_query.bias", "W_key.weight", "W_key.bias", "W_value.weight", "W_value.bias", "norm.weight",
          "norm.bias")


def _check(out, ref, what):
    err = (out.cpu() - ref).abs()
    assert err.max().item() < 8e-2, f"{what}: max abs err {err.max().item():.3e}"
    assert err.mean().item() < 8e-3, f"{what}: mean abs err {err.mean().item():.3e}"


@pytest.mark.parametrize("B,L,K", [(3, 197, 1024), (2, 256, 768), (5, 16, 80), (1, 77, 200), (4, 128, 64), (2, 129, 1024)],
                         ids=["vit_l16", "roberta", "pseudo_tokens", "ragged", "one_tile", "two_tiles_ragged"])
def test_token_self_attention_matches_oracle(pkg, B, L, K):
    from garbage_classification_rca_b200 import functional as F
    p = _block_params("sa", K, K, 128, 96, seed=L + K, gain=2.0)
    g = torch.Generator().manual_seed(B + L)
    x = torch.randn(B, L, K, generator=g).bfloat16()
    blk = F.TokenAttention([p[f"sa.{l}"].cuda() for l in LEAVES], B, L)
    out = blk(x.cuda())
    torch.cuda.synchronize()
    # the oracle sees the same bf16-rounded activations
    ref = orc.self_attention(x.float(), p, "sa")
    _check(out, ref, f"SA B={B} L={L} K={K}")
    assert blk(x.float().cuda()).equal(out)        # fp32 activations are cast at the hand-off


@pytest.mark.parametrize("reverse", [True, False], ids=["rca", "ca"])
@pytest.mark.parametrize("B,L", [(3, 197), (2, 256), (4, 16), (1, 100)])
def test_token_cross_attention_matches_oracle(pkg, B, L, reverse):
    from garbage_classification_rca_b200 import functional as F
    p = _block_params("ca", 96, 96, 64, 48, seed=7 * L + int(reverse), gain=2.0)
    g = torch.Generator().manual_seed(L)
    x1 = torch.relu(torch.randn(B, L, 96, generator=g)).bfloat16()       # SA outputs are post-ReLU
    x2 = torch.relu(torch.randn(B, L, 96, generator=g)).bfloat16()
    blk = F.TokenAttention([p[f"ca.{l}"].cuda() for l in LEAVES], B, L, reverse=reverse)
    out = blk(x1.cuda(), x2.cuda())
    torch.cuda.synchronize()
    ref = orc.reverse_cross_attention(x1.float(), x2.float(), p, "ca", reverse)
    _check(out, ref, f"CA reverse={reverse} B={B} L={L}")


def test_token_attention_rejects_what_it_does_not_cover(pkg):
    from garbage_classification_rca_b200 import functional as F
    p = _block_params("sa", 64, 64, 128, 96, seed=1)
    with pytest.raises(ValueError):
        F.TokenAttention([p[f"sa.{l}"].cuda() for l in LEAVES], 2, 300)          # L > 256
    blk = F.TokenAttention([p[f"sa.{l}"].cuda() for l in LEAVES], 2, 32)
    with pytest.raises(ValueError):
        blk(torch.zeros(2, 33, 64).cuda())                                       # shape mismatch
    with pytest.raises(RuntimeError):
        blk(torch.zeros(2, 32, 64))                                              # no CPU fallback
