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


# ---- backward (mmrca_token_attention_backward) ---------------------------------------------------------------------------
def _oracle_grads(fn, xs, p, prefix, d_out):
    """float64 autograd through the oracle restatement: gradients of sum(out * d_out) w.r.t. the parameters and inputs."""
    p64 = {k: v.double().requires_grad_(True) for k, v in p.items()}
    xs64 = [x.double().requires_grad_(True) for x in xs]
    out = fn(*xs64, p64, prefix)
    (out * d_out.double()).sum().backward()
    return {k: v.grad for k, v in p64.items()}, [x.grad for x in xs64], out.detach()


def _rel(a, ref):
    return ((a.double().cpu() - ref).norm() / ref.norm().clamp_min(1e-30)).item()


# bf16 operands, fp32 accumulation, float64 oracle: per-tensor relative L2 error of the gradients (measured 4e-4 .. 1e-2,
# tools/diag_token_bwd.py).  The blocks end in a ReLU: an output within the bf16 forward error of zero can sit on the other
# side of the kink in the product than in the float64 oracle, and each such element moves a whole d_out entry in or out
# of the gradient (5e-2 .. 2e-1 relative on these small random problems - a property of comparing across precisions at
# a kink, not of the backward).  The tests therefore zero d_out where EITHER forward has the gate closed, so both
# backward passes run with the same gates - closed gates are still exercised, the coin flips at the kink are not.
TOK_GRAD_LIMIT = 2e-2


def _gate_mask(out_product, out_oracle):
    return ((out_product.cpu() > 0) & (out_oracle > 0)).float()


@pytest.mark.parametrize("B,L,K", [(3, 197, 1024), (2, 256, 768), (4, 16, 96), (1, 77, 208), (2, 129, 1024), (3, 5, 32), (2, 128, 64)],
                         ids=["vit_l16", "roberta", "pseudo_tokens", "ragged", "two_tiles_ragged", "tiny", "one_full_tile"])
def test_token_self_attention_backward_matches_oracle(pkg, B, L, K):
    from garbage_classification_rca_b200 import functional as F
    p = _block_params("sa", K, K, 128, 96, seed=L + K, gain=2.0)
    g = torch.Generator().manual_seed(B + L)
    x = torch.randn(B, L, K, generator=g).bfloat16()
    d_out = torch.randn(B, L, 96, generator=g) / (B * L)
    params = [p[f"sa.{l}"].cuda() for l in LEAVES]
    blk = F.TokenAttention(params, B, L, training=True)
    out = blk(x.cuda())
    d_out = d_out * _gate_mask(out, orc.self_attention(x.float(), p, "sa"))
    grads = [torch.zeros_like(t) for t in params]
    dx, _ = blk.backward(d_out.cuda(), grads, need_dx_q=True)
    torch.cuda.synchronize()
    ref_g, ref_dx, ref_out = _oracle_grads(lambda x_, p_, pre: orc.self_attention(x_, p_, pre), [x.float()], p, "sa", d_out)
    _check(out, ref_out.float(), "SA training forward")
    for l, gt in zip(LEAVES, grads):
        r = ref_g[f"sa.{l}"]
        if l == "W_key.bias":       # analytically zero (softmax is shift-invariant): absolute check against the scale of d(b_q)
            assert gt.abs().max().item() < 1e-2 * ref_g["sa.W_query.bias"].abs().max().item() + 1e-7
            continue
        assert _rel(gt, r) < TOK_GRAD_LIMIT, f"SA d({l}) B={B} L={L} K={K}: rel {_rel(gt, r):.3e}"
    assert _rel(dx, ref_dx[0]) < TOK_GRAD_LIMIT, f"SA d(x): rel {_rel(dx, ref_dx[0]):.3e}"
    # the backward accumulates into the parameter gradients
    blk(x.cuda())
    blk.backward(d_out.cuda(), grads)
    torch.cuda.synchronize()
    assert _rel(grads[0], 2 * ref_g["sa.W_query.weight"]) < TOK_GRAD_LIMIT


@pytest.mark.parametrize("reverse", [True, False], ids=["rca", "ca"])
@pytest.mark.parametrize("B,L", [(3, 197), (2, 256), (4, 16), (1, 100), (2, 3), (1, 128)])
def test_token_cross_attention_backward_matches_oracle(pkg, B, L, reverse):
    from garbage_classification_rca_b200 import functional as F
    p = _block_params("ca", 96, 96, 64, 48, seed=7 * L + int(reverse), gain=2.0)
    g = torch.Generator().manual_seed(L)
    x1 = torch.relu(torch.randn(B, L, 96, generator=g)).bfloat16()
    x2 = torch.relu(torch.randn(B, L, 96, generator=g)).bfloat16()
    d_out = torch.randn(B, L, 48, generator=g) / (B * L)
    params = [p[f"ca.{l}"].cuda() for l in LEAVES]
    blk = F.TokenAttention(params, B, L, reverse=reverse, training=True)
    out = blk(x1.cuda(), x2.cuda())
    d_out = d_out * _gate_mask(out, orc.reverse_cross_attention(x1.float(), x2.float(), p, "ca", reverse))
    grads = [torch.zeros_like(t) for t in params]
    dx1, dx2 = blk.backward(d_out.cuda(), grads, need_dx_q=True, need_dx_kv=True)
    torch.cuda.synchronize()
    ref_g, ref_dx, _ = _oracle_grads(lambda a, b, p_, pre: orc.reverse_cross_attention(a, b, p_, pre, reverse),
                                     [x1.float(), x2.float()], p, "ca", d_out)
    for l, gt in zip(LEAVES, grads):
        r = ref_g[f"ca.{l}"]
        if l == "W_key.bias":
            assert gt.abs().max().item() < 1e-2 * ref_g["ca.W_query.bias"].abs().max().item() + 1e-7
            continue
        assert _rel(gt, r) < TOK_GRAD_LIMIT, f"CA d({l}) reverse={reverse} B={B} L={L}: rel {_rel(gt, r):.3e}"
    assert _rel(dx1, ref_dx[0]) < TOK_GRAD_LIMIT, f"CA d(x_q): rel {_rel(dx1, ref_dx[0]):.3e}"
    assert _rel(dx2, ref_dx[1]) < TOK_GRAD_LIMIT, f"CA d(x_kv): rel {_rel(dx2, ref_dx[1]):.3e}"


def test_token_attention_under_autograd(pkg):
    """apply(): gradients reach the parameter tensors and the inputs like loss.backward() through the reference modules."""
    from garbage_classification_rca_b200 import functional as F
    B, L, K = 2, 50, 128
    p = _block_params("sa", K, K, 128, 96, seed=3)
    params = [p[f"sa.{l}"].cuda().requires_grad_(True) for l in LEAVES]
    x = torch.randn(B, L, K, generator=torch.Generator().manual_seed(1)).bfloat16().float()
    xg = x.cuda().requires_grad_(True)
    blk = F.TokenAttention(params, B, L, training=True)
    out = blk.apply(xg)
    out.square().sum().backward()
    p64 = {k: v.double().requires_grad_(True) for k, v in p.items()}
    x64 = x.double().requires_grad_(True)
    orc.self_attention(x64, p64, "sa").square().sum().backward()
    assert _rel(params[0].grad, p64["sa.W_query.weight"].grad) < TOK_GRAD_LIMIT
    assert _rel(params[4].grad, p64["sa.W_value.weight"].grad) < TOK_GRAD_LIMIT
    assert _rel(xg.grad, x64.grad) < TOK_GRAD_LIMIT
    with pytest.raises(RuntimeError):
        F.TokenAttention([t.detach() for t in params], B, L).backward(out.detach(), [torch.zeros_like(t) for t in params])


def test_token_attention_bf16_output_into_a_caller_buffer(pkg):
    """out_dtype=bfloat16 + out=: a block writes the next block's activation format straight into a caller's buffer."""
    from garbage_classification_rca_b200 import functional as F
    B, L, K = 3, 100, 256
    p = _block_params("sa", K, K, 128, 96, seed=11)
    x = torch.randn(B, L, K, generator=torch.Generator().manual_seed(2)).bfloat16().cuda()
    params = [p[f"sa.{l}"].cuda() for l in LEAVES]
    ref = F.TokenAttention(params, B, L)(x).clone()
    ext = torch.zeros(B + 1, L, 96, dtype=torch.bfloat16, device="cuda")
    out = F.TokenAttention(params, B, L, out_dtype=torch.bfloat16)(x, out=ext[1:])
    torch.cuda.synchronize()
    assert out.data_ptr() == ext[1:].data_ptr() and ext[0].abs().max().item() == 0
    assert torch.equal(ext[1:], ref.to(torch.bfloat16))
    with pytest.raises(ValueError):
        F.TokenAttention(params, B, L, out_dtype=torch.bfloat16)(x, out=torch.zeros(B, L, 96, device="cuda"))


def test_token_block_under_cuda_graph_capture(pkg):
    """Forward + backward of a cross block captured into a CUDA graph (the calls fork GEMMs onto library-owned side streams and
    join them with events: legal under capture) and replayed: same results as the eager calls."""
    from garbage_classification_rca_b200 import functional as F
    B, L = 3, 150
    p = _block_params("ca", 96, 96, 64, 48, seed=21)
    g = torch.Generator().manual_seed(9)
    x1 = torch.relu(torch.randn(B, L, 96, generator=g)).bfloat16().cuda()
    x2 = torch.relu(torch.randn(B, L, 96, generator=g)).bfloat16().cuda()
    d_out = (torch.randn(B, L, 48, generator=g) / (B * L)).cuda()
    params = [p[f"ca.{l}"].cuda() for l in LEAVES]
    blk = F.TokenAttention(params, B, L, reverse=True, training=True)
    grads = [torch.zeros_like(t) for t in params]
    out_eager = blk(x1, x2).clone()
    dx1_e, dx2_e = blk.backward(d_out, grads, True, True)
    torch.cuda.synchronize()
    grads_eager = [t.clone() for t in grads]
    for t in grads:
        t.zero_()
    blk.refresh_weights()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):          # warm-up on the capture stream (lazy initialisation must not happen inside the capture)
        blk(x1, x2)
        blk.backward(d_out, [torch.zeros_like(t) for t in params], True, True)
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    blk.refresh_weights()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        out_g = blk(x1, x2)
        dx1_g, dx2_g = blk.backward(d_out, grads, True, True)
    for t in grads:
        t.zero_()
    graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(out_g, out_eager)
    assert torch.allclose(dx1_g, dx1_e, rtol=1e-5, atol=1e-8) and torch.allclose(dx2_g, dx2_e, rtol=1e-5, atol=1e-8)
    for a, b in zip(grads, grads_eager):
        assert torch.allclose(a, b, rtol=1e-4, atol=1e-7)
