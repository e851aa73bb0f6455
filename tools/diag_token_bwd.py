"""Development diagnostic: per-tensor relative errors of the token-level attention backward against the float64 restatement."""
import sys
sys.path.insert(0, ".")
import torch
from tests.test_token_gpu import _block_params, LEAVES, _oracle_grads, _rel
from oracle import mmrca_oracle as orc
from garbage_classification_rca_b200 import functional as F

BETA = float(sys.argv[1]) if len(sys.argv) > 1 else 0.0      # shift of norm.bias: > 3 keeps the ReLU gates away from the kink
for (B, L, K, gain) in [(3, 197, 1024, 2.0), (3, 197, 1024, 1.0), (2, 256, 768, 2.0), (4, 16, 96, 2.0), (1, 77, 208, 2.0), (2, 129, 1024, 2.0)]:
    p = _block_params("sa", K, K, 128, 96, seed=L + K, gain=gain)
    p["sa.norm.bias"] += BETA
    g = torch.Generator().manual_seed(B + L)
    x = torch.randn(B, L, K, generator=g).bfloat16()
    d_out = torch.randn(B, L, 96, generator=g) / (B * L)
    params = [p[f"sa.{l}"].cuda() for l in LEAVES]
    blk = F.TokenAttention(params, B, L, training=True)
    blk(x.cuda())
    grads = [torch.zeros_like(t) for t in params]
    dx, _ = blk.backward(d_out.cuda(), grads, need_dx_q=True)
    torch.cuda.synchronize()
    ref_g, ref_dx, _ = _oracle_grads(lambda x_, p_, pre: orc.self_attention(x_, p_, pre), [x.float()], p, "sa", d_out)
    print(f"SA B={B} L={L} K={K} gain={gain}: " + " ".join(f"{l.split('.')[0][2:] + l[-1]}={_rel(gt, ref_g['sa.' + l]):.2e}" for l, gt in zip(LEAVES, grads)),
          f"dx={_rel(dx, ref_dx[0]):.2e}", f"|dbk|={grads[3].abs().max().item():.1e} |dbq|={grads[1].abs().max().item():.1e}")
for reverse in (True, False):
    for (B, L) in [(3, 197), (2, 256), (4, 16), (1, 100)]:
        p = _block_params("ca", 96, 96, 64, 48, seed=7 * L + int(reverse), gain=2.0)
        p["ca.norm.bias"] += BETA
        g = torch.Generator().manual_seed(L)
        x1 = torch.relu(torch.randn(B, L, 96, generator=g)).bfloat16()
        x2 = torch.relu(torch.randn(B, L, 96, generator=g)).bfloat16()
        d_out = torch.randn(B, L, 48, generator=g) / (B * L)
        params = [p[f"ca.{l}"].cuda() for l in LEAVES]
        blk = F.TokenAttention(params, B, L, reverse=reverse, training=True)
        blk(x1.cuda(), x2.cuda())
        grads = [torch.zeros_like(t) for t in params]
        dx1, dx2 = blk.backward(d_out.cuda(), grads, need_dx_q=True, need_dx_kv=True)
        torch.cuda.synchronize()
        ref_g, ref_dx, _ = _oracle_grads(lambda a, b, p_, pre: orc.reverse_cross_attention(a, b, p_, pre, reverse),
                                         [x1.float(), x2.float()], p, "ca", d_out)
        print(f"CA rev={reverse} B={B} L={L}: " + " ".join(f"{l.split('.')[0][2:] + l[-1]}={_rel(gt, ref_g['ca.' + l]):.2e}" for l, gt in zip(LEAVES, grads)),
              f"dx1={_rel(dx1, ref_dx[0]):.2e} dx2={_rel(dx2, ref_dx[1]):.2e}")
