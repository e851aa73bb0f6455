"""GPU parity of the fused bf16 tensor-core pipeline (mmrca_head_tc*.cuh) against the fp32 CPU oracle.

Tolerances (BASELINE.json north_star): bf16 logits within 2e-2 absolute, argmax agreement >= 99.9 %.
Intermediates (the SA output images the CA kernels consume) are decoded from the workspace and held to a
bf16-rounding bound so that a wrong tile/row mapping cannot hide behind the bias-dominated logits.

Gradients ("fusion-head gradients within 1e-2 relative"): the fp32 kernels meet it per tensor, max-norm, with
two orders of margin (tests/test_parity_gpu.py holds them to 2e-4).  The bf16 pipeline is held to it on the whole
head gradient — l2 error of the flat 94 820-vector relative to its l2 norm — for the full concat (the benchmark
configuration; measured 5e-3 .. 8e-3, dropout on or off), with per-tensor and per-entry bounds beside it.  The
cross_attention_only ablation is worse conditioned at random init (its gradients are sums of nearly cancelling
per-sample terms behind two LayerNorm backward projections, without the feature columns that dominate the full
model): ANY bf16 evaluation of the reference's formulas lands at a few per cent there — torch's own bf16 autocast
of the reference included, tools/diag_bf16.py prints it next to ours (ours 1.3e-2 .. 4.4e-2, autocast 1.4e-2 ..
2.8e-2) — so that variant gets the looser row of BF16_GRAD.  Both rows are an order of magnitude below what a
wrong row mapping, a dropped term or a mis-scaled tensor produces."""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import mmrca_oracle as orc
from tests._util import grad_summary, make_inputs

pytestmark = pytest.mark.gpu

LOGITS_ABS_BF16 = 2e-2
TILE_BYTES, CS = 12 * 2064, 2064


@pytest.fixture(scope="module")
def pkg(native_lib):
    import garbage_classification_rca_b200 as g
    return g


def decode_sa_image(ws: torch.Tensor, off: int, batch: int) -> torch.Tensor:
    """[tiles][12 column groups (stride 2064 B)][128 rows x 8 bf16] -> [batch, 16, 96] fp32."""
    tiles = (batch + 7) // 8
    raw = ws[off:off + tiles * TILE_BYTES].cpu().view(tiles, TILE_BYTES)
    out = torch.empty(tiles, 128, 96)
    for kc in range(12):
        blk = raw[:, kc * CS:kc * CS + 2048].contiguous().view(torch.bfloat16).view(tiles, 128, 8)
        out[:, :, kc * 8:(kc + 1) * 8] = blk.float()
    return out.view(tiles * 8, 16, 96)[:batch]


def run_fused_forward(g, p, img, txt, flags, training=False):
    from garbage_classification_rca_b200 import _native as N
    from garbage_classification_rca_b200 import functional as F
    rev, fo, co = flags
    names = g.head_param_names(fo, co)
    params = [p[n].cuda().contiguous() for n in names]
    B = img.shape[0]
    desc = N.HeadDesc(B, 1280, 768, 4, F.make_flags(rev, fo, co) | (N.FLAG_TRAINING if training else 0),
                      N.COMPUTE_BF16_FUSED)
    L = N.lib()
    nbytes = int(L.mmrca_head_workspace_bytes(C.byref(desc), 1 if training else 0))
    ws = torch.zeros(nbytes, dtype=torch.uint8, device="cuda")
    logits = torch.full((B, 4), float("nan"), device="cuda")
    hp = F._head_struct(params)
    imgc, txtc = img.cuda().contiguous(), txt.cuda().contiguous()
    N.check(L.mmrca_head_forward(C.byref(desc), C.byref(hp), imgc.data_ptr(), txtc.data_ptr(), None, 1.0,
                                 logits.data_ptr(), ws.data_ptr(), ws.numel(),
                                 torch.cuda.current_stream().cuda_stream), "mmrca_head_forward")
    torch.cuda.synchronize()
    offs = {k: int(L.mmrca_head_workspace_offset(C.byref(desc), 1 if training else 0, w))
            for k, w in (("t", N.WS_TEXT_SA_IMAGE), ("i", N.WS_IMAGE_SA_IMAGE))}
    return logits.cpu(), ws, offs


@pytest.mark.parametrize("B", [1, 8, 13, 200])
@pytest.mark.parametrize("qk_gain", [1.0, 40.0])
def test_fused_sa_images_and_logits(pkg, B, qk_gain):
    p = orc.init_head_params(seed=11, qk_gain=qk_gain)
    img, txt, _ = make_inputs(B, 100 + B)
    logits, ws, offs = run_fused_forward(pkg, p, img, txt, (True, False, False))
    t_ref = orc.self_attention(orc.l2_normalise(txt).reshape(B, 16, 48), p, "self_attention_text")
    i_ref = orc.self_attention(orc.l2_normalise(img).reshape(B, 16, 80), p, "self_attention_image")
    t_got = decode_sa_image(ws, offs["t"], B)
    i_got = decode_sa_image(ws, offs["i"], B)
    # LayerNorm'd O(1) outputs after three bf16-operand matmuls; the image itself is bf16 (2^-9 relative)
    for name, got, ref in (("text SA", t_got, t_ref), ("image SA", i_got, i_ref)):
        err = (got - ref).abs()
        assert err.max().item() < 8e-2, f"{name}: max abs err {err.max().item():.3e}"
        assert err.mean().item() < 8e-3, f"{name}: mean abs err {err.mean().item():.3e}"
    ref = orc.head_forward(p, img, txt, True)
    assert (logits - ref).abs().max().item() < LOGITS_ABS_BF16


@pytest.mark.parametrize("flags", [(True, False, False), (False, False, False), (True, False, True),
                                   (False, False, True), (True, True, False)],
                         ids=["rca", "ca", "rca_cross_only", "ca_cross_only", "features_only"])
def test_fused_logits_switches(pkg, flags):
    rev, fo, co = flags
    B = 333
    p = orc.init_head_params(features_only=fo, cross_attention_only=co, seed=5, qk_gain=40.0)
    img, txt, _ = make_inputs(B, 5)
    logits, _, _ = run_fused_forward(pkg, p, img, txt, flags)
    ref = orc.head_forward(p, img, txt, rev, fo, co)
    err = (logits - ref).abs().max().item()
    assert err < LOGITS_ABS_BF16, f"max abs logit err {err:.3e}"


def test_fused_argmax_agreement(pkg):
    """>= 99.9 % argmax agreement at the benchmark batch, with a sharpened classifier so that the argmax is
    decided by the attention outputs and not by the bias (SURVEY.md §8 c)."""
    B = 4096
    p = orc.init_head_params(seed=17, qk_gain=40.0)
    g = torch.Generator().manual_seed(17)
    wf = torch.randn(4, 3584, generator=g) * 1.5
    wf[:, :1536] *= 0.03         # T_I / I_T are non-negative and nearly sample-independent at random init: keep
                                 # their class offset small so the (zero-mean) features decide the argmax
    p["final_with_everything.weight"] = wf
    p["final_with_everything.bias"] = torch.zeros(4)
    img = torch.randn(B, 1280, generator=g)
    txt = torch.randn(B, 768, generator=g)
    ref = orc.head_forward(p, img, txt, True)
    logits, _, _ = run_fused_forward(pkg, p, img, txt, (True, False, False))
    assert ref.argmax(1).unique().numel() == 4
    err = (logits - ref).abs().max().item()
    assert err < LOGITS_ABS_BF16 * max(1.0, ref.abs().max().item()), f"max abs logit err {err:.3e}"
    top2 = ref.topk(2, dim=1).values
    decided = (top2[:, 0] - top2[:, 1]) > 2 * LOGITS_ABS_BF16
    agree = logits.argmax(1) == ref.argmax(1)
    assert agree[decided].float().mean().item() == 1.0
    assert agree.float().mean().item() >= 0.999


# ---- backward ------------------------------------------------------------------------------------------------------
# north_star: "fusion-head gradients within 1e-2 relative".  The bf16 pipeline is held to it NORM-WISE on the whole head
# gradient (||g - g_ref||_2 / ||g_ref||_2 over the flat 94 820-entry bucket: what an optimizer step sees) on every
# variant of configs[3]; measured 2.4e-3 (full concat) / 8.9e-3 (cross-only) at batch 200 and 0.9e-3 / 2.8e-3 at batch
# 2048 (gpurun_out/r2_diag_*.log).  Per tensor the worst L2 error is 1 - 3.5 % (full) and up to 4.6 % (cross-only, on
# 48- / 96-entry LayerNorm / bias tensors whose batch sum cancels): that residue is the bf16 rounding of the FORWARD
# activations (X, V, SA output), each worth 0.5 - 2 % on those tensors (tools/emulate_bf16.py reproduces the GPU's
# figures to 4 digits and attributes them site by site); only hi + lo pairs of every forward operand would remove it
# (DESIGN.md §2).  The fp32 kernels meet 1e-2 per tensor with two orders of magnitude to spare (test_parity_gpu.py).
BF16_GRAD = {False: dict(cos=0.99999, flat_l2_rel=5e-3, worst_l2_rel=4e-2, max_err_over_global=1e-2),    # full concat
             True: dict(cos=0.9999, flat_l2_rel=1e-2, worst_l2_rel=6e-2, max_err_over_global=2.5e-2)}    # cross-only


def check_bf16_grads(ours, ref, what, cross_only=False):
    s, lim = grad_summary(ours, ref), BF16_GRAD[bool(cross_only)]
    assert s["cos"] >= lim["cos"], f"{what}: gradient cosine {s['cos']:.5f}"
    for k in ("flat_l2_rel", "worst_l2_rel", "max_err_over_global"):
        assert s[k] <= lim[k], f"{what}: {k} = {s[k]:.3e} ({s['worst_name']})"
    for n, v in ours.items():
        if n.endswith("W_key.bias"):      # analytically zero; the Z = Xq M + u algebra never forms it
            assert np.abs(v).max() == 0.0, n


@pytest.mark.parametrize("flags", [(True, False, False), (False, False, False), (True, False, True)],
                         ids=["rca", "ca", "rca_cross_only"])
@pytest.mark.parametrize("qk_gain", [1.0, 40.0])
@pytest.mark.parametrize("drop_p", [0.0, 0.6])
def test_fused_logits_loss_and_gradients(pkg, flags, qk_gain, drop_p):
    """autograd path (mmrca_head_forward + mmrca_cross_entropy + mmrca_head_backward), seeded dropout included:
    the oracle gets the mask the kernels regenerate on chip."""
    from garbage_classification_rca_b200 import _native as N
    from garbage_classification_rca_b200 import functional as F
    from garbage_classification_rca_b200.training import CrossEntropyLoss
    rev, fo, co = flags
    B, seed = 512, 77
    p = orc.init_head_params(features_only=fo, cross_attention_only=co, seed=31, qk_gain=qk_gain)
    img, txt, labels = make_inputs(B, 31)
    mask, scale = None, 1.0
    if drop_p > 0:
        mask = F.dropout_mask(seed, drop_p, B, F.concat_width(1280, 768, fo, co), "cuda").cpu().numpy()
        scale = 1.0 / (1.0 - drop_p)
        assert abs(mask.mean() - (1.0 - drop_p)) < 0.01
    ref = orc.np_head_forward_backward(p, img.numpy(), txt.numpy(), rev, fo, co, labels=labels.numpy(),
                                       drop_mask=mask, drop_scale=scale)
    names = pkg.head_param_names(fo, co)
    params = [p[n].cuda().requires_grad_(True) for n in names]
    N.kernel_launches(reset=True)
    logits = pkg.mmrca_head(img.cuda(), txt.cuda(), params, reverse=rev, features_only=fo, cross_attention_only=co,
                            compute=N.COMPUTE_BF16, drop_p=drop_p, drop_seed=seed)
    loss = CrossEntropyLoss()(logits, labels.cuda())
    loss.backward()
    torch.cuda.synchronize()
    assert N.kernel_launches() >= 3      # forward launches of this thread (autograd runs the backward on its own)
    lg = logits.detach().cpu().numpy()
    assert np.abs(lg - ref["logits"]).max() < LOGITS_ABS_BF16 * max(1.0, scale)
    assert abs(loss.item() - ref["loss"]) < 5e-3
    check_bf16_grads({n: t.grad.cpu().numpy() for n, t in zip(names, params)}, ref["grads"],
                     f"{flags} {qk_gain} {drop_p}", cross_only=co)


@pytest.mark.parametrize("B", [1, 7, 64, 333])
def test_fused_train_step_weighted_smoothed_dropout(pkg, B):
    """One-call step (mmrca_head_train_step): class-weighted, label-smoothed CE (main_both.py:87-93) + dropout,
    gradients accumulated over two calls like loss.backward() does."""
    from garbage_classification_rca_b200 import _native as N
    from garbage_classification_rca_b200 import functional as F
    p = orc.init_head_params(seed=5, qk_gain=40.0)
    img, txt, labels = make_inputs(B, 500 + B)
    cw = torch.tensor([0.6, 1.7, 1.0, 0.9])
    seed, drop_p = 99, 0.6
    mask = F.dropout_mask(seed, drop_p, B, 3584, "cuda").cpu().numpy()
    ref = orc.np_head_forward_backward(p, img.numpy(), txt.numpy(), True, False, False, labels=labels.numpy(),
                                       class_weight=cw.numpy(), label_smoothing=0.1, drop_mask=mask,
                                       drop_scale=1.0 / (1.0 - drop_p))
    names = pkg.head_param_names()
    step = pkg.HeadTrainStep([p[n].cuda() for n in names], B, 1280, 768, reverse=True, class_weight=cw.cuda(),
                             label_smoothing=0.1, compute=N.COMPUTE_BF16, drop_p=drop_p)
    step.zero_grad()
    for _ in range(2):
        loss, logits = step(img.cuda(), txt.cuda(), labels.cuda(), drop_seed=seed)
    torch.cuda.synchronize()
    assert np.abs(logits.cpu().numpy() - ref["logits"]).max() < LOGITS_ABS_BF16 * 2.5
    assert abs(loss.item() - ref["loss"]) < 5e-3
    if B >= 64:
        ours = {n: v.cpu().numpy() / 2.0 for n, v in zip(names, step.grads.views)}
        check_bf16_grads(ours, ref["grads"], f"train step B={B}", cross_only=B < 200)   # 64 samples: the looser row
    # the feature-source rows of the classifier gradient (ce_feat_kernel): fp32 products of the bf16-rounded normalised
    # features (2^-9 relative per element)
    gw = step.grads.views[names.index("final_with_everything.weight")].cpu().numpy() / 2.0
    rw = ref["grads"]["final_with_everything.weight"]
    assert np.abs(gw[:, 1536:] - rw[:, 1536:]).max() <= 6e-3 * np.abs(rw[:, 1536:]).max() + 1e-7


def test_dropout_mask_is_a_pure_function_of_the_seed(pkg):
    from garbage_classification_rca_b200 import functional as F
    a = F.dropout_mask(7, 0.6, 64, 3584, "cuda")
    b = F.dropout_mask(7, 0.6, 64, 3584, "cuda")
    c = F.dropout_mask(8, 0.6, 64, 3584, "cuda")
    assert torch.equal(a, b) and not torch.equal(a, c)
    assert abs(a.float().mean().item() - 0.4) < 0.01
    assert (a.float().mean(0) - 0.4).abs().max().item() < 0.3        # no dead / always-on columns
    assert F.dropout_mask(7, 0.0, 4, 3584, "cuda").all()
    # a sample's mask does not depend on the batch it sits in (row b is the same in a larger batch)
    assert torch.equal(F.dropout_mask(7, 0.6, 128, 3584, "cuda")[:64], a)


def test_seeded_dropout_fp32_kernels_match_oracle(pkg):
    """Same seed through the fp32 kernels (they read the materialised mask): the tight fp32 tolerances."""
    from garbage_classification_rca_b200 import _native as N
    from garbage_classification_rca_b200 import functional as F
    from garbage_classification_rca_b200.training import CrossEntropyLoss
    from tests._util import GRAD_REL_FP32_TIGHT, LOGITS_REL_FP32, assert_grad_close, rel_err
    B, seed, drop_p = 37, 3, 0.6
    p = orc.init_head_params(seed=9, qk_gain=40.0)
    img, txt, labels = make_inputs(B, 9)
    mask = F.dropout_mask(seed, drop_p, B, 3584, "cuda").cpu().numpy()
    ref = orc.np_head_forward_backward(p, img.numpy(), txt.numpy(), True, False, False, labels=labels.numpy(),
                                       drop_mask=mask, drop_scale=2.5)
    names = pkg.head_param_names()
    params = [p[n].cuda().requires_grad_(True) for n in names]
    logits = pkg.mmrca_head(img.cuda(), txt.cuda(), params, reverse=True, compute=N.COMPUTE_FP32, drop_p=drop_p,
                            drop_seed=seed)
    CrossEntropyLoss()(logits, labels.cuda()).backward()
    assert rel_err(logits.detach().cpu().numpy(), ref["logits"]) < LOGITS_REL_FP32
    scale = max(np.abs(v).max() for v in ref["grads"].values())
    for n, t in zip(names, params):
        assert_grad_close(n, t.grad.cpu().numpy(), ref["grads"][n], GRAD_REL_FP32_TIGHT, scale)


@pytest.mark.parametrize("co", [False, True], ids=["features_only", "features_only+cross_only"])
@pytest.mark.parametrize("B,drop_p", [(1, 0.0), (61, 0.6), (4096, 0.6)])
def test_fused_features_only_train_step(pkg, B, drop_p, co):
    """--features_only through the bf16 build: two streaming kernels (prep_feat: normalise + fp32 classifier terms;
    ce_feat: cross-entropy + dWf from the bf16 feature images).  Logits are fp32 arithmetic: the 1e-4-relative contract;
    dWf sees bf16-rounded features: 1e-2 of its largest entry.  The attention blocks stay outside the graph (reference
    multimodal_model.py:694-699): their gradients are untouched."""
    from garbage_classification_rca_b200 import _native as N
    from garbage_classification_rca_b200 import functional as F
    # with both switches set the reference's if / elif gives features_only precedence (multimodal_model.py:694-726)
    p = orc.init_head_params(features_only=True, cross_attention_only=co, seed=9)
    img, txt, labels = make_inputs(B, 9)
    names = pkg.head_param_names(True, co)
    cw = torch.tensor([0.8, 1.3, 1.0, 0.9])
    step = pkg.HeadTrainStep([p[n].cuda() for n in names], B, 1280, 768, reverse=True, features_only=True,
                             cross_attention_only=co, class_weight=cw.cuda(), label_smoothing=0.1,
                             compute=N.COMPUTE_BF16, drop_p=drop_p)
    step.zero_grad()
    N.kernel_launches(reset=True)
    loss, logits = step(img.cuda(), txt.cuda(), labels.cuda(), drop_seed=5)
    torch.cuda.synchronize()
    assert N.kernel_launches() == 2
    mask, scale = None, 1.0
    if drop_p > 0:
        mask = F.dropout_mask(5, drop_p, B, 2048, "cuda").cpu().numpy()
        scale = 1.0 / (1.0 - drop_p)
    ref = orc.np_head_forward_backward(p, img.numpy(), txt.numpy(), True, True, co, labels=labels.numpy(),
                                       class_weight=cw.numpy(), label_smoothing=0.1, drop_mask=mask, drop_scale=scale)
    lg = logits.cpu().numpy()
    assert np.abs(lg - ref["logits"]).max() <= 1e-4 * np.abs(ref["logits"]).max()
    assert abs(loss.item() - ref["loss"]) < 1e-4
    for n, v in zip(names, step.grads.views):
        g = v.cpu().numpy()
        if n.startswith("final_features_only_linear"):
            r = ref["grads"][n]
            assert np.abs(g - r).max() <= 1e-2 * np.abs(r).max() + 1e-9, n
        else:
            assert np.abs(g).max() == 0.0, n


def test_fused_gradients_at_benchmark_batch(pkg):
    """BASELINE.json configs[1] itself: batch 4096, --reverse, dropout 0.6 - the whole head gradient against the float64
    oracle (same keep mask).  The norm-wise error shrinks with the batch (the systematic weight-rounding term is gone
    since the hi + lo W_value / classifier blobs): held to 3e-3, a third of north_star's 1e-2."""
    from garbage_classification_rca_b200 import _native as N
    from garbage_classification_rca_b200 import functional as F
    B, seed, drop_p = 4096, 4242, 0.6
    p = orc.init_head_params(seed=12, qk_gain=40.0)
    img, txt, labels = make_inputs(B, 12)
    mask = F.dropout_mask(seed, drop_p, B, 3584, "cuda").cpu().numpy()
    ref = orc.np_head_forward_backward(p, img.numpy(), txt.numpy(), True, False, False, labels=labels.numpy(),
                                       drop_mask=mask, drop_scale=1.0 / (1.0 - drop_p))
    names = pkg.head_param_names()
    step = pkg.HeadTrainStep([p[n].cuda() for n in names], B, 1280, 768, reverse=True, compute=N.COMPUTE_BF16, drop_p=drop_p)
    step.zero_grad()
    loss, logits = step(img.cuda(), txt.cuda(), labels.cuda(), drop_seed=seed)
    torch.cuda.synchronize()
    assert np.abs(logits.cpu().numpy() - ref["logits"]).max() < LOGITS_ABS_BF16 * 2.5
    s = grad_summary({n: v.cpu().numpy() for n, v in zip(names, step.grads.views)}, ref["grads"])
    assert s["flat_l2_rel"] <= 3e-3 and s["cos"] >= 0.99999, s
    assert s["worst_l2_rel"] <= 3e-2 and s["max_err_over_global"] <= 5e-3, s


@pytest.mark.parametrize("B", [8, 333])
def test_fused_train_step_bf16_features(pkg, B):
    """MMRCA_FLAG_FEATURES_BF16: the features arrive as bf16 (a backbone under bf16 autocast / a host hand-off shipping
    half the bytes).  Against the oracle on the SAME (bf16-rounded) features the usual limits hold; against the oracle on
    the original fp32 features the logits stay within north_star's 2e-2 absolute."""
    from garbage_classification_rca_b200 import _native as N
    p = orc.init_head_params(seed=21, qk_gain=40.0)
    img, txt, labels = make_inputs(B, 600 + B)
    img16, txt16 = img.bfloat16(), txt.bfloat16()
    ref32 = orc.np_head_forward_backward(p, img.numpy(), txt.numpy(), True, False, False, labels=labels.numpy())
    ref16 = orc.np_head_forward_backward(p, img16.float().numpy(), txt16.float().numpy(), True, False, False,
                                         labels=labels.numpy())
    names = pkg.head_param_names()
    step = pkg.HeadTrainStep([p[n].cuda() for n in names], B, 1280, 768, reverse=True, compute=N.COMPUTE_BF16)
    step.zero_grad()
    loss, logits = step(img16.cuda(), txt16.cuda(), labels.cuda())
    torch.cuda.synchronize()
    lg = logits.cpu().numpy()
    assert np.abs(lg - ref32["logits"]).max() < LOGITS_ABS_BF16
    assert np.abs(lg - ref16["logits"]).max() < LOGITS_ABS_BF16
    assert abs(loss.item() - ref16["loss"]) < 5e-3
    if B >= 200:
        check_bf16_grads({n: v.cpu().numpy() for n, v in zip(names, step.grads.views)}, ref16["grads"], "bf16 features")
    # the same step object takes fp32 features again (the flag is per call)
    step.zero_grad()
    _, logits32 = step(img.cuda(), txt.cuda(), labels.cuda())
    assert np.abs(logits32.cpu().numpy() - ref32["logits"]).max() < LOGITS_ABS_BF16


def test_zero_grad_folded_into_the_step(pkg):
    """MMRCA_FLAG_ZERO_GRADS: the step clears the contiguous gradient bucket in its first kernel - the same gradients as
    zero_grad() followed by the step, and no accumulation across calls."""
    from garbage_classification_rca_b200 import _native as N
    B = 40
    p = orc.init_head_params(seed=3, qk_gain=40.0)
    img, txt, labels = make_inputs(B, 3)
    names = pkg.head_param_names()
    step = pkg.HeadTrainStep([p[n].cuda() for n in names], B, 1280, 768, reverse=True, compute=N.COMPUTE_BF16, drop_p=0.6)
    step.zero_grad()
    step(img.cuda(), txt.cuda(), labels.cuda(), drop_seed=4)
    ref = step.grads.flat.clone()
    step.grads.flat.fill_(123.0)
    for _ in range(2):
        step(img.cuda(), txt.cuda(), labels.cuda(), drop_seed=4, zero_grad=True)
    torch.cuda.synchronize()
    n = step.grads.n
    assert (step.grads.flat[:n] - ref[:n]).abs().max().item() <= 1e-6 * max(1.0, ref[:n].abs().max().item())
    # the fp32 kernels take the flag too (a memset node instead of the fused clear)
    step32 = pkg.HeadTrainStep([p[n].cuda() for n in names], B, 1280, 768, reverse=True, compute=N.COMPUTE_FP32)
    step32.zero_grad()
    step32(img.cuda(), txt.cuda(), labels.cuda())
    ref32 = step32.grads.flat.clone()
    step32.grads.flat.fill_(-7.0)
    step32(img.cuda(), txt.cuda(), labels.cuda(), zero_grad=True)
    assert (step32.grads.flat[:n] - ref32[:n]).abs().max().item() <= 1e-6 * max(1.0, ref32[:n].abs().max().item())


@pytest.mark.parametrize("flags", [(True, False, False), (False, False, False), (True, False, True)],
                         ids=["rca", "ca", "rca_cross_only"])
@pytest.mark.parametrize("B,drop_p", [(8, 0.0), (61, 0.6), (512, 0.6)])
def test_fused_feature_gradients(pkg, flags, B, drop_p):
    """Fine-tune phase (reference main_both.py:687-694): MMRCA_FLAG_FEATURE_GRADS inside the bf16 pipeline (sa_bwd exports
    dZ / dS / dV, sa_dx_kernel finishes d(loss)/d(features)), autograd path and one-call step, against the float64 oracle.
    Relative L2 error of the whole [B, d] gradient: measured 0.4 - 1.8 % (image) and 2.3 - 5.1 % (text: its 48-wide chunks
    and the x40 query / key gain of this test put most of its gradient on the bf16 attention path rather than on the
    fp32 classifier term); held to 3e-2 / 6e-2 and a cosine of 0.998.  The fp32 kernels give 1e-4 (test_parity_gpu.py)."""
    from garbage_classification_rca_b200 import _native as N
    from garbage_classification_rca_b200 import functional as F
    from garbage_classification_rca_b200.training import CrossEntropyLoss
    rev, fo, co = flags
    seed = 19
    p = orc.init_head_params(features_only=fo, cross_attention_only=co, seed=8, qk_gain=40.0)
    img, txt, labels = make_inputs(B, 800 + B)
    mask, scale = None, 1.0
    if drop_p > 0:
        mask = F.dropout_mask(seed, drop_p, B, F.concat_width(1280, 768, fo, co), "cuda").cpu().numpy()
        scale = 1.0 / (1.0 - drop_p)
    ref = orc.np_head_forward_backward(p, img.numpy(), txt.numpy(), rev, fo, co, labels=labels.numpy(), drop_mask=mask,
                                       drop_scale=scale)
    names = pkg.head_param_names(fo, co)

    def rel(a, b):
        return float(np.linalg.norm(a - b) / np.linalg.norm(b))

    # one-call step
    step = pkg.HeadTrainStep([p[n].cuda() for n in names], B, 1280, 768, reverse=rev, cross_attention_only=co,
                             compute=N.COMPUTE_BF16, drop_p=drop_p, feature_grads=True)
    step.zero_grad()
    N.kernel_launches(reset=True)
    step(img.cuda(), txt.cuda(), labels.cuda(), drop_seed=seed)
    torch.cuda.synchronize()
    assert N.kernel_launches() == 8                      # the 7 kernels of the step + sa_dx: still the tensor-core pipeline
    e_img, e_txt = rel(step.d_img.cpu().numpy(), ref["d_img"]), rel(step.d_txt.cpu().numpy(), ref["d_txt"])
    assert e_img <= 3e-2 and e_txt <= 6e-2, (e_img, e_txt)
    for ours, r in ((step.d_img, ref["d_img"]), (step.d_txt, ref["d_txt"])):
        o = ours.cpu().numpy().ravel().astype(np.float64)
        assert o @ r.ravel() / (np.linalg.norm(o) * np.linalg.norm(r)) >= 0.998
    if B >= 200:
        check_bf16_grads({n: v.cpu().numpy() for n, v in zip(names, step.grads.views)}, ref["grads"], "feature-grad step",
                         cross_only=co)
    # autograd path: features that require grad
    params = [p[n].cuda().requires_grad_(True) for n in names]
    xi, xt = img.cuda().requires_grad_(True), txt.cuda().requires_grad_(True)
    logits = pkg.mmrca_head(xi, xt, params, reverse=rev, cross_attention_only=co, compute=N.COMPUTE_BF16, drop_p=drop_p,
                            drop_seed=seed)
    CrossEntropyLoss()(logits, labels.cuda()).backward()
    assert rel(xi.grad.cpu().numpy(), ref["d_img"]) <= 3e-2 and rel(xt.grad.cpu().numpy(), ref["d_txt"]) <= 6e-2
    # both routes run the same kernels (the logits' red.global order may differ in the last bit)
    assert rel(xi.grad.cpu().numpy(), step.d_img.cpu().numpy()) < 1e-3 and rel(xt.grad.cpu().numpy(), step.d_txt.cpu().numpy()) < 1e-3
