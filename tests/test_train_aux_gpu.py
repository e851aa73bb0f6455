"""GPU: the pieces around the head through the C ABI — fused feature hand-off, the flat
gradient / parameter buckets of the module path, the fused optimizer steps, the frozen-phase feature cache and the
image_only / text_only evaluation modes.  Checkers: torch ops on the same inputs and the oracle."""
import io
from contextlib import redirect_stdout

import numpy as np
import pytest
import torch

from oracle import mmrca_oracle as orc
from tests._util import LOGITS_REL_FP32, make_inputs, rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pkg(native_lib):
    import garbage_classification_rca_b200 as g
    return g


@pytest.mark.parametrize("in_dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("out_dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("channels_last", [False, True])
def test_feature_handoff_matches_torch(pkg, in_dtype, out_dtype, channels_last):
    """CLS gather + global average pool + cast in one kernel vs hidden[:, 0] and avgpool + flatten (reference
    multimodal_model.py:651-658, :25-36); ragged batch, EfficientNetV2-M's 15 x 15 final map and an odd 7 x 5 one."""
    from garbage_classification_rca_b200 import functional as F
    g = torch.Generator().manual_seed(3)
    for B, T, (h, w) in ((5, 9, (15, 15)), (1, 3, (7, 5))):
        hidden = torch.randn(B, T, 768, generator=g).to(in_dtype).cuda()
        fmap = torch.randn(B, 1280, h, w, generator=g).to(in_dtype).cuda()
        if channels_last:
            fmap = fmap.contiguous(memory_format=torch.channels_last)
        img, txt = F.feature_handoff(hidden, fmap, out_dtype)
        ref_txt = hidden[:, 0].float()
        ref_img = fmap.float().mean(dim=(2, 3))
        tol = 1e-6 if out_dtype == torch.float32 else 2 ** -8
        assert img.dtype == out_dtype and txt.dtype == out_dtype
        assert (txt.float() - ref_txt).abs().max().item() <= tol * max(1.0, ref_txt.abs().max().item())
        assert (img.float() - ref_img).abs().max().item() <= tol * max(1.0, ref_img.abs().max().item()) + 1e-6
    # a strided view of a longer hidden state (what backbone(...)[0] is when sliced) is accepted as it is
    hidden = torch.randn(4, 6, 768, generator=g).cuda()
    img, txt = F.feature_handoff(hidden[:, 1:], torch.ones(4, 1280, 2, 2).cuda(), torch.float32)
    assert torch.equal(txt, hidden[:, 1]) and torch.equal(img, torch.ones(4, 1280).cuda())


def _buckets(n, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(n, generator=g).cuda(), [torch.randn(n, generator=g).cuda() for _ in range(5)]


@pytest.mark.parametrize("cfg", [dict(lr=0.0016, weight_decay=0.03), dict(lr=0.01, momentum=0.9, weight_decay=1e-3),
                                 dict(lr=0.01, momentum=0.8, nesterov=True), dict(lr=0.05, momentum=0.9, dampening=0.1)],
                         ids=["reference_sgd", "momentum", "nesterov", "dampening"])
def test_fused_sgd_matches_torch(pkg, cfg):
    """One-launch SGD over the flat bucket vs torch.optim.SGD (reference main_both.py:548-549: lr, weight_decay = --reg)."""
    from garbage_classification_rca_b200 import functional as F, training as T
    n = 94824
    p0, grads = _buckets(n, 1)
    ref_p = torch.nn.Parameter(p0.clone())
    opt = torch.optim.SGD([ref_p], **cfg)
    params = [torch.nn.Parameter(p0[:94820].clone()), torch.nn.Parameter(p0[94820:94824].clone())]
    fg = F.FlatGrads([q.detach() for q in params])
    fp = T.FlatParams(params, fg)
    fused = T.FusedSGD(fp, fg, **cfg)
    for g_ in grads:
        ref_p.grad = g_.clone()
        opt.step()
        fg.flat[:n].copy_(g_)
        fused.step()
    assert (fp.flat[:n] - ref_p.detach()).abs().max().item() < 1e-6
    assert params[0].data_ptr() == fp.flat.data_ptr()      # the parameters live in the bucket


def test_fused_adamw_matches_torch(pkg):
    from garbage_classification_rca_b200 import functional as F, training as T
    n = 4096
    p0, grads = _buckets(n, 2)
    ref_p = torch.nn.Parameter(p0.clone())
    opt = torch.optim.AdamW([ref_p], lr=1e-3, weight_decay=0.03)
    params = [torch.nn.Parameter(p0.clone())]
    fg = F.FlatGrads([q.detach() for q in params])
    fp = T.FlatParams(params, fg)
    fused = T.FusedAdamW(fp, fg, lr=1e-3, weight_decay=0.03)
    for g_ in grads:
        ref_p.grad = g_.clone()
        opt.step()
        fg.flat[:n].copy_(g_)
        fused.step()
    assert (fp.flat[:n] - ref_p.detach()).abs().max().item() < 2e-6


class _StubText(torch.nn.Module):
    calls = 0

    def forward(self, input_ids=None, attention_mask=None, **kw):
        type(self).calls += 1
        return (self.feat[input_ids[:, 0]].unsqueeze(1),)      # "text" of sample i is row ids[i, 0] of the table


class _StubImage(torch.nn.Module):
    calls = 0

    def forward(self, x):
        type(self).calls += 1
        return None, None, self.feat[x[:, 0, 0, 0].long()]


def _stub_model(compute=None, drop=0.0):
    from garbage_classification_rca_b200 import _native as N, multimodal_model as M
    torch.manual_seed(0)
    with redirect_stdout(io.StringIO()):
        m = M.MM_RCA(4, drop, 0.0, 0.7, 256, "distilbert", 16, True, False, False, pretrained=False,
                     compute=N.COMPUTE_FP32 if compute is None else compute)
    m.text_model, m.image_model = _StubText(), _StubImage()
    return m.cuda()


def test_module_flat_grads_and_fused_step(pkg):
    """attach_flat_grads: the backward kernels accumulate straight into ONE persistent bucket whose views are the
    parameters' .grad (no per-backward allocation), equal to the autograd-returned gradients; with flat_params a fused SGD
    step equals torch.optim.SGD on an identical model."""
    from garbage_classification_rca_b200 import training as T
    B = 12
    img, txt, labels = make_inputs(B, 5)
    ids = torch.arange(B).view(B, 1).cuda()
    images = torch.arange(B, dtype=torch.float32).view(B, 1, 1, 1).expand(B, 3, 2, 2).contiguous().cuda()
    models = [_stub_model(), _stub_model()]
    for m in models:
        m.text_model.feat, m.image_model.feat = txt.cuda(), img.cuda()
        m.train()
    plain, flat = models
    fg = flat.attach_flat_grads(flat_params=True)
    crit = T.CrossEntropyLoss()
    opt_ref = torch.optim.SGD([p for p in plain.head_parameters()], lr=0.05, weight_decay=0.03)
    opt_fused = T.FusedSGD(flat._flat_params, fg, lr=0.05, weight_decay=0.03)
    for it in range(3):
        for m in models:
            crit(m(_input_ids=ids, _attention_mask=torch.ones_like(ids), _images=images), labels.cuda()).backward()
        for p, q in zip(plain.head_parameters(), flat.head_parameters()):
            assert q.grad.data_ptr() >= fg.flat.data_ptr() and q.grad.data_ptr() < fg.flat.data_ptr() + 4 * fg.flat.numel()
            assert (p.grad - q.grad).abs().max().item() <= 1e-6 * max(1.0, p.grad.abs().max().item())
        opt_ref.step(); opt_ref.zero_grad(set_to_none=False)
        opt_fused.step(); opt_fused.zero_grad()
        for p, q in zip(plain.head_parameters(), flat.head_parameters()):
            assert (p.detach() - q.detach()).abs().max().item() < 1e-6
    # state_dict is unaffected by the flat layout: strict round trip into a fresh module
    fresh = _stub_model()
    fresh.load_state_dict(flat.state_dict(), strict=True)


def test_feature_cache_skips_the_frozen_backbones(pkg):
    """Frozen phase: the second pass over the same sample ids does not run the backbones (the reference recomputes them
    every step, main_both.py:562-577) and gives the same logits (fp32 cache: bit-identical)."""
    B = 6
    img, txt, _ = make_inputs(B, 9)
    m = _stub_model()
    m.text_model.feat, m.image_model.feat = txt.cuda(), img.cuda()
    m.eval()
    cache = m.enable_feature_cache(64, dtype=torch.float32)
    ids = torch.arange(B).view(B, 1).cuda()
    images = torch.arange(B, dtype=torch.float32).view(B, 1, 1, 1).expand(B, 3, 2, 2).contiguous().cuda()
    sample_ids = torch.tensor([40, 3, 17, 5, 63, 0])
    _StubText.calls = _StubImage.calls = 0
    with torch.no_grad():
        a = m(ids, torch.ones_like(ids), images, sample_ids=sample_ids)
        assert (_StubText.calls, _StubImage.calls) == (1, 1)
        b = m(ids, torch.ones_like(ids), images, sample_ids=sample_ids)
        assert (_StubText.calls, _StubImage.calls) == (1, 1) and cache.hits == B
        c = m(ids, torch.ones_like(ids), images)                  # no ids: the stock path
        assert (_StubText.calls, _StubImage.calls) == (2, 2)
        perm = torch.tensor([5, 4, 3, 2, 1, 0])
        d = m(ids, torch.ones_like(ids), images, sample_ids=sample_ids[perm])      # cached rows in another order
    assert torch.equal(a, b) and torch.equal(a, c) and torch.equal(d, a[perm])
    ref = orc.head_forward({k: v.detach().cpu() for k, v in m.state_dict().items()}, img, txt, True)
    assert rel_err(a.cpu().numpy(), ref.numpy()) < LOGITS_REL_FP32


def test_calculate_set_accuracy_modes(pkg):
    """image_only / text_only / both (reference main_both.py:43-47, :141-198): the mode's remove_* flags reach
    drop_modalities, which zeroes the other modality's inputs before the backbones."""
    from garbage_classification_rca_b200 import training as T
    B = 8
    img, txt, labels = make_inputs(B + 1, 21)      # row 0: what a zeroed input selects (ids / images all zero)
    m = _stub_model()
    m.text_model.feat, m.image_model.feat = txt.cuda(), img.cuda()
    m.eval()
    ids = torch.arange(1, B + 1).view(B, 1)
    images = torch.arange(1, B + 1, dtype=torch.float32).view(B, 1, 1, 1).expand(B, 3, 2, 2).contiguous()
    loader = [({"text": {"tokens": ids, "attention_mask": torch.ones_like(ids)}, "image": {"raw_image": images}},
               labels[1:])]
    sd = {k: v.detach().cpu() for k, v in m.state_dict().items()}
    for mode in ("both", "image_only", "text_only"):
        with redirect_stdout(io.StringIO()):
            acc, report = T.calculate_set_accuracy(m, loader, B, "cuda", B, T.mode_config_dict[mode], True)
        ti = txt[1:] if mode != "image_only" else txt[:1].expand(B, -1)       # zeroed ids pick row 0
        ii = img[1:] if mode != "text_only" else img[:1].expand(B, -1)
        ref = orc.head_forward(sd, ii, ti, True).argmax(1)
        assert abs(acc - 100.0 * (ref == labels[1:]).float().mean().item()) < 1e-4
        assert abs(report["accuracy"] * 100.0 - acc) < 1e-4 and sum(r["support"] for k, r in report.items() if k != "accuracy") == B


def test_module_bf16_features_path(pkg):
    """bf16 pooled features (a backbone under bf16 autocast) go to the bf16 pipeline without an fp32 copy."""
    from garbage_classification_rca_b200 import _native as N
    B = 40
    img, txt, _ = make_inputs(B, 33)
    m = _stub_model(compute=N.COMPUTE_BF16)
    m.eval()
    with torch.no_grad():
        N.kernel_launches(reset=True)
        out = m.forward_features(img.bfloat16().cuda(), txt.bfloat16().cuda())
        assert N.kernel_launches() == 3
    sd = {k: v.detach().cpu() for k, v in m.state_dict().items()}
    ref = orc.head_forward(sd, img, txt, True)
    assert (out.cpu() - ref).abs().max().item() < 2e-2
