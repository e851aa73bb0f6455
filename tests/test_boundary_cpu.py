"""CPU (no GPU): the C-ABI library builds, loads and exports every symbol include/mmrca.h declares; the
Python mirror reproduces the reference's constructor / state_dict layout / flag surface; the product
path refuses to run without its CUDA device instead of falling back."""
import hashlib
import io
import json
import os
import re
from contextlib import redirect_stdout

import pytest
import torch

from tests._util import GOLDEN

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(native_lib):
    from garbage_classification_rca_b200 import _native
    hdr = open(os.path.join(ROOT, "include", "mmrca.h")).read()
    declared = set(re.findall(r"\b(mmrca_[a-z_0-9]+)\s*\(", hdr))
    assert declared == set(_native.EXPORTS)
    for sym in declared:
        assert hasattr(native_lib, sym), sym
    assert native_lib.mmrca_query(_native.QUERY_ABI_VERSION) == _native.ABI_VERSION


def test_workspace_size_is_monotone(native_lib):
    from garbage_classification_rca_b200 import functional as F
    prev = 0
    for b in (0, 1, 7, 256, 4096):
        inf = F.workspace_bytes(b, 1280, 768, 4, 1, 0, training=False)
        trn = F.workspace_bytes(b, 1280, 768, 4, 1, 0, training=True)
        assert trn >= inf >= prev
        prev = inf


def test_token_and_fusion_descriptors_are_validated_on_the_host(native_lib):
    """Host-side logic of the token-level / classic-head entry points that needs no GPU: workspace sizing (the training and
    bf16 flags grow it), shape validation, and the backward's refusal to run without the training flag."""
    import ctypes as C
    from garbage_classification_rca_b200 import _native as N
    ws = lambda d: int(native_lib.mmrca_token_attention_workspace_bytes(C.byref(d)))
    inf = N.TokenDesc(8, 197, 1024, 1024, 128, 96, 0, 0)
    trn = N.TokenDesc(8, 197, 1024, 1024, 128, 96, 0, N.TOKEN_TRAINING)
    assert 0 < ws(inf) < ws(trn)
    assert ws(N.TokenDesc(16, 197, 1024, 1024, 128, 96, 0, N.TOKEN_TRAINING)) > ws(trn)
    for bad in (N.TokenDesc(8, 257, 1024, 1024, 128, 96, 0, 0),       # L > 256
                N.TokenDesc(8, 1, 1024, 1024, 128, 96, 0, 0),         # L < 2
                N.TokenDesc(8, 197, 1024, 1024, 100, 96, 0, 0),       # widths outside {(128, 96), (64, 48)}
                N.TokenDesc(8, 197, 1020, 1020, 128, 96, 0, 0)):      # d_in not a multiple of 8
        assert ws(bad) == 0 and N.last_error()
    # the backward needs MMRCA_TOKEN_TRAINING on the descriptor: refused before any device is touched
    rc = native_lib.mmrca_token_attention_backward(C.byref(inf), None, None, None, None, None, None, None, None, 0, None)
    assert rc == 1 and "MMRCA_TOKEN_TRAINING" in N.last_error()      # MMRCA_ERR_INVALID
    fws = lambda d: int(native_lib.mmrca_fusion_workspace_bytes(C.byref(d)))
    f32 = N.FusionDesc(64, 1280, 768, 256, 4, 0, 0.0, 0)
    b16 = N.FusionDesc(64, 1280, 768, 256, 4, N.FUSION_BF16, 0.0, 0)
    assert 0 < fws(f32) < fws(b16)
    assert fws(N.FusionDesc(64, 1280, 768, 40, 4, 0, 0.0, 0)) > 0                       # fp32: any hidden width % 4
    assert fws(N.FusionDesc(64, 1280, 768, 40, 4, N.FUSION_BF16, 0.0, 0)) == 0          # bf16: hidden % 16
    assert fws(N.FusionDesc(64, 1280, 768, 512, 4, N.FUSION_BF16, 0.0, 0)) == 0         # bf16: hidden <= 256


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback(native_lib):
    import garbage_classification_rca_b200 as g
    assert native_lib.mmrca_query(g._native.QUERY_DEVICE_OK) == 0
    params = g.functional.init_head_parameters("cpu")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        g.mmrca_head(torch.randn(2, 1280), torch.randn(2, 768), params, reverse=True)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        g.cross_entropy(torch.randn(2, 4), torch.zeros(2, dtype=torch.long))


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "garbage_classification_rca_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src, f"{f} mentions the oracle"


@pytest.fixture(scope="module")
def layout():
    with open(os.path.join(GOLDEN, "state_dict_layout.json")) as fh:
        return json.load(fh)


@pytest.mark.parametrize("flags", [(False, False), (True, False), (False, True)])
def test_state_dict_layout_matches_reference(layout, flags):
    """Names, shapes, dtypes and ORDER of the reference's state_dict (SURVEY.md §5 checkpoint row): 88 head
    tensors incl. the dead gated/CLIP/GRU ones, image_model.{stem,stage1..}, conditional classifier keys."""
    from garbage_classification_rca_b200 import multimodal_model as M
    fo, co = flags
    with redirect_stdout(io.StringIO()) as out:
        m = M.MM_RCA(4, 0.6, 0.0, 0.7, 256, "distilbert", 16, True, fo, co, pretrained=False)
    assert "txt patch size:  48" in out.getvalue() and "img patch size:  80" in out.getvalue()
    ref = layout[f"distilbert|features_only={int(fo)}|cross_attention_only={int(co)}"]
    ours = [[k, list(v.shape), str(v.dtype)] for k, v in m.state_dict().items()]
    assert len(ours) == ref["order_total"]
    head = [e for e in ours if not e[0].startswith(("image_model.", "text_model."))]
    assert head == ref["head"]
    assert [e[0] for e in ours].index(head[0][0]) == ref["head_first_index"]
    for pre, dg in ref["backbones"].items():
        sub = [e for e in ours if e[0].startswith(pre)]
        h = hashlib.sha256("\n".join(f"{k}:{s}:{d}" for k, s, d in sub).encode()).hexdigest()
        assert (len(sub), h, sub[0][0], sub[-1][0]) == (dg["count"], dg["sha256"], dg["first"], dg["last"])
    # the 34 tensors the forward reads, frozen backbones, helper methods
    assert len(m.head_parameters()) == 34
    assert sum(p.numel() for p in m.head_parameters()) == {(False, False): 94820, (True, False): 88676,
                                                          (False, True): 86628}[flags]
    assert not any(p.requires_grad for p in m.text_model.parameters())
    assert m.get_image_size() == (480, 480) and m.get_max_token_size() == 512


def test_reference_call_sites_with_fewer_args_construct():
    """calculate_test_accuracy_both.py:162-171 passes 9 positional args, main_both.py:319-327 passes 8
    (TypeError in the reference): the trailing switches default here."""
    from garbage_classification_rca_b200 import multimodal_model as M
    with redirect_stdout(io.StringIO()):
        m = M.MM_RCA(4, 0.6, 0.0, 0.7, 256, "distilbert", 16, True, False, pretrained=False)
    assert m.cross_attention_only is False and m.reverse is True
    with pytest.raises(SystemExit):
        with redirect_stdout(io.StringIO()):
            M.MM_RCA(4, 0.6, 0.0, 0.7, 256, "roberta", 16, True, False, False, pretrained=False)


def test_option_surface():
    from garbage_classification_rca_b200.options import args_parser
    a = args_parser(["--late_fusion=MM_RCA", "--reverse", "--features-only"])
    assert (a.late_fusion, a.reverse, a.features_only, a.cross_attention_only) == ("MM_RCA", True, True, False)
    a = args_parser(["--no-reverse", "--cross_attention_only", "--model_dropout", "0.3"])
    assert (a.reverse, a.cross_attention_only, a.model_dropout) == (False, True, 0.3)
    d = args_parser([])
    assert (d.late_fusion, d.model_dropout, d.batch_size, d.num_neurons_FC, d.label_smoothing) == \
        ("gated", 0.6, 16, 256, 0.0)


def test_drop_modalities_matches_reference_semantics():
    """eval=True + remove_* zero the image batch / ids+mask (reference :424-435); training with
    image_or_text_dropout_chance = 0 never drops (:444-455)."""
    from garbage_classification_rca_b200 import multimodal_model as M
    m = M.MM_RCA.__new__(M.MM_RCA)
    m.image_or_text_dropout_chance, m.img_dropout_prob = 0.0, 0.7
    ids = torch.arange(12).reshape(2, 6)
    m._images, m._input_ids, m._attention_mask = torch.ones(2, 3, 4, 4), ids.clone(), torch.ones_like(ids)
    with redirect_stdout(io.StringIO()):
        m.drop_modalities(False, False, False)
        assert m._images.sum() > 0 and torch.equal(m._input_ids, ids)
        m.drop_modalities(True, True, False)
        assert m._images.sum() == 0 and torch.equal(m._input_ids, ids)
        m.drop_modalities(True, False, True)
        assert m._input_ids.sum() == 0 and m._attention_mask.sum() == 0


REFERENCE = "/root/reference"


@pytest.mark.skipif(not os.path.isdir(REFERENCE), reason="the reference checkout only exists in the build container")
@pytest.mark.parametrize("flags", [(True, False, False), (True, True, False), (False, False, True)])
def test_pth_round_trip_with_the_reference_class(tmp_path, flags):
    """save_model_weights (ours) -> torch.load -> STRICT load_state_dict into the unmodified reference MM_RCA, and a
    reference-saved .pth strict-loaded into ours (main_both.py:201-226, calculate_test_accuracy_both.py:187), with the
    backbones stubbed out on both sides (their key names are pinned by test_state_dict_layout_matches_reference)."""
    import sys
    sys.path.insert(0, REFERENCE)
    import CVPR_code.multimodal_model as mm
    from garbage_classification_rca_b200 import multimodal_model as M
    from garbage_classification_rca_b200.training import save_model_weights
    rev, fo, co = flags

    class _Cfg:
        hidden_size = 768

    class Stub(torch.nn.Module):
        config = _Cfg()

    saved = mm.distilbert, mm.eff_net_v2
    mm.distilbert, mm.eff_net_v2 = (lambda: Stub()), (lambda: Stub())
    try:
        with redirect_stdout(io.StringIO()):
            torch.manual_seed(1)
            ref = mm.MM_RCA(4, 0.6, 0.0, 0.7, 256, "distilbert", 16, rev, fo, co)
            torch.manual_seed(2)
            ours = M.MM_RCA(4, 0.6, 0.0, 0.7, 256, "distilbert", 16, rev, fo, co, pretrained=False)
    finally:
        mm.distilbert, mm.eff_net_v2 = saved
    ours.text_model, ours.image_model = Stub(), Stub()
    path = save_model_weights(ours, str(tmp_path / "ours.pth"), "cpu")
    ref.load_state_dict(torch.load(path), strict=True)                       # calculate_test_accuracy_both.py:187
    for (k, a), (k2, b) in zip(ours.state_dict().items(), ref.state_dict().items()):
        assert k == k2 and torch.equal(a, b)
    torch.manual_seed(3)
    for p in ref.parameters():
        p.data.normal_()
    torch.save(ref.state_dict(), str(tmp_path / "ref.pth"))                  # main_both.py:223-224
    M.load_reference_state_dict(ours, torch.load(str(tmp_path / "ref.pth")), strict=True)
    assert all(torch.equal(a, b) for a, b in zip(ours.state_dict().values(), ref.state_dict().values()))
    # a checkpoint written from an nn.DataParallel wrapper carries a "module." prefix
    M.load_reference_state_dict(ours, {"module." + k: v for k, v in ref.state_dict().items()}, strict=True)


def test_run_one_epoch_reference_semantics():
    """run_one_epoch mirror (main_both.py:81-134) on a CPU stand-in model: the optimizer steps every acc_steps batches
    and on the last one, gradients are NOT scaled by 1/acc_steps (the division happens after backward, :112-114), the
    returned losses are the divided ones."""
    from garbage_classification_rca_b200 import training as T

    class Toy(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.lin = torch.nn.Linear(6, 4)

        def forward(self, _input_ids, _attention_mask, _images):
            return self.lin(_images.flatten(1))

    class CE(torch.nn.Module):          # the kernels need a GPU; the loop's semantics do not
        def __init__(self, weight=None, label_smoothing=0.0):
            super().__init__()
            self.f = torch.nn.CrossEntropyLoss(weight=weight, label_smoothing=label_smoothing)

        def forward(self, a, b):
            return self.f(a, b)

    g = torch.Generator().manual_seed(0)
    batches = [({"text": {"tokens": torch.zeros(3, 2, dtype=torch.long), "attention_mask": torch.ones(3, 2, dtype=torch.long)},
                 "image": {"raw_image": torch.randn(3, 6, generator=g)}}, torch.randint(0, 4, (3,), generator=g))
               for _ in range(5)]
    saved, T.CrossEntropyLoss = T.CrossEntropyLoss, CE
    try:
        torch.manual_seed(0)
        m = Toy()
        ref = Toy()
        ref.load_state_dict(m.state_dict())
        opt = torch.optim.SGD(m.parameters(), lr=0.1)
        n_batches, losses = T.run_one_epoch(0, m, batches, 15, "cpu", 3, opt, [1.0, 2.0, 0.5, 1.0], True, 2, 0.1)
    finally:
        T.CrossEntropyLoss = saved
    # restatement of the reference loop
    ropt = torch.optim.SGD(ref.parameters(), lr=0.1)
    crit = torch.nn.CrossEntropyLoss(weight=torch.tensor([1.0, 2.0, 0.5, 1.0]), label_smoothing=0.1)
    rl = []
    for i, (d, y) in enumerate(batches):
        loss = crit(ref(None, None, d["image"]["raw_image"]), y)
        loss.backward()
        rl.append(loss.detach() / 2)
        if (i + 1) % 2 == 0 or i + 1 == len(batches):
            ropt.step()
            ropt.zero_grad()
    assert n_batches == 5 and len(losses) == 5
    assert all(abs(float(a) - float(b)) < 1e-6 for a, b in zip(losses, rl))
    assert all(torch.allclose(a, b, atol=1e-7) for a, b in zip(m.state_dict().values(), ref.state_dict().values()))


def test_build_model_dispatches_the_rebuilt_late_fusion_variants():
    """--late_fusion dispatch (reference main_both.py:272-343): MM_RCA, hierarchical, classic, normalized are rebuilt;
    the others exit like an unknown strategy would."""
    from garbage_classification_rca_b200 import multimodal_model as M
    from garbage_classification_rca_b200.options import args_parser
    from garbage_classification_rca_b200.training import build_model
    want = {"MM_RCA": M.MM_RCA, "hierarchical": M.Hierarchical, "classic": M.EffV2MediumAndDistilbertClassic,
            "normalized": M.EffV2MediumAndDistilbertNormalized}
    for name, cls in want.items():
        a = args_parser([f"--late_fusion={name}", "--text_model=distilbert"])
        with redirect_stdout(io.StringIO()):
            m = build_model(a, pretrained=False)
        assert type(m) is cls
        assert len(m.state_dict()) > 1000        # the shared base's full state_dict, whatever the variant
    with pytest.raises(SystemExit):
        build_model(args_parser(["--late_fusion=clip"]), pretrained=False)


def test_feature_cache_logic_on_cpu():
    """training.FeatureCache: all-or-nothing lookup by dataset index, store, invalidate (device-agnostic plumbing)."""
    from garbage_classification_rca_b200.training import FeatureCache
    c = FeatureCache(10, 4, 3, "cpu", dtype=torch.float32)
    ids = torch.tensor([7, 2, 9])
    assert c.lookup(ids) is None and c.misses == 3
    img, txt = torch.arange(12.0).view(3, 4), torch.arange(9.0).view(3, 3)
    c.store(ids, img, txt)
    got = c.lookup(torch.tensor([9, 7]))
    assert torch.equal(got[0], img[[2, 0]]) and torch.equal(got[1], txt[[2, 0]]) and c.hits == 2
    assert c.lookup(torch.tensor([9, 1])) is None          # one uncached id: the whole batch goes through the backbones
    c.invalidate()
    assert c.lookup(ids) is None
