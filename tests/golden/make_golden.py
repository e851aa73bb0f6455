"""Generate the golden fixtures in this directory from the UNMODIFIED reference.

Run in the build container only (needs /root/reference, which does not exist on the
GPU box):   python tests/golden/make_golden.py

The reference classes do all the math.  The only patches are the ones SURVEY.md §8(c)
lists: the weight-downloading backbone factories (multimodal_model.py:113-153) are
replaced by random-init / stub equivalents because there is no network.  Head-only
cases replace text_model / image_model by 3-line stubs that hand precomputed pooled
features to the unmodified MM_RCA.forward (multimodal_model.py:638-728).
"""
import hashlib
import io
import json
import os
import sys
from contextlib import redirect_stdout

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

import CVPR_code.multimodal_model as mm  # noqa: E402  (the reference, imported as-is)
from oracle import mmrca_oracle as orc   # noqa: E402  (only for init_head_params / names)


class _Cfg:
    hidden_size = 768


class StubText(torch.nn.Module):
    """Returns (feat.unsqueeze(1),) so that text_output[0][:, 0] == feat (multimodal_model.py:655-658)."""

    def __init__(self):
        super().__init__()
        self.config = _Cfg()
        self.feat = None

    def forward(self, input_ids=None, attention_mask=None, **kw):
        return (self.feat.unsqueeze(1),)


class StubImage(torch.nn.Module):
    """Returns (None, None, feat) like EfficientNetV2MFullFeatureExtractor (multimodal_model.py:36)."""

    def __init__(self):
        super().__init__()
        self.feat = None

    def forward(self, x):
        return None, None, self.feat


def build_reference(reverse, features_only, cross_attention_only, drop=0.0, stub=True):
    if stub:
        mm.distilbert = lambda: StubText()
        mm.bert = lambda: StubText()
        mm.eff_net_v2 = lambda: StubImage()
    with redirect_stdout(io.StringIO()):
        m = mm.MM_RCA(4, drop, 0.0, 0.7, 256, "distilbert", 16, reverse, features_only, cross_attention_only)
    return m


def run_case(name, reverse, features_only, cross_attention_only, B, seed,
             class_weight=None, label_smoothing=0.0, use_drop_mask=False, qk_gain=1.0):
    torch.manual_seed(seed)
    m = build_reference(reverse, features_only, cross_attention_only)
    params = orc.init_head_params(features_only=features_only, cross_attention_only=cross_attention_only,
                                  seed=seed, qk_gain=qk_gain)
    missing, unexpected = m.load_state_dict(params, strict=False)
    assert not unexpected, unexpected
    g = torch.Generator().manual_seed(1000 + seed)
    img = torch.randn(B, 1280, generator=g) * 0.7 + 0.1
    txt = torch.randn(B, 768, generator=g) * 1.3 - 0.05
    labels = torch.randint(0, 4, (B,), generator=g)
    img.requires_grad_(True)
    txt.requires_grad_(True)
    m.text_model.feat = txt
    m.image_model.feat = img
    m.train()
    ids = torch.zeros(B, 8, dtype=torch.long)
    d_cat = orc.concat_width(1280, 768, features_only, cross_attention_only)
    drop_mask = None
    drop_scale = 1.0
    if use_drop_mask:
        # torch's Philox stream cannot be reproduced by a custom kernel (SURVEY.md §7 hard parts):
        # take the mask torch.nn.Dropout itself draws and store it in the fixture.
        p_drop = 0.6
        m.drop = torch.nn.Dropout(p=p_drop)
        probe = torch.ones(B, d_cat)
        st = torch.get_rng_state()
        drop_mask = (m.drop(probe) != 0)
        torch.set_rng_state(st)      # the real forward below draws the identical mask
        drop_scale = 1.0 / (1.0 - p_drop)
    logits = m(_input_ids=ids, _attention_mask=torch.ones_like(ids), _images=torch.zeros(B, 3, 4, 4))
    cw = None if class_weight is None else torch.tensor(class_weight, dtype=torch.float32)
    crit = torch.nn.CrossEntropyLoss(weight=cw, label_smoothing=label_smoothing)   # main_both.py:87-93
    loss = crit(logits, labels)
    loss.backward()                                                               # main_both.py:112
    out = dict(img=img.detach().numpy(), txt=txt.detach().numpy(), labels=labels.numpy(),
               logits=logits.detach().numpy(), loss=np.float32(loss.item()),
               d_img=img.grad.numpy(), d_txt=txt.grad.numpy(),
               flags=np.array([reverse, features_only, cross_attention_only], dtype=np.uint8),
               seed=np.int64(seed), qk_gain=np.float32(qk_gain), label_smoothing=np.float32(label_smoothing),
               drop_scale=np.float32(drop_scale))
    if cw is not None:
        out["class_weight"] = cw.numpy()
    if drop_mask is not None:
        out["drop_mask"] = drop_mask.numpy().astype(np.uint8)
    sd = dict(m.named_parameters())
    n_grad = 0
    for k in orc.head_param_names(features_only, cross_attention_only):
        gk = sd[k].grad
        if gk is None:
            continue
        n_grad += 1
        out["grad/" + k] = gk.numpy().astype(np.float32)
    out["n_grad_tensors"] = np.int64(n_grad)
    np.savez_compressed(os.path.join(HERE, f"head_{name}.npz"), **out)
    print(f"{name}: loss={loss.item():.6f} logits[0]={logits[0].tolist()} grads={n_grad}")


def components():
    """Standalone SelfAttention / ReverseCrossAttention KATs incl. the SURVEY §8(c) anchors."""
    torch.manual_seed(1234)
    rca = mm.ReverseCrossAttention(96, 96, 64, 48, True)
    sa = mm.SelfAttention(80, 128, 96, "k")
    x1 = torch.randn(2, 16, 96)
    x2 = torch.randn(2, 16, 96)
    xi = torch.randn(2, 16, 80)
    out = dict(x1=x1.numpy(), x2=x2.numpy(), xi=xi.numpy())
    with torch.no_grad():
        o_rca = rca(x1, x2)
        rca.reverse = False
        o_ca = rca(x1, x2)
        o_sa = sa(xi)
    print("anchor sums:", o_rca.sum().item(), o_ca.sum().item(), o_sa.sum().item())
    out.update(rca_out=o_rca.numpy(), ca_out=o_ca.numpy(), sa_out=o_sa.numpy(),
               anchor_sums=np.array([o_rca.sum().item(), o_ca.sum().item(), o_sa.sum().item()]))
    for k, v in rca.state_dict().items():
        out["rca/" + k] = v.numpy()
    for k, v in sa.state_dict().items():
        out["sa/" + k] = v.numpy()
    # a second KAT at another square L (token-level extension, SURVEY §0 row "config 5")
    torch.manual_seed(77)
    rca2 = mm.ReverseCrossAttention(96, 96, 64, 48, True)
    y1, y2 = torch.randn(3, 24, 96), torch.randn(3, 24, 96)
    with torch.no_grad():
        out.update(L24_x1=y1.numpy(), L24_x2=y2.numpy(), L24_out=rca2(y1, y2).numpy())
    for k, v in rca2.state_dict().items():
        out["rca24/" + k] = v.numpy()
    np.savez_compressed(os.path.join(HERE, "components.npz"), **out)


def state_dict_layout():
    """Names / shapes / dtypes of the full reference state_dict with real (random-init) backbones,
    for the drop-in layout test (SURVEY.md §5 'Checkpoint')."""
    from torchvision.models import efficientnet_v2_m
    from transformers import BertConfig, BertModel, DistilBertConfig, DistilBertModel

    def eff():
        model = efficientnet_v2_m(weights=None)
        model.classifier = torch.nn.Sequential(*[model.classifier[i] for i in range(1)])   # mirrors :120-121
        return mm.EfficientNetV2MFullFeatureExtractor(model)

    layouts = {}
    for text in ("distilbert", "bert"):
        mm.distilbert = lambda: DistilBertModel(DistilBertConfig())
        mm.bert = lambda: BertModel(BertConfig())
        mm.eff_net_v2 = eff
        for flags in ((True, False, False), (True, True, False), (True, False, True), (True, True, True)):
            with redirect_stdout(io.StringIO()):
                m = mm.MM_RCA(4, 0.6, 0.0, 0.7, 256, text, 16, *flags)
            sd = m.state_dict()
            key = f"{text}|features_only={int(flags[1])}|cross_attention_only={int(flags[2])}"
            ent = [[k, list(v.shape), str(v.dtype)] for k, v in sd.items()]
            head = [e for e in ent if not e[0].startswith(("image_model.", "text_model."))]
            digest = {}
            for pre in ("image_model.", "text_model."):
                sub = [e for e in ent if e[0].startswith(pre)]
                h = hashlib.sha256("\n".join(f"{k}:{s}:{d}" for k, s, d in sub).encode()).hexdigest()
                digest[pre] = dict(count=len(sub), sha256=h, first=sub[0][0], last=sub[-1][0])
            layouts[key] = dict(order_total=len(ent), head=head, backbones=digest,
                                head_first_index=[e[0] for e in ent].index(head[0][0]))
            print(key, len(sd), sum(v.numel() for v in sd.values()))
    with open(os.path.join(HERE, "state_dict_layout.json"), "w") as f:
        json.dump(layouts, f)


if __name__ == "__main__":
    components()
    run_case("rca_full", True, False, False, B=6, seed=0)
    run_case("ca_full", False, False, False, B=6, seed=1)
    run_case("rca_features_only", True, True, False, B=6, seed=2)
    run_case("rca_cross_only", True, False, True, B=6, seed=3)
    run_case("rca_full_weighted_smooth", True, False, False, B=5, seed=4,
             class_weight=[0.6, 1.7, 0.9, 1.2], label_smoothing=0.1)
    run_case("rca_full_dropout", True, False, False, B=4, seed=5, use_drop_mask=True)
    run_case("rca_full_sharp", True, False, False, B=6, seed=6, qk_gain=40.0)
    run_case("ca_full_sharp", False, False, False, B=6, seed=7, qk_gain=40.0)
    run_case("rca_cross_only_sharp", True, False, True, B=5, seed=8, qk_gain=40.0,
             class_weight=[1.3, 0.5, 1.0, 0.8], label_smoothing=0.05)
    state_dict_layout()
