"""Golden fixtures of the Classic / Normalized late-fusion heads from the UNMODIFIED reference classes
(EffV2MediumAndDistilbertClassic / ...Normalized, multimodal_model.py:489-579).

Run in the build container only (needs /root/reference):   python tests/golden/make_golden_fusion.py

The reference's forward hands the image extractor's (stage3, stage6, pooled) TUPLE to image_to_hidden_size (:519-521),
a TypeError as shipped.  The only patch, besides the offline backbone factories (SURVEY.md §8 c), is therefore an image
stub that returns the pooled tensor itself; every arithmetic line of the two forwards runs unmodified.
Weights are a pure function of the seed (oracle.init_fusion_params), so the fixtures hold inputs' seeds + outputs only.
"""
import io
import os
import sys
from contextlib import redirect_stdout

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

import CVPR_code.multimodal_model as mm  # noqa: E402
from oracle import mmrca_oracle as orc   # noqa: E402


class _Cfg:
    hidden_size = 768


class StubText(torch.nn.Module):
    config = _Cfg()
    feat = None

    def forward(self, input_ids=None, attention_mask=None, **kw):
        return (self.feat.unsqueeze(1),)


class StubImage(torch.nn.Module):
    feat = None

    def forward(self, x):
        return self.feat


def sample_index(size):
    return np.random.default_rng(12345).choice(size, 256, replace=False)


def run_case(name, cls, normalized, B, seed, class_weight=None, label_smoothing=0.0, drop=0.0):
    mm.distilbert, mm.eff_net_v2 = (lambda: StubText()), (lambda: StubImage())
    with redirect_stdout(io.StringIO()):
        m = cls(4, drop, 0.0, 0.7, 256, "distilbert", 16, False, False, False)
    p = orc.init_fusion_params(seed=seed)
    missing, unexpected = m.load_state_dict(p, strict=False)
    assert not unexpected
    g = torch.Generator().manual_seed(3000 + seed)
    img = (torch.randn(B, 1280, generator=g) * 0.7 + 0.1).requires_grad_(True)
    txt = (torch.randn(B, 768, generator=g) * 1.3 - 0.05).requires_grad_(True)
    labels = torch.randint(0, 4, (B,), generator=g)
    m.text_model.feat, m.image_model.feat = txt, img
    mask = None
    if drop > 0:
        m.train()
        # nn.Dropout draws its mask from torch's generator: recover it from the reference's own output
        torch.manual_seed(777 + seed)
        with torch.no_grad(), redirect_stdout(io.StringIO()):
            probe = torch.nn.functional.dropout(torch.ones(B, 256), p=drop, training=True)
        mask = (probe > 0).to(torch.uint8)
        torch.manual_seed(777 + seed)
    else:
        m.eval()
    ids = torch.zeros(B, 4, dtype=torch.long)
    with redirect_stdout(io.StringIO()):
        logits = m(_input_ids=ids, _attention_mask=torch.ones_like(ids), _images=torch.zeros(B, 3, 4, 4))
    cw = None if class_weight is None else torch.tensor(class_weight)
    loss = torch.nn.CrossEntropyLoss(weight=cw, label_smoothing=label_smoothing)(logits, labels)
    loss.backward()
    sd = dict(m.named_parameters())
    out = dict(seed=seed, B=B, normalized=int(normalized), drop=drop, label_smoothing=label_smoothing,
               class_weight=np.asarray(class_weight if class_weight is not None else [], dtype=np.float32),
               img=img.detach().numpy(), txt=txt.detach().numpy(), labels=labels.numpy(),
               logits=logits.detach().numpy(), loss=float(loss), d_img=img.grad.numpy(), d_txt=txt.grad.numpy(),
               mask=mask.numpy() if mask is not None else np.zeros((0, 0), np.uint8))
    for n in orc.FUSION_PARAM_NAMES:
        gnp = sd[n].grad.numpy()
        if gnp.size <= 4096:
            out["grad." + n] = gnp
        else:      # large weight gradients: row sums, column sums and 256 sampled entries keep the fixture small
            out["gradrow." + n], out["gradcol." + n] = gnp.sum(axis=1), gnp.sum(axis=0)
            out["gradsample." + n] = gnp.ravel()[sample_index(gnp.size)]
    # the dead parameters stay outside the graph
    assert sd["final_with_everything.weight"].grad is None
    np.savez_compressed(os.path.join(HERE, f"fusion_{name}.npz"), **out)
    print(name, "loss", float(loss), "logits[0]", logits[0].tolist())


if __name__ == "__main__":
    run_case("classic", mm.EffV2MediumAndDistilbertClassic, False, 9, 1)
    run_case("normalized", mm.EffV2MediumAndDistilbertNormalized, True, 9, 2)
    run_case("classic_weighted_smooth_dropout", mm.EffV2MediumAndDistilbertClassic, False, 21, 3,
             class_weight=[0.6, 1.7, 1.0, 0.9], label_smoothing=0.1, drop=0.6)
    run_case("normalized_weighted_smooth_dropout", mm.EffV2MediumAndDistilbertNormalized, True, 21, 4,
             class_weight=[0.6, 1.7, 1.0, 0.9], label_smoothing=0.1, drop=0.6)
