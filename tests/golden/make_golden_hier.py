"""Golden fixtures of the hierarchical late-fusion head from the UNMODIFIED reference.

Run in the build container only (needs /root/reference):   python tests/golden/make_golden_hier.py

The reference class Hierarchical (CVPR_code/multimodal_model.py:729-818) does all the math, including the two
AvgPool2d calls; the backbones are 3-line stubs that hand over seeded feature maps / hidden states (there is no
network for the pretrained weights).  The 12 MB weight of final_hierarchical_image is a pure function
of the seed (oracle.init_hier_params), so the fixture stores the inputs, the logits, the loss, the small gradients in
full and a strided sample + the Frobenius norm of the two large weight gradients.
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

import CVPR_code.multimodal_model as mm  # noqa: E402  (the reference, imported as-is)
from oracle import mmrca_oracle as orc   # noqa: E402  (only for init_hier_params / names)

ROW_STEP, COL_STEP = 37, 53      # strided sample of the large weight gradients


class _Cfg:
    hidden_size = 768


class _TextOut:
    """text_output[0] and text_output.hidden_states, the two things Hierarchical.forward reads (:744-757)."""

    def __init__(self, last, hidden_states):
        self._last, self.hidden_states = last, hidden_states

    def __getitem__(self, i):
        return (self._last,)[i]


class StubText(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.config = _Cfg()
        self.cls = None          # (last, layer2, layer4) CLS features, each [B,768]

    def forward(self, input_ids=None, attention_mask=None, output_hidden_states=False, **kw):
        last, l2, l4 = (t.unsqueeze(1) for t in self.cls)
        return _TextOut(last, (None, None, l2, None, l4))


class StubImage(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.maps = None         # (stage3 [B,160,28,28], stage6 [B,512,12,12], pooled [B,1280])

    def forward(self, x):
        return self.maps


def run_case(name, B, seed, class_weight=None, label_smoothing=0.0, p_drop=0.0, bias_gap=0.05):
    torch.manual_seed(seed)
    mm.distilbert = lambda: StubText()
    mm.bert = lambda: StubText()
    mm.eff_net_v2 = lambda: StubImage()
    with redirect_stdout(io.StringIO()):
        m = mm.Hierarchical(4, p_drop, 0.0, 0.7, 256, "distilbert", 16, True, False, False)
    params = orc.init_hier_params(seed=seed, bias_gap=bias_gap)
    missing, unexpected = m.load_state_dict(params, strict=False)
    assert not unexpected, unexpected
    g = torch.Generator().manual_seed(2000 + seed)
    s3 = torch.randn(B, 160, 28, 28, generator=g).abs() * 0.6            # post-activation-like maps
    s6 = torch.randn(B, 512, 12, 12, generator=g).abs() * 0.4
    pooled = torch.randn(B, 1280, generator=g) * 0.7 + 0.1
    cls = [torch.randn(B, 768, generator=g) * 1.3 - 0.05 for _ in range(3)]
    labels = torch.randint(0, 4, (B,), generator=g)
    m.image_model.maps = (s3, s6, pooled)
    m.text_model.cls = cls
    m.train()
    drop_mask, drop_scale = None, 1.0
    if p_drop > 0:
        # the two self.drop calls (:805-806) draw from torch's stream in this order: take the masks torch draws
        st = torch.get_rng_state()
        mi = (m.drop(torch.ones(B, orc.HIER_D_IMG)) != 0)
        mt = (m.drop(torch.ones(B, orc.HIER_D_TXT)) != 0)
        torch.set_rng_state(st)
        drop_mask = torch.cat((mi, mt), dim=1)
        drop_scale = 1.0 / (1.0 - p_drop)
    ids = torch.zeros(B, 8, dtype=torch.long)
    logits = m(_input_ids=ids, _attention_mask=torch.ones_like(ids), _images=torch.zeros(B, 3, 4, 4))
    cw = None if class_weight is None else torch.tensor(class_weight, dtype=torch.float32)
    loss = torch.nn.CrossEntropyLoss(weight=cw, label_smoothing=label_smoothing)(logits, labels)   # main_both.py:87-93
    loss.backward()
    # what the B200 head is handed: the pooled + flattened maps (the AvgPool2d stay on the stock side, :761-775)
    s3p = torch.nn.AvgPool2d(kernel_size=7, stride=7)(s3).flatten(1)
    s6p = torch.nn.AvgPool2d(kernel_size=6, stride=6)(s6).flatten(1)
    out = dict(img_pooled=pooled.numpy(), img_s3=s3p.numpy(), img_s6=s6p.numpy(),
               txt_last=cls[0].numpy(), txt_l2=cls[1].numpy(), txt_l4=cls[2].numpy(), labels=labels.numpy(),
               logits=logits.detach().numpy(), loss=np.float32(loss.item()), seed=np.int64(seed),
               label_smoothing=np.float32(label_smoothing), drop_scale=np.float32(drop_scale),
               sample_steps=np.array([ROW_STEP, COL_STEP]), bias_gap=np.float32(bias_gap))
    if cw is not None:
        out["class_weight"] = cw.numpy()
    if drop_mask is not None:
        out["drop_mask"] = drop_mask.numpy().astype(np.uint8)
    sd = dict(m.named_parameters())
    for k in orc.HIER_PARAM_NAMES:
        gk = sd[k].grad.numpy().astype(np.float32)
        if gk.size > 100_000:
            out["grad_sample/" + k] = gk[::ROW_STEP, ::COL_STEP].copy()
            out["grad_norm/" + k] = np.float64(np.sqrt((gk.astype(np.float64) ** 2).sum()))
        else:
            out["grad/" + k] = gk
    np.savez_compressed(os.path.join(HERE, f"hier_{name}.npz"), **out)
    print(f"{name}: loss={loss.item():.6f} logits[0]={logits[0].tolist()}")


if __name__ == "__main__":
    run_case("plain", B=5, seed=0)
    run_case("weighted_smooth", B=6, seed=1, class_weight=[0.6, 1.7, 0.9, 1.2], label_smoothing=0.1)
    run_case("dropout", B=4, seed=2, p_drop=0.6)
    run_case("init_scale", B=6, seed=3, bias_gap=0.0)      # default-init biases: ReLU units near zero exist
