"""GPU parity tests: the CUDA path, called through the C ABI (libmmrca.so via the Python host layer),
against (a) golden vectors produced by the unmodified reference and (b) the CPU oracle on seeded inputs.

Tolerances (BASELINE.json north_star): fp32 logits within 1e-4 relative, head gradients within 1e-2
relative; the fp32 kernels are additionally held to a much tighter bound (tests/_util.py).
"""
import os

import numpy as np
import pytest
import torch

from oracle import mmrca_oracle as orc
from tests._util import (GOLDEN, GRAD_REL, GRAD_REL_FP32_TIGHT, HEAD_CASES, LOGITS_REL_FP32, assert_grad_close,
                         load_case, make_inputs, rel_err)

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pkg(native_lib):
    import garbage_classification_rca_b200 as g
    assert torch.cuda.is_available(), "GPU tests need a B200"
    assert native_lib.mmrca_query(g._native.QUERY_DEVICE_OK) == 1, "device is not sm_100"
    return g


def _run_cuda_head(g, p, img, txt, labels, flags, cw=None, eps=0.0, dm=None, ds=1.0, feature_grads=True):
    from garbage_classification_rca_b200.training import CrossEntropyLoss
    rev, fo, co = flags
    names = g.head_param_names(fo, co)
    params = [p[n].cuda().requires_grad_(True) for n in names]
    img_c = img.cuda().requires_grad_(feature_grads)
    txt_c = txt.cuda().requires_grad_(feature_grads)
    g._native.kernel_launches(reset=True)
    logits = g.mmrca_head(img_c, txt_c, params, reverse=rev, features_only=fo, cross_attention_only=co,
                          drop_mask=None if dm is None else dm.cuda(), drop_scale=ds)
    crit = CrossEntropyLoss(None if cw is None else cw.cuda(), eps)
    loss = crit(logits, labels.cuda())
    loss.backward()
    torch.cuda.synchronize()
    assert g._native.kernel_launches() > 0, "no libmmrca kernel was launched"
    grads = {n: (t.grad.cpu() if t.grad is not None else None) for n, t in zip(names, params)}
    return (logits.detach().cpu(), loss.item(), grads,
            img_c.grad.cpu() if feature_grads else None, txt_c.grad.cpu() if feature_grads else None)


@pytest.mark.parametrize("case", HEAD_CASES)
def test_head_matches_reference_golden(pkg, case):
    d, flags, p = load_case(case)
    cw = torch.tensor(d["class_weight"]) if "class_weight" in d.files else None
    dm = torch.tensor(d["drop_mask"]) if "drop_mask" in d.files else None
    logits, loss, grads, d_img, d_txt = _run_cuda_head(
        pkg, p, torch.tensor(d["img"]), torch.tensor(d["txt"]), torch.tensor(d["labels"]), flags, cw,
        float(d["label_smoothing"]), dm, float(d["drop_scale"]))
    assert rel_err(logits.numpy(), d["logits"]) < LOGITS_REL_FP32
    assert abs(loss - float(d["loss"])) < 1e-5 * max(1.0, abs(float(d["loss"])))
    assert rel_err(d_img.numpy(), d["d_img"]) < GRAD_REL_FP32_TIGHT
    assert rel_err(d_txt.numpy(), d["d_txt"]) < GRAD_REL_FP32_TIGHT
    n = 0
    for k in d.files:
        if k.startswith("grad/"):
            n += 1
            assert grads[k[5:]] is not None, k
            assert_grad_close(k[5:], grads[k[5:]].numpy(), d[k], GRAD_REL_FP32_TIGHT)
    assert n == int(d["n_grad_tensors"])
    if flags[1]:   # features_only: attention blocks are outside the graph, like in the reference
        assert all(v is None for k, v in grads.items() if not k.startswith("final_"))


@pytest.mark.parametrize("flags", [(True, False, False), (False, False, False), (True, True, False),
                                   (True, False, True), (True, True, True)],
                         ids=["rca", "ca", "features_only", "cross_only", "features_only+cross_only"])
@pytest.mark.parametrize("B", [1, 3, 8, 17, 64])
def test_head_matches_oracle_seeded(pkg, flags, B):
    """Ragged batch sizes (not a multiple of the 2/4-sample tiles), all four switch combinations,
    trained-like (sharp) attention, class weights + label smoothing."""
    rev, fo, co = flags
    seed = 100 + B
    p = orc.init_head_params(features_only=fo, cross_attention_only=co, seed=seed, qk_gain=40.0)
    img, txt, labels = make_inputs(B, seed)
    cw = torch.tensor([1.3, 0.5, 1.0, 0.8])
    ref = orc.np_head_forward_backward(p, img.numpy(), txt.numpy(), rev, fo, co, labels=labels.numpy(),
                                       class_weight=cw.numpy(), label_smoothing=0.05)
    logits, loss, grads, d_img, d_txt = _run_cuda_head(pkg, p, img, txt, labels, flags, cw, 0.05)
    assert rel_err(logits.numpy(), ref["logits"]) < LOGITS_REL_FP32
    assert abs(loss - ref["loss"]) < 1e-5 * max(1.0, abs(ref["loss"]))
    assert rel_err(d_img.numpy(), ref["d_img"]) < GRAD_REL_FP32_TIGHT
    assert rel_err(d_txt.numpy(), ref["d_txt"]) < GRAD_REL_FP32_TIGHT
    for name, gt in grads.items():
        if gt is None:
            assert fo and not name.startswith("final_")
            continue
        assert_grad_close(name, gt.numpy(), ref["grads"][name], GRAD_REL_FP32_TIGHT)


@pytest.mark.parametrize("dims", [(768, 768), (1024, 768), (1024, 1024)], ids=["vitb", "convnext", "vitl-bart"])
def test_other_backbone_widths(pkg, dims):
    """Parametric d_img / d_txt (any pooled width with 16 chunks of 48/64/80)."""
    d_img, d_txt = dims
    p = orc.init_head_params(d_img=d_img, d_txt=d_txt, seed=5, qk_gain=30.0)
    img, txt, labels = make_inputs(9, 5, d_img, d_txt)
    ref = orc.np_head_forward_backward(p, img.numpy(), txt.numpy(), True, False, False, labels=labels.numpy())
    logits, loss, grads, d_i, d_t = _run_cuda_head(pkg, p, img, txt, labels, (True, False, False))
    assert rel_err(logits.numpy(), ref["logits"]) < LOGITS_REL_FP32
    assert rel_err(d_i.numpy(), ref["d_img"]) < GRAD_REL_FP32_TIGHT
    for name, gt in grads.items():
        assert_grad_close(name, gt.numpy(), ref["grads"][name], GRAD_REL_FP32_TIGHT)


def _comp(d, pre):
    order = ("W_query.weight", "W_query.bias", "W_key.weight", "W_key.bias", "W_value.weight", "W_value.bias",
             "norm.weight", "norm.bias")
    return [torch.tensor(d[pre + k]).cuda() for k in order]


def test_attention_blocks_match_reference_classes(pkg):
    """Stand-alone blocks against mm.SelfAttention / mm.ReverseCrossAttention outputs (components.npz)."""
    d = np.load(os.path.join(GOLDEN, "components.npz"))
    x1, x2, xi = (torch.tensor(d[k]).cuda() for k in ("x1", "x2", "xi"))
    rca = _comp(d, "rca/")
    out = pkg.attention_block(x1, x2, rca, reverse=True)
    assert rel_err(out.cpu().numpy(), d["rca_out"]) < LOGITS_REL_FP32
    assert abs(out.sum().item() - 585.30798) < 2e-2
    out = pkg.attention_block(x1, x2, rca, reverse=False)
    assert rel_err(out.cpu().numpy(), d["ca_out"]) < LOGITS_REL_FP32
    out = pkg.attention_block(xi, None, _comp(d, "sa/"))
    assert rel_err(out.cpu().numpy(), d["sa_out"]) < LOGITS_REL_FP32
    assert abs(out.sum().item() - 1249.70618) < 5e-2


@pytest.mark.parametrize("kind", ["sa48", "sa64", "sa80", "rca", "ca"])
def test_attention_block_gradients(pkg, kind):
    """O(1) random inputs: exercises the softmax / reverse backward far from the uniform-attention regime."""
    torch.manual_seed(11)
    B = 7
    if kind.startswith("sa"):
        d_in, dkq, dv, rev, self_ = int(kind[2:]), 128, 96, False, True
    else:
        d_in, dkq, dv, rev, self_ = 96, 64, 48, kind == "rca", False
    blk = "b"
    p = {f"{blk}.W_query.weight": torch.randn(dkq, d_in) * 0.3, f"{blk}.W_query.bias": torch.randn(dkq) * 0.1,
         f"{blk}.W_key.weight": torch.randn(dkq, d_in) * 0.3, f"{blk}.W_key.bias": torch.randn(dkq) * 0.1,
         f"{blk}.W_value.weight": torch.randn(dv, d_in) * 0.3, f"{blk}.W_value.bias": torch.randn(dv) * 0.1,
         f"{blk}.norm.weight": 1 + 0.2 * torch.randn(dv), f"{blk}.norm.bias": 0.2 * torch.randn(dv)}
    xq = torch.randn(B, 16, d_in)
    xkv = xq if self_ else torch.randn(B, 16, d_in)
    gout = torch.randn(B, 16, dv)
    pr = {k: v.double().requires_grad_(True) for k, v in p.items()}
    xq_r = xq.double().requires_grad_(True)
    xkv_r = xq_r if self_ else xkv.double().requires_grad_(True)
    ref = orc.self_attention(xq_r, pr, blk) if self_ else orc.reverse_cross_attention(xq_r, xkv_r, pr, blk, rev)
    ref.backward(gout.double())
    order = ("W_query.weight", "W_query.bias", "W_key.weight", "W_key.bias", "W_value.weight", "W_value.bias",
             "norm.weight", "norm.bias")
    pc = [p[f"{blk}.{k}"].cuda().requires_grad_(True) for k in order]
    xq_c = xq.cuda().requires_grad_(True)
    xkv_c = None if self_ else xkv.cuda().requires_grad_(True)
    out = pkg.attention_block(xq_c, xkv_c, pc, reverse=rev)
    out.backward(gout.cuda())
    assert rel_err(out.detach().cpu().numpy(), ref.detach().numpy()) < LOGITS_REL_FP32
    assert rel_err(xq_c.grad.cpu().numpy(), xq_r.grad.numpy()) < GRAD_REL_FP32_TIGHT
    if not self_:
        assert rel_err(xkv_c.grad.cpu().numpy(), xkv_r.grad.numpy()) < GRAD_REL_FP32_TIGHT
    scale = max(float(v.grad.abs().max()) for v in pr.values())
    for k, t in zip(order, pc):
        assert_grad_close(f"{blk}.{k}", t.grad.cpu().numpy(), pr[f"{blk}.{k}"].grad.numpy(), GRAD_REL_FP32_TIGHT,
                          scale)


def test_cross_entropy_kernel(pkg):
    g = torch.Generator().manual_seed(3)
    logits = torch.randn(1500, 4, generator=g) * 2
    labels = torch.randint(0, 4, (1500,), generator=g)
    w = torch.tensor([0.5, 2.0, 1.0, 1.5])
    for weight in (None, w):
        for eps in (0.0, 0.1):
            l_ref, dl_ref = orc.np_cross_entropy_fwd_bwd(logits.numpy(), labels.numpy(),
                                                         None if weight is None else weight.numpy(), eps)
            loss, dl = pkg.cross_entropy(logits.cuda(), labels.cuda(), None if weight is None else weight.cuda(), eps)
            assert abs(loss.item() - l_ref) < 2e-6 * max(1, abs(l_ref))
            assert rel_err(dl.cpu().numpy(), dl_ref) < 1e-5


def test_train_step_equals_autograd_path(pkg):
    """One-call fused step (mmrca_head_train_step) == forward + CrossEntropyLoss + backward through autograd."""
    B = 37
    p = orc.init_head_params(seed=9, qk_gain=40.0)
    img, txt, labels = make_inputs(B, 9)
    logits, loss, grads, d_img, d_txt = _run_cuda_head(pkg, p, img, txt, labels, (True, False, False))
    names = pkg.head_param_names()
    step = pkg.HeadTrainStep([p[n].cuda() for n in names], B, 1280, 768, reverse=True, feature_grads=True)
    for _ in range(2):   # gradients accumulate, like loss.backward()
        l2, lg2 = step(img.cuda(), txt.cuda(), labels.cuda())
    torch.cuda.synchronize()
    assert abs(l2.item() - loss) < 1e-6
    assert rel_err(lg2.cpu().numpy(), logits.numpy()) < 1e-6
    assert rel_err(step.d_img.cpu().numpy(), d_img.numpy()) < 1e-5
    for n, v in zip(names, step.grads.views):
        assert_grad_close(n, v.cpu().numpy() / 2.0, grads[n].numpy(), 1e-4)
    step.zero_grad()
    assert float(step.grads.flat.abs().max()) == 0.0


def test_full_size_properties(pkg):
    """BASELINE config 2 size (B = 4096): size-independent properties instead of a CPU recomputation of
    everything — permutation equivariance (bit-exact: a sample's arithmetic does not depend on its slot),
    shard additivity of the gradients (what data parallelism relies on), linearity in dlogits, plus a
    256-sample slice against the oracle."""
    B = 4096
    p = orc.init_head_params(seed=21, qk_gain=40.0)
    names = pkg.head_param_names()
    params = [p[n].cuda() for n in names]
    img, txt, labels = make_inputs(B, 21)
    img_c, txt_c, lab_c = img.cuda(), txt.cuda(), labels.cuda()
    logits = pkg.mmrca_head(img_c, txt_c, params, reverse=True)
    perm = torch.randperm(B, generator=torch.Generator().manual_seed(1)).cuda()
    logits_p = pkg.mmrca_head(img_c[perm].contiguous(), txt_c[perm].contiguous(), params, reverse=True)
    assert torch.equal(logits[perm], logits_p)
    ref = orc.head_forward(p, img[:256], txt[:256], True)
    assert rel_err(logits[:256].cpu().numpy(), ref.numpy()) < LOGITS_REL_FP32

    def grads_of(sl, scale=1.0):
        step = pkg.HeadTrainStep(params, sl.stop - sl.start, 1280, 768, reverse=True)
        step(img_c[sl].contiguous(), txt_c[sl].contiguous(), lab_c[sl].contiguous())
        return step.grads.flat.clone() * scale

    full = grads_of(slice(0, B))
    # mean-reduced loss: the full-batch gradient is the average of the two half-batch gradients
    halves = 0.5 * (grads_of(slice(0, B // 2)) + grads_of(slice(B // 2, B)))
    assert rel_err(halves.cpu().numpy(), full.cpu().numpy()) < 1e-4
    # linearity in dlogits through the explicit backward
    ps = [t.clone().requires_grad_(True) for t in params]
    out = pkg.mmrca_head(img_c, txt_c, ps, reverse=True)
    dl = torch.randn(B, 4, generator=torch.Generator().manual_seed(2)).cuda() / B
    g1 = torch.autograd.grad(out, ps, dl, retain_graph=True)
    g2 = torch.autograd.grad(out, ps, 2.0 * dl)
    for a, b, n in zip(g1, g2, names):
        assert_grad_close(n, (b / 2.0).cpu().numpy(), a.cpu().numpy(), 1e-4)


def test_empty_batch_and_errors(pkg):
    p = orc.init_head_params(seed=1)
    params = [p[n].cuda() for n in pkg.head_param_names()]
    out = pkg.mmrca_head(torch.zeros(0, 1280).cuda(), torch.zeros(0, 768).cuda(), params, reverse=True)
    assert out.shape == (0, 4)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        pkg.mmrca_head(torch.zeros(2, 1280), torch.zeros(2, 768).cuda(), params, reverse=True)
    with pytest.raises(RuntimeError, match="d_img"):
        pkg.mmrca_head(torch.zeros(2, 1000).cuda(), torch.zeros(2, 768).cuda(), params, reverse=True)
    # zero feature rows: the reference's eps-free L2 norm yields NaN (multimodal_model.py:662-665); mirrored
    out = pkg.mmrca_head(torch.zeros(2, 1280).cuda(), torch.ones(2, 768).cuda(), params, reverse=True)
    assert torch.isnan(out).all()


def test_module_drop_in_with_stub_backbones(pkg):
    """The nn.Module mirror: reference ctor signature + forward(_input_ids, _attention_mask, _images, ...)
    with the backbones replaced by feature stubs, against the oracle on the module's own state_dict."""
    import io
    from contextlib import redirect_stdout
    from garbage_classification_rca_b200 import multimodal_model as M

    class StubText(torch.nn.Module):
        def forward(self, input_ids=None, attention_mask=None, **kw):
            return (self.feat.unsqueeze(1),)

    class StubImage(torch.nn.Module):
        def forward(self, x):
            return None, None, self.feat

    torch.manual_seed(0)
    with redirect_stdout(io.StringIO()):
        m = M.MM_RCA(4, 0.6, 0.0, 0.7, 256, "distilbert", 16, True, False, False, pretrained=False)
    m.text_model, m.image_model = StubText(), StubImage()
    m = m.cuda()
    img, txt, labels = make_inputs(10, 77)
    m.text_model.feat, m.image_model.feat = txt.cuda(), img.cuda()
    ids = torch.zeros(10, 8, dtype=torch.long).cuda()
    m.eval()
    with torch.no_grad():
        out = m(_input_ids=ids, _attention_mask=torch.ones_like(ids), _images=torch.zeros(10, 3, 8, 8).cuda(),
                eval=True, remove_image=False, remove_text=False)
    sd = {k: v.detach().cpu() for k, v in m.state_dict().items()}
    ref = orc.head_forward(sd, img, txt, True)
    assert rel_err(out.cpu().numpy(), ref.numpy()) < LOGITS_REL_FP32
    # training mode: dropout p=0.6 is drawn inside the kernels from a seed the module takes from torch's
    # generator; parity through the materialised mask of that seed
    m.train()
    out = m.forward_features(img.cuda(), txt.cuda())
    mask, scale = pkg.functional.dropout_mask(m.last_dropout_seed, 0.6, 10, 3584, "cuda"), 1.0 / (1.0 - 0.6)
    ref = orc.head_forward(sd, img, txt, True, drop_mask=mask.cpu(), drop_scale=scale)
    assert rel_err(out.detach().cpu().numpy(), ref.numpy()) < LOGITS_REL_FP32
    assert 0.3 < mask.float().mean().item() < 0.5
    loss = torch.nn.functional.cross_entropy(out, labels.cuda())
    loss.backward()
    assert m.final_with_everything.weight.grad is not None and m.cross_attention_1.W_value.weight.grad is not None
    assert m.image_to_hidden_size.weight.grad is None   # dead parameters stay outside the graph


def _peer_allreduce_worker(rank, world, port, out):
    import os
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", device_id=dev)
    from garbage_classification_rca_b200.training import PeerAllReduce, allreduce_mean_
    n = 94824
    ar = PeerAllReduce(n, dev)
    worst = 0.0
    for it in range(6):
        g = torch.Generator(device=dev).manual_seed(100 * it + rank)
        x = torch.randn(n, device=dev, generator=g)
        ref = allreduce_mean_(x.clone())
        got = ar(x.clone())
        worst = max(worst, (got - ref).abs().max().item())
        chk = got.clone()
        dist.broadcast(chk, 0)
        assert torch.equal(chk, got)          # bit-identical on every rank
    out[rank] = worst
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs with NVLink peer access")
def test_peer_memory_allreduce_matches_nccl(pkg):
    """The one collective of a data-parallel step as the one-shot NVLink peer-memory kernel
    (mmrca_peer_allreduce_mean) against NCCL; skipped on a single-GPU box (tools/check_peer_allreduce.py
    runs the same check under torchrun at 2 and 8 GPUs)."""
    import torch.multiprocessing as mp
    world = 2
    with mp.Manager() as m:
        out = m.dict()
        mp.spawn(_peer_allreduce_worker, args=(world, 29631, out), nprocs=world, join=True)
        assert len(out) == world and max(out.values()) <= 1e-6


def test_module_with_stock_backbones_cfg1(pkg):
    """BASELINE.json configs[0] through the drop-in module with the real (random-init, no network) stock backbones:
    EfficientNetV2-M feature extractor + DistilBERT, batch 8, 3x384x384 images, 128-token text, eval forward.  The
    head's logits are checked against the oracle on the features the module's own backbones produced."""
    import io
    from contextlib import redirect_stdout
    from garbage_classification_rca_b200 import multimodal_model as M
    torch.manual_seed(0)
    with redirect_stdout(io.StringIO()):
        m = M.MM_RCA(4, 0.6, 0.0, 0.7, 256, "distilbert", 8, True, False, False, pretrained=False)
    m = m.cuda().eval()
    images = torch.randn(8, 3, 384, 384, device="cuda")
    ids = torch.randint(0, 30522, (8, 128), device="cuda")
    mask = torch.ones_like(ids)
    with torch.no_grad():
        out = m(ids, mask, images, eval=True)
        m._images, m._input_ids, m._attention_mask = images, ids, mask
        _, txt, (_, _, img) = m._backbone_features()
    assert out.shape == (8, 4) and torch.isfinite(out).all()
    assert img.shape == (8, 1280) and txt.shape == (8, 768)
    sd = {k: v.detach().cpu() for k, v in m.state_dict().items() if not k.startswith(("image_model.", "text_model."))}
    ref = orc.head_forward(sd, img.float().cpu(), txt.float().cpu(), True)
    assert rel_err(out.cpu().numpy(), ref.numpy()) < LOGITS_REL_FP32
