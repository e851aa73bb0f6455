"""Development tool (CPU only, uses the oracle): a float64 emulation of the bf16 tensor-core pipeline
(csrc/mmrca_head_tc*.cuh) with every operand rounding as a named, switchable site.  It answers "which roundings
dominate the gradient error of the bf16 path" without spending GPU time: run it with all sites on (reproduces the
error level tools/diag_bf16.py measures on the GPU), then with sites switched off / given a hi+lo pair.

  python tools/emulate_bf16.py [B=200] [seed=31]

A site's mode: "bf16" (one bf16 rounding), "pair" (bf16 hi + bf16 lo: ~16 mantissa bits), "tf32" (10-bit mantissa), "off".
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from oracle import mmrca_oracle as orc

L = 16
EPS = 1e-5


def _round_bits(x, keep):          # round-to-nearest-even to `keep` explicit mantissa bits via float32 bit tricks
    f = x.to(torch.float32)
    i = f.view(torch.int32)
    drop = 23 - keep
    bias = ((i >> drop) & 1) + ((1 << (drop - 1)) - 1)
    i = ((i + bias) >> drop) << drop
    return i.view(torch.float32).to(torch.float64)


def rnd(x, mode):
    if mode == "off":
        return x
    if mode == "bf16":
        return _round_bits(x, 7)
    if mode == "tf32":
        return _round_bits(x, 10)
    if mode == "pair":
        hi = _round_bits(x, 7)
        return hi + _round_bits(x - hi, 7)
    if mode == "f32":
        return x.to(torch.float32).to(torch.float64)
    raise ValueError(mode)


class Sites(dict):
    def __call__(self, name, x):
        return rnd(x, self.get(name, self.get("*", "bf16")))


def attn_fwd(R, blk, xq, xkv, p, prefix, reverse):
    """xq, xkv: [B,16,din] operand values (already rounded by the producer)."""
    wq, bq = p[f"{prefix}.W_query.weight"], p[f"{prefix}.W_query.bias"]
    wk = p[f"{prefix}.W_key.weight"]
    wv, bv = p[f"{prefix}.W_value.weight"], p[f"{prefix}.W_value.bias"]
    g, b = p[f"{prefix}.norm.weight"], p[f"{prefix}.norm.bias"]
    s = 1.0 / np.sqrt(wq.shape[0])
    M = R(f"{blk}.M", s * (wq.T @ wk))              # [din(k), din(k')]
    u = R(f"{blk}.M", s * (wk.T @ bq))              # [din]  (Wk^T bq: indexed by k')... see note
    # scores row i, col j: (xq_i Wq^T + bq)(xkv_j Wk^T)^T s = xq_i (s Wq^T Wk) xkv_j^T + s bq^T Wk xkv_j^T
    Wv_, bv_ = R(f"{blk}.Wv", wv), R(f"{blk}.Wv", bv)
    xq_p, xkv_p = R(f"{blk}.Xproj", xq), R(f"{blk}.Xproj", xkv)      # the projections' copy of the inputs
    xkv_s = R(f"{blk}.Xscore", xkv)                                    # the score contraction's copy
    Z = R(f"{blk}.Z", xq_p @ M + u)
    vb_out = R.get("vbias_out", False)
    V = R(f"{blk}.V", xkv_p @ Wv_.T + (0.0 if vb_out else bv_))
    S = Z @ xkv_s.transpose(1, 2)
    A = torch.softmax(S, dim=-1)
    Pm = (1.0 - A) / (L - 1) if reverse else A
    Pb = R(f"{blk}.P", Pm)
    C = Pb @ V + (bv_ if vb_out else 0.0)
    mu = C.mean(-1, keepdim=True)
    var = ((C - mu) ** 2).mean(-1, keepdim=True)
    rstd = 1.0 / torch.sqrt(var + EPS)
    xhat = (C - mu) * rstd
    y = xhat * g + b
    cache = dict(xq=xq_p, xkv=xkv_s, xkv_p=xkv_p, vb_out=vb_out, M=M, u=u, Wv=Wv_, Z=Z, V=V, A=A, Pm=Pm, Pb=Pb, xhat=xhat, rstd=rstd, y=y,
                 reverse=reverse, s=s)
    return torch.relu(y), cache


def attn_bwd(R, blk, dout, c, p, prefix, grads, dy_mask=None, want_dx=True, p_for_softmax="fp32"):
    """dout: dL/d(relu(LN)) (before this block's dropout mask, which dy_mask applies)."""
    wq, bq = p[f"{prefix}.W_query.weight"], p[f"{prefix}.W_query.bias"]
    wk = p[f"{prefix}.W_key.weight"]
    g = p[f"{prefix}.norm.weight"]
    dy = dout * (c["y"] > 0)
    if dy_mask is not None:
        dy = dy * dy_mask
    t1 = R(f"{blk}.dyx", dy * c["xhat"])
    t2 = R(f"{blk}.dy", dy)
    grads[f"{prefix}.norm.weight"] = t1.sum((0, 1))
    grads[f"{prefix}.norm.bias"] = t2.sum((0, 1))
    dxh = dy * g
    dC = c["rstd"] * (dxh - dxh.mean(-1, keepdim=True) - c["xhat"] * (dxh * c["xhat"]).mean(-1, keepdim=True))
    dCb = R(f"{blk}.dC", dC)
    dP = dCb @ c["V"].transpose(1, 2)
    dV = R(f"{blk}.dV", c["Pb"].transpose(1, 2) @ dCb)
    Pm = c["Pb"] if p_for_softmax == "bf16" else c["Pm"]
    if c["reverse"]:
        A = 1.0 - (L - 1) * Pm
        dA = -dP / (L - 1)
    else:
        A, dA = Pm, dP
    dS = R(f"{blk}.dS", A * (dA - (dA * A).sum(-1, keepdim=True)))
    dZ = R(f"{blk}.dZ", dS @ c["xkv"])
    # parameter gradients (fp32 accumulation in TMEM over the batch)
    xq_e = c["xq"]                            # operands of the weight-gradient MMAs: the projections' copies
    xkv_e = c["xkv_p"]
    dM = torch.einsum("brk,brj->kj", xq_e, dZ)        # dM[k][k']
    du = dZ.sum((0, 1))
    dWv = torch.einsum("brn,brk->nk", dV, xkv_e)
    dbv = dCb.sum((0, 1)) if c["vb_out"] else dV.sum((0, 1))
    s = c["s"]
    # finalize: M = s Wq^T Wk, u = s Wk^T bq   (fp32)
    grads[f"{prefix}.W_query.weight"] = s * (wk @ dM.T)                      # dWq[n][k] = s sum_k' dM[k][k'] Wk[n][k']
    grads[f"{prefix}.W_key.weight"] = s * (wq @ dM + torch.outer(bq, du))   # dWk[n][k'] = s (sum_k Wq[n][k] dM[k][k'] + bq[n] du[k'])
    grads[f"{prefix}.W_query.bias"] = s * (wk @ du)
    grads[f"{prefix}.W_key.bias"] = torch.zeros_like(bq)
    grads[f"{prefix}.W_value.weight"] = dWv
    grads[f"{prefix}.W_value.bias"] = dbv
    if not want_dx:
        return None, None
    Wv_b = R(f"{blk}.Wv_bwd", p[f"{prefix}.W_value.weight"]) if f"{blk}.Wv_bwd" in R else c["Wv"]
    dXq = R(f"{blk}.dXq", dZ @ c["M"].T)
    dXkv = R(f"{blk}.dXkv", dS.transpose(1, 2) @ c["Z"] + dV @ Wv_b)
    return dXq, dXkv


def emulate(p, img, txt, labels, reverse, co, R, drop_mask=None, drop_scale=1.0):
    p = {k: v.double() for k, v in p.items()}
    img, txt = img.double(), txt.double()
    B = img.shape[0]
    img_n = img / img.norm(dim=1, keepdim=True)
    txt_n = txt / txt.norm(dim=1, keepdim=True)
    xi = R("X", img_n).reshape(B, L, -1)
    xt = R("X", txt_n).reshape(B, L, -1)
    i_sa, c_i = attn_fwd(R, "sa", xi, xi, p, orc.SA_IMAGE, False)
    t_sa, c_t = attn_fwd(R, "sa", xt, xt, p, orc.SA_TEXT, False)
    i_sa_b, t_sa_b = R("sa.out", i_sa), R("sa.out", t_sa)
    t_i, c_1 = attn_fwd(R, "ca", t_sa_b, i_sa_b, p, orc.CA_1, reverse)
    i_t, c_2 = attn_fwd(R, "ca", i_sa_b, t_sa_b, p, orc.CA_2, reverse)
    fin = orc.final_linear_name(False, co)
    wf, bf = p[f"{fin}.weight"], p[f"{fin}.bias"]
    D = wf.shape[1]
    m = torch.ones(B, D, dtype=torch.float64) if drop_mask is None else drop_mask.double() * drop_scale
    ca_w = 48 * L
    F1 = R("ca.F", t_i.reshape(B, -1) * m[:, :ca_w])
    F2 = R("ca.F", i_t.reshape(B, -1) * m[:, ca_w:2 * ca_w])
    wfb = R("Wf", wf[:, :2 * ca_w])
    logits = bf + F1 @ wfb[:, :ca_w].T + F2 @ wfb[:, ca_w:].T
    if not co:
        feat = torch.cat((img_n, txt_n), 1) * m[:, 2 * ca_w:]       # fp32 features, fp32 weights
        logits = logits + feat @ wf[:, 2 * ca_w:].T
    loss, dl = orc.np_cross_entropy_fwd_bwd(logits.numpy(), labels.numpy())
    dl = torch.from_numpy(dl)
    grads = {}
    gwf = torch.zeros_like(wf)
    grads[f"{fin}.bias"] = dl.sum(0)
    if not co:
        xb = torch.cat((xi.reshape(B, -1), xt.reshape(B, -1)), 1) * m[:, 2 * ca_w:]      # ce_feat reads the bf16 X images
        gwf[:, 2 * ca_w:] = dl.T @ xb
    dlb = R("DL", dl)
    # CA backward: Out operand (relu*mask in bf16) for dWf; dOut = DL Wf^T
    gwf[:, :ca_w] = dlb.T @ F1
    gwf[:, ca_w:2 * ca_w] = dlb.T @ F2
    grads[f"{fin}.weight"] = gwf
    wfb2 = R("Wf_bwd", wf[:, :2 * ca_w]) if "Wf_bwd" in R else wfb
    dO1 = (dlb @ wfb2[:, :ca_w]).reshape(B, L, 48)
    dO2 = (dlb @ wfb2[:, ca_w:]).reshape(B, L, 48)
    dq1, dkv1 = attn_bwd(R, "ca", dO1, c_1, p, orc.CA_1, grads, dy_mask=m[:, :ca_w].reshape(B, L, 48))
    dq2, dkv2 = attn_bwd(R, "ca", dO2, c_2, p, orc.CA_2, grads, dy_mask=m[:, ca_w:2 * ca_w].reshape(B, L, 48))
    d_t_sa = dq1 + dkv2
    d_i_sa = dkv1 + dq2
    attn_bwd(R, "sa", d_t_sa, c_t, p, orc.SA_TEXT, grads, want_dx=False, p_for_softmax="bf16")
    attn_bwd(R, "sa", d_i_sa, c_i, p, orc.SA_IMAGE, grads, want_dx=False, p_for_softmax="bf16")
    return logits.numpy(), loss, {k: v.numpy() for k, v in grads.items()}


def summarize(ours, ref, names):
    worst, wname, worst_max, wmname = 0.0, "", 0.0, ""
    fo, fr = [], []
    for n in names:
        r, o = np.asarray(ref[n], dtype=np.float64), ours[n]
        fo.append(o.ravel()); fr.append(r.ravel())
        if np.linalg.norm(r) < 1e-12:
            continue
        e = np.linalg.norm(o - r) / np.linalg.norm(r)
        em = np.abs(o - r).max() / np.abs(r).max()
        if e > worst:
            worst, wname = e, n
        if em > worst_max:
            worst_max, wmname = em, n
    fo, fr = np.concatenate(fo), np.concatenate(fr)
    return dict(flat=np.linalg.norm(fo - fr) / np.linalg.norm(fr), worst_l2=worst, worst_l2_name=wname,
                worst_max=worst_max, worst_max_name=wmname)


def per_tensor(ours, ref, names):
    out = {}
    for n in names:
        r, o = np.asarray(ref[n], dtype=np.float64), ours[n]
        if np.linalg.norm(r) < 1e-12:
            continue
        out[n] = (np.linalg.norm(o - r) / np.linalg.norm(r), np.abs(o - r).max() / np.abs(r).max())
    return out


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 31
    verbose = len(sys.argv) > 3
    from tests._util import make_inputs
    configs = [(1.0, True, False), (1.0, True, True), (40.0, True, False), (40.0, True, True)]
    all_sites = ["X", "sa.M", "sa.Wv", "sa.Z", "sa.V", "sa.P", "sa.out", "ca.M", "ca.Wv", "ca.Z", "ca.V", "ca.P", "ca.F",
                 "Wf", "DL", "ca.dyx", "ca.dy", "ca.dC", "ca.dV", "ca.dS", "ca.dZ", "ca.dXq", "ca.dXkv",
                 "sa.dyx", "sa.dy", "sa.dC", "sa.dV", "sa.dS", "sa.dZ"]
    for qk, rev, co in configs:
        p = orc.init_head_params(cross_attention_only=co, seed=seed, qk_gain=qk)
        img, txt, labels = make_inputs(B, seed)
        ref = orc.np_head_forward_backward(p, img.numpy(), txt.numpy(), rev, False, co, labels=labels.numpy())
        names = orc.head_param_names(False, co)
        names = [n for n in names if not n.endswith("W_key.bias")]
        exact = emulate(p, img, txt, labels, rev, co, Sites({"*": "off"}))
        s0 = summarize(exact[2], ref["grads"], names)
        base = emulate(p, img, txt, labels, rev, co, Sites({"*": "bf16"}))
        sb = summarize(base[2], ref["grads"], names)
        print(f"== qk {qk} reverse {rev} cross_only {co} B {B}: exact-emulation flat {s0['flat']:.1e}; all-bf16: flat {sb['flat']:.3e} "
              f"worst l2 {sb['worst_l2']:.3e} ({sb['worst_l2_name']}) worst max {sb['worst_max']:.3e} ({sb['worst_max_name']}) "
              f"logits err {np.abs(base[0] - ref['logits']).max():.2e}")
        if verbose:
            for n, (e, em) in per_tensor(base[2], ref["grads"], names).items():
                print(f"     {n:44s} l2 {e:.3e} max {em:.3e}")
        # one site at a time on (everything else exact): which sites produce the error
        rows = []
        for site in all_sites:
            o = emulate(p, img, txt, labels, rev, co, Sites({"*": "off", site: "bf16"}))
            s = summarize(o[2], ref["grads"], names)
            rows.append((s["worst_l2"], site, s))
        rows.sort(reverse=True)
        for w, site, s in rows[:12]:
            print(f"   only {site:8s}: flat {s['flat']:.3e} worst l2 {s['worst_l2']:.3e} ({s['worst_l2_name']}) worst max {s['worst_max']:.3e}")


if __name__ == "__main__":
    main()
