"""Shared helpers for the parity tests (oracle side lives in oracle/, see its header)."""
import os

import numpy as np
import torch

from conftest import GOLDEN

BASE_CASES = ["gru_l1", "gru_tiny", "gru_l2", "lstm_l1", "lstm_l3", "gru_med", "lstm_med"]
ATTN_CASES = ["attn_gru_l1", "attn_gru_l2", "attn_lstm_l1", "attn_lstm_l2", "attn_gru_med", "attn_lstm_med"]


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    return {k: z[k] for k in z.files}


def golden_params(g, dtype=torch.float32, device="cpu"):
    return {k[len("param."):]: torch.from_numpy(v).to(dtype).to(device) for k, v in g.items()
            if k.startswith("param.")}


def golden_grads(g):
    return {k[len("grad."):]: torch.from_numpy(v) for k, v in g.items() if k.startswith("grad.")}


def rel_err(a, b):
    """max |a-b| / max(|b|_inf, tiny) -- the 'relative' of north_star's 1e-4 / 2e-2 bars."""
    a = torch.as_tensor(a, dtype=torch.float64).cpu()
    b = torch.as_tensor(b, dtype=torch.float64).cpu()
    return float((a - b).abs().max() / max(float(b.abs().max()), 1e-30))


KINKED = ("attn.encoder_att.weight", "attn.encoder_att.bias", "attn.decoder_att.weight", "attn.decoder_att.bias")


def l2_err(a, b):
    a = torch.as_tensor(a, dtype=torch.float64).cpu()
    b = torch.as_tensor(b, dtype=torch.float64).cpu()
    return float((a - b).norm() / max(float(b.norm()), 1e-30))


def reference_grads_on_gpu(model, p, feat, cap, lengths, alpha_c=1.0, autocast=None, dtype=torch.float32):
    """Gradients of the reference algorithm run by torch's own CUDA kernels on this GPU -- the UNMODIFIED reference
    modules from baseline/_ref when they are installed, else the oracle's explicit equations -- in `dtype`, optionally
    under torch.autocast(`autocast`).  A second, independent implementation of the same arithmetic: how far IT lands
    from the CPU truth is the reference's own spread at that size, the yardstick for tensors whose gradient is
    discontinuous in the pre-activation (LeakyReLU kink, Attention/rnn_attn.py:18,25)."""
    from oracle import showtell_oracle as O
    dev = torch.device("cuda:0")
    ctx = torch.autocast("cuda", dtype=autocast) if autocast is not None else torch.autocast("cuda", enabled=False)
    prev = torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = torch.backends.cudnn.allow_tf32 = False
    try:
        try:
            from baseline import reference as R
            have = R.available()
        except Exception:
            have = False
        if have:
            E, H = p["embeddings.weight"].shape[1], p["unit.weight_hh_l0"].shape[1]
            V = p["linear.weight"].shape[0]
            L = sum(1 for k in p if k.startswith("unit.weight_hh_l"))
            Cn = p["embed.weight"].shape[1] if "embed.weight" in p else 2048
            An = p["attn.encoder_att.weight"].shape[0] if "attn.encoder_att.weight" in p else 512
            net = R.make_module(model, E, H, V, L, Cn, An)
            net.load_state_dict(p)
            net = net.to(dev).to(dtype)
            with ctx:
                R.train_step(net, model, feat.to(dev).to(dtype), cap.to(dev), lengths, alpha_c)
            return {n: q.grad.detach().float().cpu() for n, q in net.named_parameters()}, "reference modules"
        pg = {k: v.to(dev).to(dtype) for k, v in p.items()}
        with ctx:
            _, grads, _ = O.train_step(pg, model, feat.to(dev).to(dtype), cap.to(dev), lengths, alpha_c)
        return {k: v.float().cpu() for k, v in grads.items()}, "oracle equations"
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = prev


def kink_ambiguity(p, model, feat, cap, lengths, alpha_c=1.0, eps=1e-5):
    """float64 oracle gradients + a componentwise bound on what the LeakyReLU kink can move in the four
    attention-projection gradients when the pre-activation s = att1 + att2 (Attention/rnn_attn.py:25) is only known to
    +-eps (fp32 rounding of a 2048-long dot product is ~1e-6).  LeakyReLU'(s) is 1 or 0.2; an implementation whose s
    differs in the last bits may land on the other side for the entries with |s| <= eps, and each such entry
    (t, b, p, a) moves d s by 0.8 * dL/du[t,b,p,a], i.e.
        d encoder_att.weight[a, :] by 0.8 |dL/du| |f[b,p,:]|,   d encoder_att.bias[a] by 0.8 |dL/du|,
        d decoder_att.weight[a, :] by 0.8 |dL/du| |h_{t-1}[b,:]|, d decoder_att.bias[a] by 0.8 |dL/du|.
    Entries of those tensors that no ambiguous term touches get a zero bound, i.e. the plain bar.
    Returns (grads64, bounds {name: tensor}, number of ambiguous entries)."""
    from oracle import showtell_oracle as O
    p64 = {k: v.double() for k, v in p.items()}
    f64 = feat.double()
    rec = []
    orig_leaky, orig_att = O.leaky_relu02, O.attention

    def leaky(x):
        u = orig_leaky(x)
        u.retain_grad()
        rec[-1]["s"], rec[-1]["u"] = x.detach(), u
        return u

    def att(pp, feat_bpc, h):
        rec.append({"h": h.detach()})
        return orig_att(pp, feat_bpc, h)

    O.leaky_relu02, O.attention = leaky, att
    try:
        _, g64, _ = O.train_step(p64, model, f64, cap, lengths, alpha_c)
    finally:
        O.leaky_relu02, O.attention = orig_leaky, orig_att
    A, C = p["attn.encoder_att.weight"].shape
    H = p["attn.decoder_att.weight"].shape[1]
    bw_e, bb = torch.zeros(A, C, dtype=torch.float64), torch.zeros(A, dtype=torch.float64)
    bw_d = torch.zeros(A, H, dtype=torch.float64)
    fbpc = f64.transpose(1, 2).abs()
    n_amb = 0
    for r in rec:
        idx = (r["s"].abs() <= eps).nonzero()
        if idx.numel() == 0:
            continue
        n_amb += idx.shape[0]
        b, pp_, a = idx[:, 0], idx[:, 1], idx[:, 2]
        mag = 0.8 * r["u"].grad[b, pp_, a].abs()
        bb.index_add_(0, a, mag)
        bw_e.index_add_(0, a, mag[:, None] * fbpc[b, pp_])
        bw_d.index_add_(0, a, mag[:, None] * r["h"][b].abs())
    bounds = {"attn.encoder_att.weight": bw_e, "attn.encoder_att.bias": bb,
              "attn.decoder_att.weight": bw_d, "attn.decoder_att.bias": bb.clone()}
    return g64, bounds, n_amb


def assert_close_with_kink_bound(name, got, ref64, bound, tol):
    """|got - ref| <= tol * max|ref| + bound, componentwise."""
    got = torch.as_tensor(got, dtype=torch.float64).cpu()
    slack = tol * float(ref64.abs().max()) + (bound if bound is not None else 0.0)
    excess = ((got - ref64).abs() - slack).max()
    assert float(excess) <= 0.0, (name, float(excess), float((got - ref64).abs().max() / ref64.abs().max()))
