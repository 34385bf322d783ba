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
