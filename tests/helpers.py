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
