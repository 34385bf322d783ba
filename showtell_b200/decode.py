"""Host wrappers of the decoding entry points (greedy / chain beam / tree beam)."""
import ctypes as C

import torch

from . import _lib
from ._lib import RnnWeights, check, ptr, stream_ptr

F32, I64, I32 = torch.float32, torch.int64, torch.int32


# nn.Linear products inside the decoding loops: "fp32" = CUDA cores (the arithmetic token ids are defined
# against), "tf32x3" = tensor cores with hi/lo-split operands (fp32-accurate, ~5x the throughput).
GEMM_MODES = {"fp32": 0, "tf32x3": 1}


class WeightPack:
    """Keeps the ctypes pointer arrays (host memory) alive for the duration of a call."""

    def __init__(self, mod):
        L = mod.num_layers
        sd = {n: p.detach() for n, p in mod.named_parameters()}
        for t in sd.values():
            if not t.is_cuda or t.dtype != F32:
                raise RuntimeError("decoding needs fp32 CUDA parameters (no CPU fallback)")
        self.tensors = sd
        arr = lambda fmt: (C.c_void_p * L)(*[sd[fmt.format(l)].contiguous().data_ptr() for l in range(L)])
        self.wih, self.whh = arr("unit.weight_ih_l{}"), arr("unit.weight_hh_l{}")
        self.bih, self.bhh = arr("unit.bias_ih_l{}"), arr("unit.bias_hh_l{}")
        self.struct = RnnWeights(mod._kind, L, mod.embed_dim, mod.num_hidden_units, mod.vocab_size,
                                 sd["embeddings.weight"].data_ptr(), self.wih, self.whh, self.bih, self.bhh,
                                 sd["linear.weight"].data_ptr(), sd["linear.bias"].data_ptr(),
                                 GEMM_MODES[getattr(mod, "decode_gemm", "fp32")])

    def workspace(self, n_img, K, max_len, device):
        nbytes = _lib.load().st_decode_workspace_bytes(C.byref(self.struct), n_img, K, max_len)
        if nbytes < 0:
            raise ValueError("st_decode_workspace_bytes: " + _lib.last_error())
        return torch.empty(nbytes, dtype=torch.uint8, device=device), nbytes


def _feature(feature, E):
    if not feature.is_cuda:
        raise RuntimeError("showtell_b200 runs on CUDA tensors only (no CPU fallback)")
    if feature.dim() != 2 or feature.shape[1] != E:
        raise ValueError(f"cnn_feature must be (B, {E})")
    return feature.detach().contiguous().to(F32)


def greedy(mod, feature, max_len=25):
    """RNN.sentence_index(cnn_feature) -> (B, max_len) int64."""
    lib = _lib.load()
    f = _feature(feature, mod.embed_dim)
    wp = WeightPack(mod)
    n = f.shape[0]
    ws, nb = wp.workspace(n, 1, max_len, f.device)
    tokens = torch.empty(n, max_len, dtype=I64, device=f.device)
    check(lib.st_decode_greedy(C.byref(wp.struct), ptr(f), n, max_len, ptr(tokens), ptr(ws), nb,
                               stream_ptr()), "st_decode_greedy")
    return tokens


def beam_chain(mod, feature, K, max_len=25, trace=False):
    """RNN.sentence_index(cnn_feature, beam_size=K), batched over images -> (B, max_len) int64.
    With trace=True also returns (scores (max_len,B,K), words (max_len,B,K))."""
    lib = _lib.load()
    f = _feature(feature, mod.embed_dim)
    wp = WeightPack(mod)
    n = f.shape[0]
    ws, nb = wp.workspace(n, K, max_len, f.device)
    tokens = torch.empty(n, max_len, dtype=I64, device=f.device)
    ts = torch.zeros(max_len, n, K, dtype=F32, device=f.device) if trace else None
    tw = torch.zeros(max_len, n, K, dtype=I32, device=f.device) if trace else None
    check(lib.st_decode_beam_chain(C.byref(wp.struct), ptr(f), n, int(K), max_len, ptr(tokens), ptr(ts),
                                   ptr(tw), ptr(ws), nb, stream_ptr()), "st_decode_beam_chain")
    return (tokens, ts, tw) if trace else tokens


def beam_tree(mod, feature, start_id, end_id, beam_width=4, num_hypotheses=1, max_length=50):
    """beam_search.beam_search() semantics, batched over images.
    Returns (tokens (B, num_hyp, max_length+1) int32 padded with -1, lengths (B, num_hyp), costs)."""
    lib = _lib.load()
    f = _feature(feature, mod.embed_dim)
    wp = WeightPack(mod)
    n = f.shape[0]
    ws, nb = wp.workspace(n, max(beam_width, 1), max_length, f.device)
    tok = torch.empty(n, num_hypotheses, max_length + 1, dtype=I32, device=f.device)
    ln = torch.empty(n, num_hypotheses, dtype=I32, device=f.device)
    cost = torch.empty(n, num_hypotheses, dtype=F32, device=f.device)
    check(lib.st_decode_beam_tree(C.byref(wp.struct), ptr(f), n, int(beam_width), int(num_hypotheses),
                                  int(max_length), int(start_id), int(end_id), ptr(tok), ptr(ln), ptr(cost),
                                  ptr(ws), nb, stream_ptr()), "st_decode_beam_tree")
    return tok, ln, cost
