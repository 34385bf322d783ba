"""Greedy decoding of the attention decoders (rnn_attn.py:77-94,120-145): every step recomputes
the attention on the full batch from the pre-step top-layer hidden state, feeds
[emb(prev token) | embed(ctx)] through the stacked GRU/LSTM and takes the arg-max of the
vocabulary projection.  Same kernels as training, single steps, fp32."""
import torch

from . import _lib, ops
from .engine import layer_params

F32, I64 = torch.float32, torch.int64


def greedy(mod, feature, start_id, max_len):
    if not feature.is_cuda:
        raise RuntimeError("showtell_b200 runs on CUDA tensors only (no CPU fallback)")
    from .attn_engine import grid_layout
    lay = grid_layout(mod)                                 # "BCP" channels-first (the reference) / "BPC" channels-last
    if feature.dim() != 3 or feature.shape[2 if lay == "BPC" else 1] != mod.nos_filters:
        raise ValueError(f"cnn_feature must be (B, {mod.nos_filters}, P)" if lay == "BCP" else
                         f"cnn_feature must be (B, P, {mod.nos_filters})")
    P = {n: p.detach() for n, p in mod.named_parameters()}
    kind, L = mod._kind, mod.num_layers
    f = feature.detach().contiguous().to(F32)
    E, H = mod.embed_dim, mod.num_hidden_units
    dev = f.device
    if lay == "BPC":
        B, Pn, C = f.shape
        F, mean_f = ops.attn_grid_bpc(f, bf16=False)
    else:
        B, C, Pn = f.shape
        F, _, mean_f = ops.attn_relayout(f, bf16=False, want_t=False)
    h = [ops.sgemm(mean_f, P["init_h.weight"], transB=True, bias=P["init_h.bias"])] * L
    c = [ops.sgemm(mean_f, P["init_c.weight"], transB=True, bias=P["init_c.bias"])] * L if kind == _lib.ST_LSTM \
        else [None] * L
    att1 = ops.sgemm(F, P["attn.encoder_att.weight"], transB=True, bias=P["attn.encoder_att.bias"])
    Fe = ops.sgemm(F, P["embed.weight"], transB=True)
    wf = P["attn.full_att.weight"].reshape(-1)
    Wih0, _, bih0, _ = layer_params(P, 0)
    tokens = torch.empty(B, max_len, dtype=I64, device=dev)
    tok = torch.full((B,), int(start_id), dtype=I64, device=dev)                   # rnn_attn.py:127-128
    X = torch.empty(B, 2 * E, dtype=F32, device=dev)
    alpha = torch.empty(B, Pn, dtype=F32, device=dev)
    S = torch.zeros(B, Pn, dtype=F32, device=dev)
    bs1 = [B]
    for t in range(max_len):
        att2 = ops.sgemm(h[L - 1], P["attn.decoder_att.weight"], transB=True, bias=P["attn.decoder_att.bias"])
        ops.attn_step_fwd(B, Pn, att1, Fe, att2, wf, P["attn.full_att.bias"], P["embed.bias"], alpha, Pn, S,
                          X[:, E:])
        ops.gather_rows(X, P["embeddings.weight"], tok)                              # rnn_attn.py:129,91
        inp = X
        for l in range(L):
            Wih, Whh, bih, bhh = layer_params(P, l)
            gx = ops.sgemm(inp, Wih, transB=True, bias=bih)
            o = ops.rnn_seq_fwd(kind, gx, Whh, bhh, bs1, h0=h[l], c0=c[l], save=False)
            h[l], c[l] = o["Hs"], o["Cs"]
            inp = h[l]
        logits = ops.sgemm(h[L - 1], P["linear.weight"], transB=True, bias=P["linear.bias"])
        tok = ops.argmax_rows(logits, out=tokens[:, t])                              # rnn_attn.py:88
    return tokens
