"""Host-side orchestration of the base (non-attention) decoder: which kernels run in which order
for RNN.forward / forward_loss / backward.  Mirrors the data flow of rnn.py:27-35 + main.py:145-151
with the time-invariant input projection hoisted out of the recurrence.

fp32 mode: CUDA-core GEMMs (st_sgemm), persistent recurrent kernels, materialised logits.
"""
import torch

from . import _lib, ops

F32 = torch.float32


def layer_params(P, l):
    return (P[f"unit.weight_ih_l{l}"], P[f"unit.weight_hh_l{l}"],
            P[f"unit.bias_ih_l{l}"], P[f"unit.bias_hh_l{l}"])


def stack_forward(P, kind, L, X, bs, save):
    """All L recurrent layers over the packed input X (N, in).  Layer l's input projection is one
    hoisted GEMM over every time step; the recurrence is one persistent kernel per layer.
    Returns (top-layer Hs (N,H), per-layer saved state)."""
    layers, inp = [], X
    for l in range(L):
        Wih, Whh, bih, bhh = layer_params(P, l)
        Gx = ops.sgemm(inp, Wih, transB=True, bias=bih, tag="ih_fwd")      # W_ih x + b_ih, all steps
        o = ops.rnn_seq_fwd(kind, Gx, Whh, bhh, bs, save=save, tag="seq_fwd")
        layers.append({"inp": inp, "out": o})
        inp = o["Hs"]
    return inp, layers


def stack_backward(P, kind, L, bs, layers, dHs_top, grads):
    """BPTT through the L layers, top down.  Fills grads[...] for the unit.* parameters and returns
    the gradient w.r.t. the packed layer-0 input (N, in_0)."""
    dH = dHs_top
    for l in reversed(range(L)):
        Wih, Whh, bih, bhh = layer_params(P, l)
        sv = layers[l]
        b = ops.rnn_seq_bwd(kind, Whh, bs, sv["out"], dH, tag="seq_bwd")
        Hprev = ops.shift_states(sv["out"]["Hs"], bs)
        grads[f"unit.weight_hh_l{l}"] = ops.sgemm(b["dGh"], Hprev, transA=True, tag="hh_dw")   # dGh^T Hprev
        grads[f"unit.bias_hh_l{l}"] = ops.colsum(b["dGh"])
        grads[f"unit.weight_ih_l{l}"] = ops.sgemm(b["dG"], sv["inp"], transA=True, tag="ih_dw")  # dG^T X
        grads[f"unit.bias_ih_l{l}"] = ops.colsum(b["dG"])
        dH = ops.sgemm(b["dG"], Wih, tag="ih_dx")                                     # dX = dG W_ih
    return dH


def base_forward(P, kind, L, feature, caption, bs, save):
    X = ops.pack_inputs(P["embeddings.weight"], feature, caption, bs, True)           # rnn.py:29-31
    Hs, layers = stack_forward(P, kind, L, X, bs, save)
    return Hs, layers


def vocab_logits(P, Hs):
    return ops.sgemm(Hs, P["linear.weight"], transB=True, bias=P["linear.bias"], tag="vocab_fwd")  # rnn.py:33


def base_backward(P, kind, L, caption, bs, layers, Hs_top, dlogits, want_dfeature, feature_shape):
    """Gradients of everything given dlogits (N, V).  Returns (grads dict, dfeature or None)."""
    grads = {}
    Wv = P["linear.weight"]
    grads["linear.weight"] = ops.sgemm(dlogits, Hs_top, transA=True, tag="vocab_dw")  # dlogits^T Hs
    grads["linear.bias"] = ops.colsum(dlogits)
    dHs = ops.sgemm(dlogits, Wv, tag="vocab_dh")                                      # dlogits W_v
    dX = stack_backward(P, kind, L, bs, layers, dHs, grads)
    dEmb = torch.zeros_like(P["embeddings.weight"])
    dfeat = torch.empty(feature_shape, dtype=F32, device=dX.device) if want_dfeature else None
    ops.pack_inputs_bwd(dX, dEmb, dfeat, caption, bs, True)
    grads["embeddings.weight"] = dEmb
    return grads, dfeat


def _check_inputs(feature, caption, lengths, E):
    if not (feature.is_cuda and caption.is_cuda):
        raise RuntimeError("showtell_b200 runs on CUDA tensors only (no CPU fallback)")
    if feature.dim() != 2 or feature.shape[1] != E:
        raise ValueError(f"cnn_feature must be (B, {E}), got {tuple(feature.shape)}")
    if caption.dim() != 2 or caption.shape[0] != feature.shape[0] or caption.dtype != torch.int64:
        raise ValueError("image_caption must be (B, T) int64")
    if len(lengths) != feature.shape[0]:
        raise ValueError("caption_size must have one entry per batch row")
    bs = _lib.batch_sizes(lengths)
    if len(bs) > caption.shape[1] + 1:
        raise ValueError("caption_size exceeds the padded caption length + 1")
    return bs


class BaseLogitsFn(torch.autograd.Function):
    """RNN.forward as the reference exposes it: returns the (N, V) logits; gradients flow to
    cnn_feature and every parameter."""

    @staticmethod
    def forward(ctx, mod, feature, caption, lengths, *params):
        names = [n for n, _ in mod.named_parameters()]
        P = {n: p.detach() for n, p in zip(names, params)}
        feature_c = feature.detach().contiguous().to(F32)
        caption_c = caption.contiguous()
        bs = _check_inputs(feature_c, caption_c, lengths, mod.embed_dim)
        save = any(ctx.needs_input_grad)
        Hs, layers = base_forward(P, mod._kind, mod.num_layers, feature_c, caption_c, bs, save)
        logits = vocab_logits(P, Hs)
        ctx.names, ctx.P, ctx.mod = names, P, mod
        ctx.bs, ctx.layers, ctx.Hs, ctx.caption = bs, layers, Hs, caption_c
        ctx.feature_shape = feature_c.shape
        return logits

    @staticmethod
    def backward(ctx, dlogits):
        mod = ctx.mod
        dlogits = dlogits.contiguous().to(F32)
        grads, dfeat = base_backward(ctx.P, mod._kind, mod.num_layers, ctx.caption, ctx.bs, ctx.layers,
                                     ctx.Hs, dlogits, ctx.needs_input_grad[1], ctx.feature_shape)
        return (None, dfeat, None, None) + tuple(grads[n] for n in ctx.names)


class BaseLossFn(torch.autograd.Function):
    """forward_loss: mean cross-entropy over the packed tokens (main.py:145,149), with forward and
    backward run back to back so the (N, V) logits buffer is turned into its own gradient in place
    and released before the call returns.  `denom` = number of tokens the mean runs over (the
    global token count under data parallelism)."""

    @staticmethod
    def forward(ctx, mod, feature, caption, lengths, denom, *params):
        names = [n for n, _ in mod.named_parameters()]
        P = {n: p.detach() for n, p in zip(names, params)}
        feature_c = feature.detach().contiguous().to(F32)
        caption_c = caption.contiguous()
        bs = _check_inputs(feature_c, caption_c, lengths, mod.embed_dim)
        if len(bs) > caption_c.shape[1]:
            raise ValueError("caption_size exceeds the padded caption length")
        need = any(ctx.needs_input_grad)
        N = sum(bs)
        denom = float(denom if denom is not None else N)
        Hs, layers = base_forward(P, mod._kind, mod.num_layers, feature_c, caption_c, bs, need)
        logits = vocab_logits(P, Hs)
        target = ops.pack_targets(caption_c, bs)
        loss_sum, _, dl = ops.ce_fwd_bwd(logits, target, grad_scale=(1.0 / denom) if need else None,
                                         inplace=True)
        ctx.names = names
        ctx.grads, ctx.dfeat = None, None
        if need:
            ctx.grads, ctx.dfeat = base_backward(P, mod._kind, mod.num_layers, caption_c, bs, layers, Hs, dl,
                                                 ctx.needs_input_grad[1], feature_c.shape)
        return (loss_sum / denom).reshape(())

    @staticmethod
    def backward(ctx, g):
        if ctx.grads is None:
            raise RuntimeError("forward_loss was run without grad enabled")
        dfeat = ctx.dfeat * g if (ctx.dfeat is not None and ctx.needs_input_grad[1]) else None
        return (None, dfeat, None, None, None) + tuple(ctx.grads[n] * g for n in ctx.names)
