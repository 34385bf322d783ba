"""Host-side orchestration of the attention decoders (Attention/rnn_attn.py, rnn_attn_LSTM.py +
the loss of main_attn.py:126-131).

Reference per step t (rnn_attn.py:66-74): attention over the full grid with a freshly recomputed
encoder projection, embed(context), one GRU/LSTM step on [emb_t | embed(ctx)], vocabulary
projection.  Here (see csrc/attn.cu): the encoder projection att1 and the embedded grid Fe are
hoisted GEMMs, each step is {small GEMM for decoder_att(h), one fused attention kernel, small GEMM
for the context half of W_ih, one recurrent step per layer}, and the vocabulary projection + CE
runs once over all packed top-layer states after the loop (the reference's `op` is never fed back,
rnn_attn.py:71-72).  Backward mirrors it, with d att1 / d w_f / d W_embed / all weight gradients
hoisted out of the reverse loop.
"""
import torch

from . import _lib, graphs, ops
from .engine import Linear, layer_params, vocab_ce, weight_grad

F32, BF16 = torch.float32, torch.bfloat16


def _offsets(bs):
    off, o = [], 0
    for b in bs:
        off.append(o)
        o += b
    return off


def _lin_small(x, W, bias=None, out=None, beta=0.0):
    """Per-step product on fp32 operands (M ~ live batch rows)."""
    return ops.sgemm(x, W, transB=True, bias=bias, out=out, beta=beta)


def _use_tc(mode, kind, P, B):
    """bf16 mode runs the per-step products and the recurrence on the tensor-core kernels when the
    shapes qualify (TMA operands need 16-byte rows; the persistent recurrent grid must be co-resident)."""
    if mode != "bf16":
        return False
    H = P["unit.weight_hh_l0"].shape[1]
    E = P["embeddings.weight"].shape[1]
    A = P["attn.decoder_att.weight"].shape[0]
    return (E % 8 == 0 and A % 8 == 0 and H % 8 == 0 and ops.rnn_seq_tc_fits(kind, H, B))


def grid_operands(mode, feature, layout="BCP", out=None):
    """The decoder's view of the feature grid: F (B*P, C) rows in the compute storage type + the channel means (B, C).
    layout "BCP" = the reference's channels-first grid (cnn_attn.py:49: one re-layout pass), "BPC" = channels-last
    (the grid IS the operand).  out = (F, mean_f): write into these (graphs.run's static operands)."""
    # (no transposed copy of the grid: dW_enc = datt1^T F reads F in place as an MN-major GEMM operand)
    if layout == "BPC":
        return ops.attn_grid_bpc(feature, bf16=(mode == "bf16"), out=out)
    F, _, mean_f = ops.attn_relayout(feature, bf16=(mode == "bf16"), want_t=False, out=out)
    return F, mean_f


def attn_forward(mode, P, kind, L, feature, caption, bs, save, layout="BCP", grid=None):
    """Returns (Hs_top (N,H), alphas (B,Tcap,P), saved dict).  layout: see grid_operands.  grid = (F, mean_f, Pn): the
    operands grid_operands already produced (`feature` is then unused)."""
    if grid is not None:
        F, mean_f, Pn = grid
        B, C = mean_f.shape
    else:
        if layout == "BPC":
            B, Pn, C = feature.shape
        else:
            B, C, Pn = feature.shape
        F, mean_f = grid_operands(mode, feature, layout)
    FT = None
    T, N, off = len(bs), sum(bs), _offsets(bs)
    Tcap = caption.shape[1]
    E = P["embeddings.weight"].shape[1]
    dev = F.device
    tc = _use_tc(mode, kind, P, B)
    sv = {"bs": bs, "off": off, "Pn": Pn, "B": B, "tc": tc}

    sv.update(F=F, mean_f=mean_f)
    # rnn_attn.py:62: every layer starts from init_h(mean_P f) (init_c likewise for the LSTM)
    # (a skinny fp32 product: it runs on the side stream beside the hoisted grid projections below)
    if tc and C % 8 == 0:
        # tensor cores: (B, C) x (H, C)^T with the bf16 mean (an fp32 CUDA-core product here took 230 us at B = 128,
        # C = 2048 and sat on the critical path in front of the time loop)
        mean_b, _ = ops.cast_bf16(mean_f, True, False)
        sv["mean_b"] = mean_b
        init = lambda name: ops.gemm_bf16(mean_b, ops.bf16_shadow(P[name + ".weight"]), bias=P[name + ".bias"], tag="init_fwd")
    else:
        init = lambda name: _lin_small(mean_f, P[name + ".weight"], P[name + ".bias"])
    (h0, c0), init_done = ops.fork(lambda: (init("init_h"), init("init_c") if kind == _lib.ST_LSTM else None),
                                   uses=(mean_f,))
    sv.update(h0=h0, c0=c0)
    # hoisted, state-independent projections of the grid
    store = BF16 if mode == "bf16" else F32
    enc = Linear(mode, P["attn.encoder_att.weight"], P["attn.encoder_att.bias"], save, "att1")
    att1 = enc.fwd(F, FT, out_dtype=store)                                            # rnn_attn.py:23
    emb_lin = Linear(mode, P["embed.weight"], None, False, "fe")
    Fe = emb_lin.fwd(F, None, out_dtype=store)                                        # embed() applied per location
    sv.update(att1=att1, Fe=Fe)
    # layer-0 input rows [emb(caption[:, t]) | embed(ctx_t)]  (rnn_attn.py:70: no input/target shift)
    X0 = ops.pack_inputs(P["embeddings.weight"], None, caption, bs, False, width=2 * E)
    Wih0, _, bih0, _ = layer_params(P, 0)
    H = P["unit.weight_hh_l0"].shape[1]
    G = Wih0.shape[0]
    alphas = torch.zeros(B, Tcap, Pn, dtype=F32, device=dev)                          # rnn_attn.py:65
    S = torch.zeros(B, Pn, dtype=F32, device=dev)
    A = P["attn.decoder_att.weight"].shape[0]
    att2_all = torch.empty(N, A, dtype=F32, device=dev)
    outs = [None] * L
    wf = P["attn.full_att.weight"].reshape(-1)
    bd = P["attn.decoder_att.bias"]
    if tc:
        # bf16 shadows of the weights (cached until the optimizer changes them); the transposes are the K-major
        # TMA operands of the reverse step kernels -- the GEMMs read either major in place
        sh = lambda w, tr: (ops.bf16_shadow(w), ops.bf16_shadow(w, transposed=True) if tr else None)
        W = {"d": sh(P["attn.decoder_att.weight"], save), "ihe": sh(Wih0[:, :E], False), "ihc": sh(Wih0[:, E:], save)}
        for l in range(L):
            Wih, Whh, _, _ = layer_params(P, l)
            W[f"hh{l}"] = sh(Whh, save)
            if l > 0:
                W[f"ih{l}"] = sh(Wih, False)
        sv["W"] = W
        Xe_b, _ = ops.cast_bf16(X0[:, :E], True, False)
        Gx = [ops.gemm_bf16(Xe_b, W["ihe"][0], bias=bih0, tag="ih_fwd")]              # emb half hoisted
        ops.join(init_done)
        h0_b, _ = ops.cast_bf16(h0, True, False)
        ctx_b = torch.empty(N, E, dtype=BF16, device=dev)
    else:
        Gx = [ops.sgemm(X0[:, :E], Wih0[:, :E], transB=True, bias=bih0, tag="ih_fwd")]
        ops.join(init_done)
    for l in range(1, L):
        Gx.append(torch.empty(N, G, dtype=F32, device=dev))
    for t in range(T):
        bt, o0 = bs[t], off[t]
        o1 = o0 + bt
        # query = pre-step top-layer hidden (rnn_attn.py:69)
        if tc:
            q_b = h0_b[:bt] if t == 0 else outs[L - 1]["Hsb"][off[t - 1]:off[t - 1] + bt]
            ops.gemm_bf16(q_b, W["d"][0], bias=bd, out=att2_all[o0:o1], tag="att2_fwd")
        else:
            q = h0[:bt] if t == 0 else outs[L - 1]["Hs"][off[t - 1]:off[t - 1] + bt]
            _lin_small(q, P["attn.decoder_att.weight"], bd, out=att2_all[o0:o1])
        ops.attn_step_fwd(bt, Pn, att1, Fe, att2_all[o0:o1], wf, P["attn.full_att.bias"], P["embed.bias"],
                          alphas[:, t, :], Tcap * Pn, S, X0[o0:o1, E:], ctx_bf16=ctx_b[o0:o1] if tc else None)
        # + W_ih[:, E:] embed(ctx): on the tensor-core path it is accumulated inside the layer-0 step kernel
        # (K-concatenated operand [h | embed(ctx)]); otherwise a small product into the hoisted pre-activations
        fused0 = None
        if tc:
            fused0 = ops.rnn_step_x_tc_fwd(kind, Gx[0], ctx_b, W["hh0"][0], W["ihc"][0], layer_params(P, 0)[3], bs, t,
                                           h0=h0, h0_b=h0_b, c0=c0, save=save, out=outs[0], tag="step_fwd")
            if fused0 is None:
                ops.gemm_bf16(ctx_b[o0:o1], W["ihc"][0], out=Gx[0][o0:o1], beta=1.0, tag="ihc_fwd")
            else:
                outs[0] = fused0
        else:
            _lin_small(X0[o0:o1, E:], Wih0[:, E:], out=Gx[0][o0:o1], beta=1.0)
        for l in range(L):
            if l == 0 and fused0 is not None:
                continue
            Wih, Whh, bih, bhh = layer_params(P, l)
            if tc:
                if l > 0:
                    ops.gemm_bf16(outs[l - 1]["Hsb"][o0:o1], W[f"ih{l}"][0], bias=bih, out=Gx[l][o0:o1])
                outs[l] = ops.rnn_seq_tc_fwd(kind, Gx[l], W[f"hh{l}"][0], bhh, bs, h0=h0, h0_b=h0_b, c0=c0, save=save,
                                             t_range=(t, t + 1), out=outs[l], tag="step_fwd")
                if outs[l] is None:
                    raise RuntimeError("rnn_seq_tc_fwd refused a shape that rnn_seq_tc_fits accepted")
            else:
                if l > 0:
                    _lin_small(outs[l - 1]["Hs"][o0:o1], Wih, bih, out=Gx[l][o0:o1])
                outs[l] = ops.rnn_seq_fwd(kind, Gx[l], Whh, bhh, bs, h0=h0, c0=c0, save=save, t_range=(t, t + 1),
                                          out=outs[l], tag="step_fwd")
    sv.update(X0=X0, outs=outs, att2_all=att2_all, alphas=alphas, S=S, enc=enc)
    return outs[L - 1]["Hs"], alphas, sv


def _ctx_bf16(bs, Pn, F, alphas):
    ctx, _ = ops.attn_ctx_all(bs, Pn, F, alphas, out_dtype=BF16)             # row-major bf16: read in place by the GEMM
    return ctx, None


def attn_backward(mode, P, kind, L, caption, sv, dHs_top, dalphas=None, Gpen=None, emb_out=None, recurrent_done=None):
    """Gradients of every parameter given dHs_top (N,H) and the gradient w.r.t. alphas, either a
    full (B,Tcap,P) tensor `dalphas` (drop-in forward) or the per-(b,p) penalty term `Gpen`.
    Data parallelism: `emb_out()` supplies the embedding-gradient buffer, `recurrent_done(grads)` is
    called once the unit.* and embedding gradients are final (the attention parameters follow)."""
    bs, off, Pn, B, tc = sv["bs"], sv["off"], sv["Pn"], sv["B"], sv["tc"]
    T, N = len(bs), sum(bs)
    E = P["embeddings.weight"].shape[1]
    H = P["unit.weight_hh_l0"].shape[1]
    dev = dHs_top.device
    outs, X0, att1, Fe, att2_all, alphas, h0, c0 = (sv[k] for k in
                                                    ("outs", "X0", "att1", "Fe", "att2_all", "alphas", "h0", "c0"))
    Tcap = alphas.shape[1]
    Wih0 = P["unit.weight_ih_l0"]
    wf = P["attn.full_att.weight"].reshape(-1)
    Wd = P["attn.decoder_att.weight"]
    A = Wd.shape[0]
    W = sv.get("W")
    grads = {}
    de_all = torch.empty(N, Pn, dtype=F32, device=dev)
    datt2_all = torch.empty(N, A, dtype=F32, device=dev)
    gt_all = torch.empty(N, A, dtype=F32, device=dev)        # datt2 before the factor w_f (feeds dw_f, see attn_hoist_bwd)
    datt2_b = torch.empty(N, A, dtype=BF16, device=dev) if tc else None
    dctx_all = torch.empty(N, E, dtype=F32, device=dev)
    dHs = [torch.empty(N, H, dtype=F32, device=dev) for _ in range(L - 1)] + [dHs_top]
    bouts = [None] * L
    # embed.weight needs the feature-space contexts ctx = sum_p alpha_p F_p of every (t,b): known since the
    # forward pass, so the pass over F runs on the side stream beside the latency-bound reverse loop
    embed_q = mode == "bf16" and E % 8 == 0      # dW_embed = Q^T F as one GEMM (ops.attn_embed_q) instead of rebuilt contexts
    if mode == "bf16" and not embed_q:
        (ctx, ctxT), ctx_done = ops.fork(lambda: _ctx_bf16(bs, Pn, sv["F"], alphas), uses=(sv["F"], alphas))
    fused_bwd = None                # layer-0 fused step kernel: None = not tried yet, then True / False for the pass
    for t in reversed(range(T)):
        bt, o0 = bs[t], off[t]
        o1 = o0 + bt
        fused0 = False
        for l in reversed(range(L)):
            Wih, Whh, _, _ = layer_params(P, l)
            if tc and l == 0 and fused_bwd is not False:
                # layer 0: gate gradients and d embed(ctx_t) = dG W_ih[:, E:] from one stream of the gate-gradient tile
                # (single-layer decoders: also the query gradient of step t+1's attention, datt2_{t+1} . W_dec)
                qfold = L == 1 and ops.query_fold_ok(A)
                r = ops.rnn_step_x_tc_bwd(kind, W["hh0"][1], W["ihc"][1], bs, t, outs[0], dHs[0], dctx_all, h0=h0, c0=c0,
                                          out=bouts[0], tag="step_bwd", query=(W["d"][1], datt2_b) if qfold else None)
                fused_bwd = r is not None
                if fused_bwd:
                    bouts[0], fused0 = r, True
                    continue
            if tc:
                bouts[l] = ops.rnn_seq_tc_bwd(kind, W[f"hh{l}"][1], bs, outs[l], dHs[l], h0=h0, c0=c0,
                                              t_range=(t + 1, t), out=bouts[l], tag="step_bwd")
                if bouts[l] is None:
                    raise RuntimeError("rnn_seq_tc_bwd refused a shape that rnn_seq_tc_fits accepted")
                if l > 0:                                                             # dX of layer l at step t
                    ops.gemm_bf16(bouts[l]["dGb"][o0:o1], W[f"ih{l}"][0], b_t=True, out=dHs[l - 1][o0:o1])
            else:
                bouts[l] = ops.rnn_seq_bwd(kind, Whh, bs, outs[l], dHs[l], h0=h0, c0=c0, t_range=(t + 1, t),
                                           out=bouts[l], tag="step_bwd")
                if l > 0:
                    ops.sgemm(bouts[l]["dG"][o0:o1], Wih, out=dHs[l - 1][o0:o1])
        # d embed(ctx_t) = dG W_ih[:, E:]
        if fused0:
            pass
        elif tc:
            ops.gemm_bf16(bouts[0]["dGb"][o0:o1], W["ihc"][1], out=dctx_all[o0:o1], tag="ihc_dx")
        else:
            ops.sgemm(bouts[0]["dG"][o0:o1], Wih0[:, E:], out=dctx_all[o0:o1])
        if dalphas is not None:
            dal, dal_stride = dalphas[:, t, :], Tcap * Pn
        else:
            dal, dal_stride = Gpen, Pn
        ops.attn_step_bwd(bt, Pn, att1, Fe, att2_all[o0:o1], wf, alphas[:, t, :], Tcap * Pn, dal, dal_stride,
                          dctx_all[o0:o1], de_all[o0:o1], datt2_all[o0:o1],
                          datt2_bf16=datt2_b[o0:o1] if tc else None, gt=gt_all[o0:o1])
        # the query was the PRE-step top-layer hidden: its gradient joins the carried dh of the top layer
        if tc and fused0 and L == 1 and ops.query_fold_ok(A) and t > 0:
            pass                         # consumed by the next (t-1) fused step kernel
        elif tc:
            ops.gemm_bf16(datt2_b[o0:o1], W["d"][1], out=bouts[L - 1]["dstate"][0][:bt], beta=1.0, tag="att2_dx")
        else:
            dq = ops.sgemm(datt2_all[o0:o1], Wd)
            ops.add_rows(bouts[L - 1]["dstate"][0], dq, bt)

    # ---- everything below reads what the reverse loop left behind and is mutually independent: it runs as concurrent
    # chains on side streams (ops.fork lanes) -- each far too small to fill the GPU, and issued serially they were a
    # quarter of the step -- while the main stream carries the chain that ends in the embedding gradient.
    pending = []

    def side(fn, uses, lane):
        out, ev = ops.fork(fn, uses=uses, lane=lane)
        pending.append(ev)
        return out, ev

    # the encoder-projection chain: one pass over att1, then the largest GEMM of the step
    def enc_chain():
        datt1, _, dwf = ops.attn_hoist_bwd(bs, Pn, att1, att2_all, de_all, wf, want_t=False, gt_all=gt_all)
        if mode == "bf16":
            dWe = ops.gemm_bf16(datt1, sv["F"], a_t=True, b_t=True, tag="att1_dw")               # datt1^T F, both in place
        else:
            dWe = ops.sgemm(datt1, sv["F"], transA=True, tag="att1_dw")
        return dWe, ops.colsum(datt1), dwf

    (dWe, dbe, dwf), _ = side(enc_chain, (att1, att2_all, de_all, wf, sv["F"]), 0)
    if embed_q:
        # dW_embed = Q^T F: one pass over alphas / dctx, then a 2*E*B*P*C FLOP GEMM
        grads["embed.weight"], _ = side(
            lambda: ops.gemm_bf16(ops.attn_embed_q(bs, Pn, alphas, dctx_all), sv["F"], a_t=True, b_t=True, tag="embed_dw"),
            (alphas, dctx_all, sv["F"]), 1)
    # initial state: every layer started from the same h0 / c0 (rnn_attn.py:62)
    dh0 = bouts[0]["dstate"][0]
    dc0 = bouts[0]["dstate"][1]
    for l in range(1, L):
        dh0 = dh0 + bouts[l]["dstate"][0]
        dc0 = dc0 + bouts[l]["dstate"][1]

    # the bias gradients and the other small column sums: one chain of their own
    def small_sums():
        out = {}
        if tc:                      # bias gradients = column sums of the bf16 gate gradients
            for l in range(L):
                out[f"unit.bias_ih_l{l}"] = ops.colsum(bouts[l]["dGb"])
                out[f"unit.bias_hh_l{l}"] = ops.colsum(bouts[l]["dGhb"]) if kind == _lib.ST_GRU else out[f"unit.bias_ih_l{l}"]
        else:
            for l in range(L):
                out[f"unit.bias_hh_l{l}"] = ops.colsum(bouts[l]["dGh"])
                out[f"unit.bias_ih_l{l}"] = ops.colsum(bouts[l]["dG"])
        out["attn.decoder_att.bias"] = ops.colsum(datt2_all)
        # sum of every de (== 0 up to rounding: softmax is shift invariant): two-stage column sum
        nde = de_all.numel()
        wde = 256 if nde % 256 == 0 else (Pn if nde % Pn == 0 else 1)
        out["attn.full_att.bias"] = ops.colsum(ops.colsum(de_all.reshape(-1, wde)).reshape(-1, 1))
        out["embed.bias"] = ops.colsum(dctx_all)
        out["init_h.bias"] = ops.colsum(dh0)
        if kind == _lib.ST_LSTM:
            out["init_c.bias"] = ops.colsum(dc0)
        return out

    sums, sums_done = side(small_sums, (datt2_all, de_all, dctx_all, dh0), 4)

    if "mean_b" in sv:
        dinit = lambda d: ops.gemm_bf16(ops.cast_bf16(d, True, False)[0], sv["mean_b"], a_t=True, b_t=True, tag="init_dw")
    else:
        dinit = lambda d: ops.sgemm(d, sv["mean_f"], transA=True)
    (grads["init_h.weight"], dic), _ = side(lambda: (dinit(dh0), dinit(dc0) if kind == _lib.ST_LSTM else None),
                                            (dh0, sv["mean_f"]), 5)
    if kind == _lib.ST_LSTM:
        grads["init_c.weight"] = dic

    # recurrent weight gradients: W_hh (and the decoder_att weight, which reads the same shifted states) on one lane,
    # W_ih on another
    h0_b16 = ops.cast_bf16(h0, True, False)[0] if tc else None

    def hh_chain():
        out, top = {}, None
        for l in range(L):
            if tc:
                Hprev = ops.shift_states(outs[l]["Hsb"], bs, h0_b16)
                out[f"unit.weight_hh_l{l}"] = ops.gemm_bf16(bouts[l]["dGhb"], Hprev, a_t=True, b_t=True, tag="hh_dw")
            else:
                Hprev = ops.shift_states(outs[l]["Hs"], bs, h0)
                out[f"unit.weight_hh_l{l}"] = weight_grad(mode, bouts[l]["dGh"], Hprev, "hh_dw")
            if l == L - 1:
                top = Hprev
        out["attn.decoder_att.weight"] = weight_grad(mode, datt2_b if tc else datt2_all, top, "att2_dw")
        return out

    def ih_chain():
        out = {}
        for l in range(L):
            if tc:
                inp_b = ops.cast_bf16(X0, True, False)[0] if l == 0 else outs[l - 1]["Hsb"]
                out[f"unit.weight_ih_l{l}"] = ops.gemm_bf16(bouts[l]["dGb"], inp_b, a_t=True, b_t=True, tag="ih_dw")
            else:
                inp = X0 if l == 0 else outs[l - 1]["Hs"]
                out[f"unit.weight_ih_l{l}"] = weight_grad(mode, bouts[l]["dG"], inp, "ih_dw")
        return out

    whh, hh_done = side(hh_chain, (X0, datt2_all), 2)
    wih, ih_done = side(ih_chain, (X0,), 3)
    # main stream: the embedding gradient (the buffer is cleared on a lane of its own)
    dEmb = emb_out() if emb_out is not None else torch.empty_like(P["embeddings.weight"])
    _, zeroed = ops.fork(dEmb.zero_, lane=6)
    if tc:
        dXemb = ops.gemm_bf16(bouts[0]["dGb"], W["ihe"][0], b_t=True, tag="ih_dx")
    else:
        dXemb = ops.sgemm(bouts[0]["dG"], Wih0[:, :E], tag="ih_dx")
    ops.join(zeroed)
    ops.pack_inputs_bwd(dXemb, dEmb, None, caption, bs, False)
    grads["embeddings.weight"] = dEmb
    for ev in (hh_done, ih_done, sums_done):
        ops.join(ev)
    grads.update(whh)
    grads.update(wih)
    grads.update(sums)
    if recurrent_done is not None:
        recurrent_done(grads)
    # attention parameters
    grads["attn.encoder_att.weight"], grads["attn.encoder_att.bias"] = dWe, dbe
    grads["attn.full_att.weight"] = dwf.reshape(1, -1)
    # embed: ctx_e = W_embed ctx + b with ctx = sum_p alpha_p F_p rebuilt for all (t,b) in one pass
    if embed_q:
        pass                                     # forked above
    elif mode == "bf16":
        dc_b, _ = ops.cast_bf16(dctx_all, True, False)
        ops.join(ctx_done)
        grads["embed.weight"] = ops.gemm_bf16(dc_b, ctx, a_t=True, b_t=True, tag="embed_dw")
    else:
        ctx, _ = ops.attn_ctx_all(bs, Pn, sv["F"], alphas)
        grads["embed.weight"] = ops.sgemm(dctx_all, ctx, transA=True, tag="embed_dw")
    for ev in pending:
        ops.join(ev)
    return grads


def grid_layout(mod):
    lay = getattr(mod, "grid_layout", "BCP")
    if lay not in ("BCP", "BPC"):
        raise ValueError('grid_layout must be "BCP" (channels-first, the reference) or "BPC" (channels-last)')
    return lay


def _check_inputs(feature, caption, lengths, C, layout="BCP"):
    if not (feature.is_cuda and caption.is_cuda):
        raise RuntimeError("showtell_b200 runs on CUDA tensors only (no CPU fallback)")
    if feature.dim() != 3 or feature.shape[2 if layout == "BPC" else 1] != C:
        want = f"(B, P, {C}) channels-last" if layout == "BPC" else f"(B, {C}, P) channels-first"
        raise ValueError(f"cnn_feature must be {want}, got {tuple(feature.shape)}")
    if caption.dim() != 2 or caption.shape[0] != feature.shape[0] or caption.dtype != torch.int64:
        raise ValueError("image_caption must be (B, T) int64")
    if len(lengths) != feature.shape[0]:
        raise ValueError("caption_size must have one entry per batch row")
    bs = _lib.batch_sizes(lengths)
    if len(bs) > caption.shape[1]:
        raise ValueError("caption_size exceeds the padded caption length")
    return bs


class AttnLogitsFn(torch.autograd.Function):
    """RNN_Attn.forward as the reference exposes it: (logits (N,V) packed, alphas (B,T,P))."""

    @staticmethod
    def forward(ctx, mod, feature, caption, lengths, *params):
        names = [n for n, _ in mod.named_parameters()]
        P = {n: p.detach() for n, p in zip(names, params)}
        mode = mod.compute_dtype
        if ctx.needs_input_grad[1]:
            raise NotImplementedError("showtell_b200 attention decoders produce no gradient for cnn_feature (the reference "
                                      "detaches the grid, Attention/cnn_attn.py:47): pass cnn_feature.detach()")
        f = feature.detach().contiguous()
        f = f if f.dtype == BF16 else f.to(F32)       # bf16 grids (autocast trunk) are consumed as they are
        cap = caption.contiguous()
        lay = grid_layout(mod)
        bs = _check_inputs(f, cap, lengths, mod.nos_filters, lay)
        save = any(ctx.needs_input_grad)
        Hs, alphas, sv = attn_forward(mode, P, mod._kind, mod.num_layers, f, cap, bs, save, layout=lay)
        vocab = Linear(mode, P["linear.weight"], P["linear.bias"], save, "vocab")
        logits = vocab.fwd(Hs)                                                        # rnn_attn.py:71,115
        ctx.names, ctx.P, ctx.mod, ctx.vocab, ctx.sv, ctx.caption = names, P, mod, vocab, sv, cap
        return logits, alphas

    @staticmethod
    def backward(ctx, dlogits, dalphas):
        mod = ctx.mod
        dHs, dWv, dbv = ctx.vocab.bwd(dlogits.contiguous().to(F32))
        if dalphas is None:
            dalphas = torch.zeros_like(ctx.sv["alphas"])
        grads = attn_backward(mod.compute_dtype, ctx.P, mod._kind, mod.num_layers, ctx.caption, ctx.sv, dHs,
                              dalphas=dalphas.contiguous().to(F32))
        grads["linear.weight"], grads["linear.bias"] = dWv, dbv
        return (None, None, None, None) + tuple(grads[n] for n in ctx.names)


class AttnLossFn(torch.autograd.Function):
    """forward_loss: CE(mean over packed tokens, same-index targets) + alpha_c * mean_{B,P}((1 -
    sum_t alpha)^2)  (main_attn.py:126,130-131), forward and backward in one pass.  `denom_tokens` /
    `denom_batch` are the global token / batch counts under data parallelism."""

    @staticmethod
    def forward(ctx, mod, feature, caption, lengths, alpha_c, denom_tokens, denom_batch, *params):
        names = [n for n, _ in mod.named_parameters()]
        P = {n: p.detach() for n, p in zip(names, params)}
        mode = mod.compute_dtype
        if ctx.needs_input_grad[1]:
            raise NotImplementedError("showtell_b200 attention decoders produce no gradient for cnn_feature (the reference "
                                      "detaches the grid, Attention/cnn_attn.py:47): pass cnn_feature.detach()")
        f = feature.detach().contiguous()
        f = f if f.dtype == BF16 else f.to(F32)       # bf16 grids (autocast trunk) are consumed as they are
        cap = caption.contiguous()
        lay = grid_layout(mod)
        bs = _check_inputs(f, cap, lengths, mod.nos_filters, lay)
        need = any(ctx.needs_input_grad)
        kind, L = mod._kind, mod.num_layers
        dt = float(denom_tokens if denom_tokens is not None else sum(bs))
        db = float(denom_batch if denom_batch is not None else f.shape[0])
        coef = float(alpha_c) / (db * f.shape[1 if lay == "BPC" else 2])
        red = getattr(mod, "grad_reducer", None)       # data parallelism: parallel.GradReducer

        Pn = f.shape[1 if lay == "BPC" else 2]

        def prologue(inputs, out):
            # eager, in front of the step graph: the grid is read ONCE, re-laid / cast straight into the graph's
            # static operands (no device copy of the 200 MB grid into a static input first)
            F, mean_f = grid_operands(mode, inputs[0], lay, out=None if out is None else (out[0], out[1]))
            if out is None:
                return [F, mean_f, inputs[1]]
            out[2].copy_(inputs[1], non_blocking=True)
            return out

        def body(F, mean_f, capt):
            Hs, alphas, sv = attn_forward(mode, P, kind, L, None, capt, bs, need, layout=lay, grid=(F, mean_f, Pn))
            target = ops.pack_targets(capt, bs, P["linear.weight"].shape[0])
            gout = red.slots([P["linear.weight"].shape, P["linear.bias"].shape]) if (red is not None and need) else None
            # the stream-K dHs product reduces into a cleared buffer: cleared early, beside the forward loop
            dHs_out = ops.fork(lambda: torch.zeros_like(Hs), lane=7) if (need and mode == "bf16") else None
            loss, dHs, grads, vdone = vocab_ce(mode, P, Hs, target, dt, need, gout=gout, dHs_out=dHs_out)
            pen_sum, Gpen = ops.attn_penalty(sv["S"], coef)
            loss = loss + coef * pen_sum.reshape(())
            g2 = None
            if need and red is None:
                g2 = attn_backward(mode, P, kind, L, capt, sv, dHs, Gpen=Gpen)
                ops.join(vdone)
                g2.update(grads)
            elif need:
                # data parallel: the vocabulary projection's exchange overlaps the reverse loop, the
                # recurrent + embedding gradients' overlaps the hoisted attention passes, the rest is last
                lin = ["linear.weight", "linear.bias"]
                grads.update(zip(lin, red.reduce([grads[n] for n in lin], ready=vdone)))
                rec = sorted(n for n in names if n.startswith("unit.")) + ["embeddings.weight"]

                def emb_out():
                    v = red.slots([P[n].shape for n in rec])
                    return v[-1] if v else torch.empty_like(P["embeddings.weight"])

                def recurrent_done(g):
                    g.update(zip(rec, red.reduce([g[n] for n in rec])))

                g2 = attn_backward(mode, P, kind, L, capt, sv, dHs, Gpen=Gpen, emb_out=emb_out,
                                   recurrent_done=recurrent_done)
                rest = [n for n in names if n in g2 and n not in rec]
                g2.update(zip(rest, red.reduce([g2[n] for n in rest])))
                red.finish()
                g2.update(grads)
            return loss, alphas, g2

        key = ("attn", mode, kind, L, lay, tuple(bs), tuple(f.shape), str(f.dtype), tuple(cap.shape), need, dt, db, float(alpha_c),
               tuple(p.data_ptr() for p in params))
        loss, alphas, ctx.grads = graphs.run(mod, key, body, (f, cap), prologue=prologue)
        ctx.names, ctx.mod, ctx.ticket = names, mod, graphs.ticket(mod)
        loss, alphas = loss.clone(), alphas.clone()
        ctx.mark_non_differentiable(alphas)
        return loss, alphas

    @staticmethod
    def backward(ctx, g, _dalphas):
        if ctx.grads is None:
            raise RuntimeError("forward_loss was run without grad enabled")
        graphs.check_ticket(ctx.mod, ctx.ticket)
        return (None,) * 7 + tuple(ops.scale_multi([ctx.grads[n] for n in ctx.names], g))   # chain rule, one launch
