"""Host-side orchestration of the base (non-attention) decoder: which kernels run in which order
for RNN.forward / forward_loss / backward.  Mirrors the data flow of rnn.py:27-35 + main.py:145-151
with the time-invariant input projection hoisted out of the recurrence.

Two arithmetic modes (module ctor `dtype=`):
  fp32  CUDA-core GEMMs (st_sgemm), materialised logits; parity bar 1e-4 relative.
  bf16  tcgen05 tensor-core GEMMs on bf16 copies of fp32 master weights / activations with fp32
        accumulation (st_gemm_bf16), vocabulary projection fused with cross-entropy so the (N,V)
        logits never reach HBM (st_vocab_ce_fwd/bwd); parity bar 2e-2 relative.
The recurrent kernels keep fp32 state in both modes.
"""
import torch

from . import _lib, graphs, ops

F32 = torch.float32


def _bf16(X):
    """bf16 view of an activation: as is if it already is one, else one cast."""
    return X if X.dtype == torch.bfloat16 else ops.cast_bf16(X, True, False)[0]


class Linear:
    """y = x W^T + b and its backward, in the arithmetic of `mode`.  bf16 mode keeps ONE bf16 copy of every
    operand: the backward products dW = dY^T X and dX = dY W read X, W and dY in place as MN-major UMMA operands
    (ops.gemm_bf16 a_t / b_t) -- no transposed copies, and the weight's bf16 shadow is cached until the optimizer
    changes the parameter (ops.bf16_shadow)."""

    def __init__(self, mode, W, bias, need_bwd, tag):
        self.mode, self.W, self.bias, self.need_bwd, self.tag = mode, W, bias, need_bwd, tag
        if mode == "bf16":
            self.Wb = ops.bf16_shadow(W)

    def fwd(self, X, XT=None, out_dtype=F32):
        """X fp32 or (bf16 mode) already bf16.  (`XT` is accepted for older callers and ignored.)"""
        if self.mode == "fp32":
            self.X = X
            return ops.sgemm(X, self.W, transB=True, bias=self.bias, tag=self.tag + "_fwd")
        self.Xb = _bf16(X)
        return ops.gemm_bf16(self.Xb, self.Wb, bias=self.bias, out_dtype=out_dtype, tag=self.tag + "_fwd")

    def bwd_bf16(self, dYb, dYT=None, need_dx=True):
        """Backward from gradients that already are bf16 (row-major): (dX or None, dW)."""
        dW = ops.gemm_bf16(dYb, self.Xb, a_t=True, b_t=True, tag=self.tag + "_dw")           # dY^T X
        dX = ops.gemm_bf16(dYb, self.Wb, b_t=True, tag=self.tag + "_dx") if need_dx else None  # dY W
        return dX, dW

    def bwd(self, dY, need_dx=True):
        """Returns (dX or None, dW, db)."""
        db = ops.colsum(dY) if self.bias is not None else None
        if self.mode == "fp32":
            dW = ops.sgemm(dY, self.X, transA=True, tag=self.tag + "_dw")
            dX = ops.sgemm(dY, self.W, tag=self.tag + "_dx") if need_dx else None
            return dX, dW, db
        dX, dW = self.bwd_bf16(_bf16(dY), None, need_dx)
        return dX, dW, db


def weight_grad(mode, dY, X, tag):
    """dW = dY^T X for a product whose forward ran inside another kernel (W_hh)."""
    if mode == "fp32":
        return ops.sgemm(dY, X, transA=True, tag=tag)
    return ops.gemm_bf16(_bf16(dY), _bf16(X), a_t=True, b_t=True, tag=tag)


def layer_params(P, l):
    return (P[f"unit.weight_ih_l{l}"], P[f"unit.weight_hh_l{l}"],
            P[f"unit.bias_ih_l{l}"], P[f"unit.bias_hh_l{l}"])


def stack_forward(mode, P, kind, L, X, bs, save):
    """All L recurrent layers over the packed input X (N, in).  Layer l's input projection is one
    hoisted GEMM over every time step; the recurrence is one persistent kernel per layer.
    Returns (top-layer Hs (N,H), per-layer saved state)."""
    layers, inp = [], X
    for l in range(L):
        Wih, Whh, bih, bhh = layer_params(P, l)
        lin = Linear(mode, Wih, bih, save, "ih")
        Gx = lin.fwd(inp)                                                  # W_ih x + b_ih, all steps
        o, tc = None, False
        if mode == "bf16" and ops.rnn_seq_tc_supported(kind, Whh.shape[1]):
            o = ops.rnn_seq_tc_fwd(kind, Gx, ops.bf16_shadow(Whh), bhh, bs, save=save, tag="seq_fwd")
            tc = o is not None
        if o is None:                                                      # CUDA-core recurrent kernel
            o = ops.rnn_seq_fwd(kind, Gx, Whh, bhh, bs, save=save, tag="seq_fwd")
        layers.append({"lin": lin, "out": o, "tc": tc})
        inp = o["Hsb"] if tc else o["Hs"]                                  # the tensor-core kernel also emits h_t as bf16
    return layers[-1]["out"]["Hs"], layers


def bptt_prealloc(P, kind, L, bs, layers):
    """Buffers of the tensor-core BPTT kernels, created (and the two small ones zero-filled) on a side stream: called
    before the vocabulary products are issued, so nothing but the dHs product stands between them and the BPTT kernel
    (the fills used to sit there: ~10 us of launch latency on the critical path).  {layer: (buffers, event)}."""
    pre = {}
    for l in range(L):
        if layers[l]["tc"]:
            WhhT = ops.bf16_shadow(layer_params(P, l)[1], transposed=True)
            Wih = layer_params(P, l)[0]

            def make(W=WhhT, n_in=Wih.shape[1]):
                o = ops.rnn_seq_tc_bwd_buffers(kind, W, bs, W.device)
                o["dX"] = torch.zeros(sum(bs), n_in, dtype=F32, device=W.device)   # output of the stream-K dG W_ih product
                return o

            pre[l] = ops.fork(make, lane=6)
    return pre


def stack_backward(mode, P, kind, L, bs, layers, dHs_top, grads, need_dx0=True, dx0_ready=None, after_bptt=None,
                   prealloc=None):
    """BPTT through the L layers, top down.  Fills grads[...] for the unit.* parameters and returns
    the gradient w.r.t. the packed layer-0 input (N, in_0).  `dx0_ready(dX)`: called as soon as that
    input gradient exists (bf16 mode: before layer 0's weight-gradient products), so the embedding
    gradient and its exchange can start while the last weight gradients are still being formed.
    `after_bptt()`: called right after the top layer's BPTT kernel is issued (work to be ordered behind it)."""
    dH = dHs_top
    pending = []
    for l in reversed(range(L)):
        _, Whh, _, _ = layer_params(P, l)
        sv = layers[l]
        b = None
        if sv["tc"]:
            bufs = None
            if prealloc is not None and l in prealloc:
                bufs, ev = prealloc[l]
                ops.join(ev)
            b = ops.rnn_seq_tc_bwd(kind, ops.bf16_shadow(Whh, transposed=True), bs, sv["out"], dH, tag="seq_bwd", out=bufs)
            if b is not None and after_bptt is not None and l == L - 1:
                after_bptt()
                after_bptt = None
        if b is not None:                                                  # tensor-core BPTT: bf16 gate gradients out
            need_dx = l > 0 or need_dx0
            dGb, dGhb, lin, Hsb = b["dGb"], b["dGhb"], sv["lin"], sv["out"]["Hsb"]
            # Everything below reads the gate gradients and nothing else of this step: the weight-gradient products
            # and the bias sums run as concurrent chains on side streams (each far too small to fill the GPU) beside
            # the input-gradient chain  dG W_ih -> embedding / feature gradients  on the main stream.
            (grads[f"unit.weight_hh_l{l}"],), e1 = ops.fork(
                lambda: (ops.gemm_bf16(dGhb, ops.shift_states(Hsb, bs), a_t=True, b_t=True, tag="hh_dw"),),
                uses=(dGhb, Hsb), lane=1)
            (grads[f"unit.weight_ih_l{l}"],), e2 = ops.fork(lambda: (lin.bwd_bf16(dGb, None, need_dx=False)[1],),
                                                            uses=(dGb, lin.Xb), lane=2)
            (grads[f"unit.bias_ih_l{l}"], grads[f"unit.bias_hh_l{l}"]), e3 = ops.fork(
                lambda: (lambda s: (s, ops.colsum(dGhb) if dGhb is not dGb else s))(ops.colsum(dGb)),
                uses=(dGb, dGhb), lane=3)
            pending += [e1, e2, e3]
            if need_dx and bufs is not None and "dX" in bufs:                                # dG W_ih
                with ops.gemm_c_zeroed():
                    dH = ops.gemm_bf16(dGb, lin.Wb, b_t=True, tag="ih_dx", out=bufs["dX"])
            else:
                dH = ops.gemm_bf16(dGb, lin.Wb, b_t=True, tag="ih_dx") if need_dx else None
            if l == 0 and dx0_ready is not None:
                dx0_ready(dH)
            continue
        Hprev = ops.shift_states(sv["out"]["Hs"], bs)
        b = ops.rnn_seq_bwd(kind, Whh, bs, sv["out"], dH, tag="seq_bwd")
        if after_bptt is not None and l == L - 1:
            after_bptt()
            after_bptt = None
        grads[f"unit.weight_hh_l{l}"] = weight_grad(mode, b["dGh"], Hprev, "hh_dw")  # dGh^T Hprev
        grads[f"unit.bias_hh_l{l}"] = ops.colsum(b["dGh"])
        dH, dW, db = sv["lin"].bwd(b["dG"], need_dx=(l > 0 or need_dx0))
        grads[f"unit.weight_ih_l{l}"], grads[f"unit.bias_ih_l{l}"] = dW, db
        if l == 0 and dx0_ready is not None:
            dx0_ready(dH)
    for e in pending:
        ops.join(e)
    return dH


def base_forward(mode, P, kind, L, feature, caption, bs, save):
    X = ops.pack_inputs(P["embeddings.weight"], feature, caption, bs, True, bf16=(mode == "bf16"))   # rnn.py:29-31
    return stack_forward(mode, P, kind, L, X, bs, save)


def base_backward_from_dHs(mode, P, kind, L, caption, bs, layers, dHs, grads, want_dfeature, feature_shape,
                           emb_out=None, emb_done=None, after_bptt=None, prealloc=None):
    """`emb_out()` / `emb_done()`: data parallelism -- buffer for the embedding gradient, and a call
    the moment it is final (the layer-0 weight gradients are formed after it)."""
    box = {}
    # the embedding gradient is a scatter-add into zeros: clear the (V, E) buffer beside the BPTT kernel, not after it
    dEmb = emb_out() if emb_out is not None else torch.empty_like(P["embeddings.weight"])
    _, zeroed = ops.fork(dEmb.zero_, lane=4)

    def dx0_ready(dX):
        ops.join(zeroed)
        box["dfeat"] = torch.empty(feature_shape, dtype=F32, device=dX.device) if want_dfeature else None
        ops.pack_inputs_bwd(dX, dEmb, box["dfeat"], caption, bs, True)
        grads["embeddings.weight"] = dEmb
        if emb_done is not None:
            emb_done()

    stack_backward(mode, P, kind, L, bs, layers, dHs, grads, dx0_ready=dx0_ready, after_bptt=after_bptt,
                   prealloc=prealloc)
    return box["dfeat"]


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
        mode = mod.compute_dtype
        feature_c = feature.detach().contiguous().to(F32)
        caption_c = caption.contiguous()
        bs = _check_inputs(feature_c, caption_c, lengths, mod.embed_dim)
        save = any(ctx.needs_input_grad)
        Hs, layers = base_forward(mode, P, mod._kind, mod.num_layers, feature_c, caption_c, bs, save)
        vocab = Linear(mode, P["linear.weight"], P["linear.bias"], save, "vocab")
        logits = vocab.fwd(layers[-1]["out"]["Hsb"] if layers[-1]["tc"] else Hs)      # rnn.py:33
        ctx.names, ctx.P, ctx.mod, ctx.vocab = names, P, mod, vocab
        ctx.bs, ctx.layers, ctx.caption = bs, layers, caption_c
        ctx.feature_shape = feature_c.shape
        return logits

    @staticmethod
    def backward(ctx, dlogits):
        mod, grads = ctx.mod, {}
        dlogits = dlogits.contiguous().to(F32)
        dHs, grads["linear.weight"], grads["linear.bias"] = ctx.vocab.bwd(dlogits)
        dfeat = base_backward_from_dHs(mod.compute_dtype, ctx.P, mod._kind, mod.num_layers, ctx.caption, ctx.bs,
                                       ctx.layers, dHs, grads, ctx.needs_input_grad[1], ctx.feature_shape)
        return (None, dfeat, None, None) + tuple(grads[n] for n in ctx.names)


def vocab_ce(mode, P, Hs, target, denom, need, gout=None, Hs_bf16=None, side_sms=None, defer_db=False, dHs_out=None,
             fork_loss=True):
    """Mean cross-entropy of the vocabulary projection of Hs (N,H) and, if `need`, its gradients.
    Returns (loss (0-d), dHs or None, grads dict, event).  bf16 mode: the weight / bias gradients do
    not feed the rest of the backward pass, so they run on the side stream (ops.fork) beside the BPTT
    kernels; `event` marks their completion (None when they ran in line) -- ops.join it, or hand it to
    the gradient reducer, before the gradients are read.  `gout`: optional [dW, db] output tensors (a
    gradient reducer's symmetric bucket) for bf16 mode.  `Hs_bf16`: the bf16 copy of Hs the recurrent kernel
    already wrote (else Hs is cast once).  `side_sms`: SMs the side-stream products may occupy (None: all) -- the
    cooperative BPTT kernel that follows on the main stream needs the rest free to start.  `defer_db` (bf16 mode):
    the bias gradient -- column sums over the 100 MB dlogits matrix, a long grid that would sit in front of the BPTT
    kernel's CTAs -- is not issued here: grads["_late_db"] = a function that issues it (the caller orders it behind the
    BPTT kernel, where HBM is idle) and returns (db, event).  `dHs_out` (bf16 mode): (zero-filled (N, H) fp32 buffer,
    event) -- cleared early on a side stream, so the stream-K dHs product does not clear it on the critical path."""
    Wv, bv = P["linear.weight"], P["linear.bias"]
    grads = {}
    if mode == "fp32":
        logits = ops.sgemm(Hs, Wv, transB=True, bias=bv, tag="vocab_fwd")             # rnn.py:33
        loss_sum, _, dl = ops.ce_fwd_bwd(logits, target, grad_scale=(1.0 / denom) if need else None,
                                         inplace=True)                                # main.py:149
        dHs = None
        if need:
            grads["linear.weight"] = ops.sgemm(dl, Hs, transA=True, tag="vocab_dw")   # dlogits^T Hs
            grads["linear.bias"] = ops.colsum(dl)
            dHs = ops.sgemm(dl, Wv, tag="vocab_dx")                                   # dlogits W_v
        return (loss_sum / denom).reshape(()), dHs, grads, None
    Wb = ops.bf16_shadow(Wv)
    Hb = Hs_bf16 if Hs_bf16 is not None else _bf16(Hs)
    loss_sum, lse = ops.vocab_ce_fwd(Hb, Wb, bv, target, tag="vocab_fwd")
    # (the mean: a one-element kernel, kept off the main stream -- there it sat between the dHs product and BPTT)
    loss, loss_done = (ops.fork(lambda: (loss_sum / denom).reshape(()), uses=(loss_sum,), lane=7) if fork_loss
                       else (None, None))
    dHs, done = None, None
    if need:
        # dlogits = (softmax - onehot) / denom recomputed tile by tile and written ONCE, row-major bf16; the three
        # products below read it in place: dW = dlogits^T Hs (both operands MN-major), db = column sums, dHs = dlogits W_v
        Pm, _ = ops.vocab_ce_bwd(Hb, Wb, bv, target, lse, 1.0 / denom, want_t=False, tag="vocab_dlogits")
        # dW / db do not feed the rest of the backward pass: side stream, beside dHs = dlogits W_v and whatever follows.
        # (A cooperative launch -- the BPTT kernel -- starts only on a drained GPU, so nothing launched before it can
        # run beside it; `side_sms` caps the side products' persistent grid for callers that want to leave SMs free.)
        def weight_grads():
            with ops.gemm_sm_limit(side_sms):
                dW = ops.gemm_bf16(Pm, Hb, a_t=True, b_t=True, tag="vocab_dw", out=gout[0] if gout else None)
            return dW, (None if defer_db else ops.colsum(Pm, out=gout[1] if gout and len(gout) > 1 else None))

        (grads["linear.weight"], db), done = ops.fork(weight_grads, uses=(Pm, Hb))
        if defer_db:
            grads["_late_db"] = lambda: ops.fork(lambda: ops.colsum(Pm), uses=(Pm,), lane=5)
        else:
            grads["linear.bias"] = db
        if dHs_out is not None:
            ops.join(dHs_out[1])
            with ops.gemm_c_zeroed():
                dHs = ops.gemm_bf16(Pm, Wb, b_t=True, tag="vocab_dx", out=dHs_out[0])
        else:
            dHs = ops.gemm_bf16(Pm, Wb, b_t=True, tag="vocab_dx")
    ops.join(loss_done)
    if loss is None:
        loss = (loss_sum / denom).reshape(())
    return loss, dHs, grads, done


class DirectCtx:
    """Stand-in for autograd's ctx when a loss function's forward is run outside autograd (forward_backward): the
    fused step already produces every gradient, so nothing needs to be recorded."""

    def __init__(self, want_dfeature, n_fixed, n_params):
        self.needs_input_grad = (False, bool(want_dfeature)) + (False,) * (n_fixed - 2) + (True,) * n_params

    def mark_non_differentiable(self, *a):
        pass


def assign_grads(mod, ctx, feature):
    """.grad of every parameter := the step's gradient tensors (what loss.backward() would have accumulated into
    freshly zeroed / None grads, main.py:146,151); the gradient w.r.t. cnn_feature continues into its producer."""
    for n, p in mod.named_parameters():
        p.grad = ctx.grads[n]
    dfeat = getattr(ctx, "dfeat", None)
    if dfeat is not None and feature.requires_grad:
        if feature.is_leaf:
            feature.grad = dfeat if feature.grad is None else feature.grad + dfeat
        else:
            feature.backward(dfeat)


class BaseLossFn(torch.autograd.Function):
    """forward_loss: mean cross-entropy over the packed tokens (main.py:145,149), with forward and
    backward run back to back (fp32 mode turns the logits buffer into its own gradient in place;
    bf16 mode never materialises logits).  `denom` = number of tokens the mean runs over (the
    global token count under data parallelism)."""

    @staticmethod
    def forward(ctx, mod, feature, caption, lengths, denom, *params):
        names = [n for n, _ in mod.named_parameters()]
        P = {n: p.detach() for n, p in zip(names, params)}
        mode = mod.compute_dtype
        feature_c = feature.detach().contiguous().to(F32)
        caption_c = caption.contiguous()
        bs = _check_inputs(feature_c, caption_c, lengths, mod.embed_dim)
        if len(bs) > caption_c.shape[1]:
            raise ValueError("caption_size exceeds the padded caption length")
        need = any(ctx.needs_input_grad)
        want_dfeat = bool(ctx.needs_input_grad[1])
        denom = float(denom if denom is not None else sum(bs))
        kind, L = mod._kind, mod.num_layers
        red = getattr(mod, "grad_reducer", None)       # data parallelism: parallel.GradReducer

        def body(feat, cap):
            # the packed targets depend on the captions only: formed beside the forward pass, not between the forward
            # recurrence and the vocabulary product
            # (issued after pack_inputs: that call is where a bad token id of an EARLIER step is reported)
            # Single GPU only: under data parallelism the extra side-stream work (this, the early BPTT buffers, the loss
            # mean) changed which kernels reach the SMs first around the BPTT kernel and the vocabulary bucket's exchange
            # no longer overlapped it (2 GPUs: 0.634 -> 0.684 ms) -- that path keeps the serial order it was measured with.
            lean = red is None
            X = ops.pack_inputs(P["embeddings.weight"], feat, cap, bs, True, bf16=(mode == "bf16"))   # rnn.py:29-31
            if lean:
                target, tdone = ops.fork(lambda: ops.pack_targets(cap, bs, P["linear.weight"].shape[0]), uses=(cap,), lane=7)
            Hs, layers = stack_forward(mode, P, kind, L, X, bs, need)
            if lean:
                ops.join(tdone)
            else:
                target = ops.pack_targets(cap, bs, P["linear.weight"].shape[0])
            # bf16 mode: the vocabulary bias gradient is issued behind the BPTT kernel (see vocab_ce) and, under data
            # parallelism, travels with the last bucket instead of the vocabulary weight's
            defer = False     # measured: the sums then lengthen the backward tail by what they save in front of BPTT
            lin = ["linear.weight"] if defer else ["linear.weight", "linear.bias"]
            gout = red.slots([P[n].shape for n in lin]) if (red is not None and need) else None
            pre = bptt_prealloc(P, kind, L, bs, layers) if (need and lean) else None
            # (the dHs product keeps its own clear: with it the side stream's dW product, issued first, also STARTS first
            # and the two persistent grids share the GPU -- dW and db are then final before BPTT starts, which is what the
            # data-parallel exchange of that bucket overlaps; cleared early, dHs took every SM first and the bucket was
            # ready only after BPTT: 0.653 -> 0.745 ms at 8 GPUs)
            dHs_out = None
            loss, dHs, grads, vdone = vocab_ce(mode, P, Hs, target, denom, need, gout=gout, defer_db=defer, dHs_out=dHs_out,
                                               fork_loss=lean,
                                               Hs_bf16=layers[-1]["out"]["Hsb"] if layers[-1]["tc"] else None)
            late = grads.pop("_late_db", None)
            box = {}

            def after_bptt():
                if late is not None:
                    grads["linear.bias"], box["db_done"] = late()

            dfeat = None
            if need and red is None:
                dfeat = base_backward_from_dHs(mode, P, kind, L, cap, bs, layers, dHs, grads, want_dfeat, feat.shape,
                                               after_bptt=after_bptt, prealloc=pre)
                ops.join(vdone)
                ops.join(box.get("db_done"))
            elif need:
                # data parallel: two exchanges on the reducer's side stream -- the vocabulary projection's overlaps BPTT ...
                grads.update(zip(lin, red.reduce([grads[n] for n in lin], ready=vdone)))
                # ... and ONE exchange for everything the end of the backward pass produces (recurrent weights, biases,
                # the embedding): their gradients become final within a few microseconds of each other, and every
                # exchange pays a launch and two cross-GPU barriers
                last = [n for n in names if n not in lin and n != "embeddings.weight"] + ["embeddings.weight"]

                def emb_out():
                    v = red.slots([P[n].shape for n in last])
                    return v[-1] if v else torch.empty_like(P["embeddings.weight"])

                dfeat = base_backward_from_dHs(mode, P, kind, L, cap, bs, layers, dHs, grads, want_dfeat, feat.shape,
                                               emb_out=emb_out, after_bptt=after_bptt, prealloc=pre)
                ops.join(box.get("db_done"))
                grads.update(zip(last, red.reduce([grads[n] for n in last])))
                red.finish()
            return loss, (grads if need else None), dfeat

        key = ("base", mode, kind, L, tuple(bs), tuple(feature_c.shape), tuple(caption_c.shape), need, want_dfeat,
               denom, tuple(p.data_ptr() for p in params))
        loss, ctx.grads, ctx.dfeat = graphs.run(mod, key, body, (feature_c, caption_c))
        ctx.names, ctx.mod, ctx.ticket = names, mod, graphs.ticket(mod)
        return loss.clone()

    @staticmethod
    def backward(ctx, g):
        if ctx.grads is None:
            raise RuntimeError("forward_loss was run without grad enabled")
        graphs.check_ticket(ctx.mod, ctx.ticket)
        src = [ctx.grads[n] for n in ctx.names]
        want_dfeat = ctx.dfeat is not None and ctx.needs_input_grad[1]
        out = ops.scale_multi(src + ([ctx.dfeat] if want_dfeat else []), g)       # chain rule, one launch
        return (None, out[-1] if want_dfeat else None, None, None, None) + tuple(out[:len(src)])
