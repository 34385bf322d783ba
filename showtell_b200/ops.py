"""Tensor-level wrappers over the C ABI: each takes CUDA tensors, passes raw pointers and sizes.

These are host-side plumbing only (allocation with torch.empty, stream = torch's current stream);
all arithmetic happens inside libshowtell_b200.so.
"""
import torch

from . import _lib
from ._lib import check, int_array, ptr, stream_ptr

F32, I64, I32 = torch.float32, torch.int64, torch.int32


class KernelTimer:
    """Optional CUDA-event timing of tagged kernel launches (used by bench.py for the roofline of
    the dominant kernel; events are recorded on the launching stream, inside the timed step)."""

    def __init__(self):
        self.spans = {}

    def begin(self, tag):
        e0 = torch.cuda.Event(enable_timing=True)
        e0.record()
        return (tag, e0)

    def end(self, tok):
        e1 = torch.cuda.Event(enable_timing=True)
        e1.record()
        self.spans.setdefault(tok[0], []).append((tok[1], e1))

    def summary(self):
        """tag -> (launches, mean ms).  Call after a synchronize."""
        return {k: (len(v), sum(a.elapsed_time(b) for a, b in v) / len(v)) for k, v in self.spans.items()}


TIMER = None      # set to a KernelTimer to time tagged launches

OVERLAP = True    # run independent kernel groups of a step on a second stream (see fork)
_SIDE = {}


def fork(fn, uses=(), lane=0):
    """Runs fn() on this device's side stream number `lane` (`uses`: the tensors it reads, kept alive for it), ordered after everything issued so far on the current
    stream, and returns (fn's result, event to join on).  Used for work that the rest of the step
    does not depend on (the vocabulary dW GEMM runs beside the BPTT kernel, which occupies at most
    128 of the 148 SMs and is latency-bound).  Capturable: fork/join become graph dependencies.
    Serial (no second stream) while TIMER records per-kernel times or OVERLAP is off."""
    if not OVERLAP or TIMER is not None:
        return fn(), None
    dev = torch.cuda.current_device()
    side = _SIDE.get((dev, lane))
    if side is None:
        side = _SIDE[(dev, lane)] = torch.cuda.Stream(device=dev)
    main = torch.cuda.current_stream()
    ready = torch.cuda.Event()
    ready.record(main)
    for t in uses:
        t.record_stream(side)
    with torch.cuda.stream(side):
        side.wait_event(ready)
        out = fn()
        done = torch.cuda.Event()
        done.record(side)
    for t in (out if isinstance(out, (tuple, list)) else (out,)):
        if torch.is_tensor(t):
            t.record_stream(main)
    return out, done


class gemm_sm_limit:
    """Context: the tensor-core GEMMs launched inside occupy at most `n` SMs (None / 0: all)."""

    def __init__(self, n):
        self.n = int(n or 0)

    def __enter__(self):
        if self.n:
            _lib.load().st_gemm_set_sm_limit(self.n)

    def __exit__(self, *a):
        if self.n:
            _lib.load().st_gemm_set_sm_limit(0)


class gemm_c_zeroed:
    """Context: the fp32 `out=` buffers of the gemm_bf16 calls inside were cleared by the caller (torch.zeros on a side
    stream, early): a stream-K launch then skips the memset it would put in front of its kernel."""

    def __enter__(self):
        _lib.load().st_gemm_set_c_zeroed(1)

    def __exit__(self, *a):
        _lib.load().st_gemm_set_c_zeroed(0)


def bptt_side_sms(kind, H, B):
    """SMs left over beside the tensor-core BPTT kernel for a batch of B rows (at least a quarter of the GPU)."""
    lib = _lib.load()
    sms = torch.cuda.get_device_properties(torch.cuda.current_device()).multi_processor_count
    return max(sms - lib.st_rnn_seq_tc_bwd_ctas(int(kind), int(H), int(B)), sms // 4)


def join(done):
    """The current stream waits for a fork()ed group."""
    if done is not None:
        torch.cuda.current_stream().wait_event(done)


def sgemm(A, B, *, transA=False, transB=False, bias=None, out=None, alpha=1.0, beta=0.0, tag=None):
    """out[M,N] = alpha * op(A) op(B) + beta * out + bias.  2-D fp32 tensors; the last dim must be
    contiguous, rows may be strided (sub-matrix views of a wider matrix are fine)."""
    lib = _lib.load()
    for t in (A, B):
        if t.dim() != 2 or t.stride(1) != 1 or t.dtype != F32 or not t.is_cuda:
            raise ValueError("sgemm operands must be 2-D fp32 CUDA tensors with unit inner stride")
    M, K = (A.shape[1], A.shape[0]) if transA else (A.shape[0], A.shape[1])
    Kb, N = (B.shape[1], B.shape[0]) if transB else (B.shape[0], B.shape[1])
    if K != Kb:
        raise ValueError(f"sgemm: inner dimensions differ ({K} vs {Kb})")
    if out is None:
        out = torch.empty(M, N, dtype=F32, device=A.device)
    elif out.shape != (M, N) or out.stride(1) != 1 or out.dtype != F32:
        raise ValueError("sgemm: bad `out`")
    if bias is not None and (bias.numel() != N or bias.dtype != F32):
        raise ValueError("sgemm: bad bias")
    import ctypes as C
    tok = TIMER.begin(tag) if (TIMER is not None and tag) else None
    check(lib.st_sgemm(int(transA), int(transB), M, N, K, float(alpha), C.c_void_p(A.data_ptr()),
                       max(A.stride(0), 1), C.c_void_p(B.data_ptr()), max(B.stride(0), 1), float(beta),
                       C.c_void_p(out.data_ptr()), max(out.stride(0), 1), ptr(bias), stream_ptr()),
          "st_sgemm")
    if tok:
        TIMER.end(tok)
    return out


def pack_inputs(emb, feature, caption, bs, with_feature, width=None, bf16=False):
    """Packed inputs (N, width >= E); only the first E columns are written (attention models use
    width = 2E and let the attention kernel fill the other half).  bf16: rows written as bf16 (the operand of the
    hoisted tensor-core input projection), row stride padded to a multiple of 8."""
    lib = _lib.load()
    _lib.raise_token_error()
    N, E = sum(bs), emb.shape[1]
    if bf16:
        X = torch.empty(N, ((width or E) + 7) // 8 * 8, dtype=torch.bfloat16, device=emb.device)[:, :width or E]
        check(lib.st_pack_inputs_bf16(_raw(X), X.stride(0), ptr(emb, F32), E, emb.shape[0],
                                      ptr(feature, F32) if with_feature else None, ptr(caption, I64), caption.shape[1],
                                      int(with_feature), len(bs), int_array(bs), stream_ptr()), "st_pack_inputs_bf16")
        return X
    X = torch.empty(N, width or E, dtype=F32, device=emb.device)
    check(lib.st_pack_inputs(ptr(X, F32), X.shape[1], ptr(emb, F32), E, emb.shape[0],
                             ptr(feature, F32) if with_feature else None,
                             ptr(caption, I64), caption.shape[1], int(with_feature), len(bs),
                             int_array(bs), stream_ptr()), "st_pack_inputs")
    return X


def pack_inputs_bwd(dX, dEmb, dfeature, caption, bs, with_feature):
    lib = _lib.load()
    check(lib.st_pack_inputs_bwd(ptr(dX, F32), dX.stride(0), ptr(dEmb, F32), dEmb.shape[1], dEmb.shape[0],
                                 ptr(dfeature, F32), ptr(caption, I64), caption.shape[1],
                                 int(with_feature), len(bs), int_array(bs), stream_ptr()),
          "st_pack_inputs_bwd")


def pack_targets(caption, bs, V):
    """Packed targets; ids outside [0, V) are clamped and reported (st_token_error)."""
    lib = _lib.load()
    out = torch.empty(sum(bs), dtype=I64, device=caption.device)
    check(lib.st_pack_targets(ptr(out, I64), ptr(caption, I64), caption.shape[1], int(V), len(bs), int_array(bs),
                              stream_ptr()), "st_pack_targets")
    return out


def colsum(M, out=None, accumulate=False):
    lib = _lib.load()
    if M.dim() != 2 or M.stride(1) != 1:
        raise ValueError("colsum: 2-D tensor with unit inner stride expected")
    if out is None:
        out = torch.empty(M.shape[1], dtype=F32, device=M.device)
    import ctypes as C
    check(lib.st_colsum(ptr(out, F32), C.c_void_p(M.data_ptr()), int(M.dtype == torch.bfloat16), M.shape[0],
                        M.shape[1], max(M.stride(0), 1), int(accumulate), stream_ptr()), "st_colsum")
    return out


def scale_multi(tensors, g, outs=None):
    """[t * g for t in tensors] (fp32, contiguous; g a 0-d device tensor) with one launch per 32 tensors.
    The results are views of one flat buffer, or the given contiguous `outs`."""
    import ctypes as C
    lib = _lib.load()
    tensors = [t.contiguous() for t in tensors]
    if not tensors:
        return []
    g = g.to(F32).reshape(1)
    if outs is None:
        pad = lambda n: (n + 3) // 4 * 4
        flat = torch.empty(sum(pad(t.numel()) for t in tensors), dtype=F32, device=tensors[0].device)
        outs, off = [], 0
        for t in tensors:
            outs.append(flat[off:off + t.numel()].view_as(t))
            off += pad(t.numel())
    for i in range(0, len(tensors), 32):
        src, dst = tensors[i:i + 32], outs[i:i + 32]
        n = len(src)
        sp = (C.c_void_p * n)(*[t.data_ptr() for t in src])
        dp = (C.c_void_p * n)(*[t.data_ptr() for t in dst])
        cnt = (C.c_int64 * n)(*[t.numel() for t in src])
        check(lib.st_scale_multi(n, sp, dp, cnt, ptr(g, F32), stream_ptr()), "st_scale_multi")
    return outs


def rowsum_bf16(M, out=None):
    lib = _lib.load()
    if out is None:
        out = torch.empty(M.shape[0], dtype=F32, device=M.device)
    import ctypes as C
    check(lib.st_rowsum_bf16(ptr(out), C.c_void_p(M.data_ptr()), M.shape[0], M.shape[1], M.stride(0), stream_ptr()),
          "st_rowsum_bf16")
    return out


def _barrier(device):
    return torch.zeros(64, dtype=I32, device=device)


def rnn_seq_fwd(kind, Gx, Whh, bhh, bs, *, h0=None, c0=None, save=True, t_range=None, out=None, tag=None):
    """Returns dict(Hs, Cs, gates, ghn).  `out` lets a caller that runs single steps (attention
    models) keep writing into the same packed buffers."""
    lib = _lib.load()
    N, H = sum(bs), Whh.shape[1]
    dev = Gx.device
    o = out or {}
    if "Hs" not in o:
        o["Hs"] = torch.empty(N, H, dtype=F32, device=dev)
        o["Cs"] = torch.empty(N, H, dtype=F32, device=dev) if kind == _lib.ST_LSTM else None
        o["gates"] = torch.empty(N, Whh.shape[0], dtype=F32, device=dev) if save else None
        o["ghn"] = torch.empty(N, H, dtype=F32, device=dev) if (save and kind == _lib.ST_GRU) else None
        o["barrier"] = _barrier(dev)
    t0, t1 = t_range if t_range is not None else (0, len(bs))
    tok = TIMER.begin(tag) if (TIMER is not None and tag) else None
    check(lib.st_rnn_seq_fwd(kind, H, len(bs), int_array(bs), t0, t1, ptr(Gx, F32), ptr(Whh, F32),
                             ptr(bhh, F32), ptr(h0, F32), ptr(c0, F32), ptr(o["Hs"], F32), ptr(o["Cs"]),
                             ptr(o["gates"]), ptr(o["ghn"]), ptr(o["barrier"], I32), stream_ptr()),
          "st_rnn_seq_fwd")
    if tok:
        TIMER.end(tok)
    return o


def rnn_seq_bwd(kind, Whh, bs, saved, dHs, *, h0=None, c0=None, t_range=None, out=None, tag=None):
    """Returns dict(dG, dGh, dstate).  dstate (2, B0, H): [0] = dh0, [1] = dc0 once t reaches 0."""
    lib = _lib.load()
    N, H, GH = sum(bs), Whh.shape[1], Whh.shape[0]
    dev = dHs.device
    o = out or {}
    if "dG" not in o:
        o["dG"] = torch.empty(N, GH, dtype=F32, device=dev)
        o["dGh"] = torch.empty(N, GH, dtype=F32, device=dev) if kind == _lib.ST_GRU else o["dG"]
        o["dstate"] = torch.zeros(2, bs[0], H, dtype=F32, device=dev)
        o["barrier"] = _barrier(dev)
    t_hi, t_lo = t_range if t_range is not None else (len(bs), 0)
    tok = TIMER.begin(tag) if (TIMER is not None and tag) else None
    check(lib.st_rnn_seq_bwd(kind, H, len(bs), int_array(bs), t_hi, t_lo, ptr(Whh, F32), ptr(h0, F32),
                             ptr(c0, F32), ptr(saved["Hs"], F32), ptr(saved["Cs"]), ptr(saved["gates"], F32),
                             ptr(saved["ghn"]), ptr(dHs, F32), ptr(o["dG"], F32), ptr(o["dGh"], F32),
                             ptr(o["dstate"], F32), ptr(o["barrier"], I32), stream_ptr()),
          "st_rnn_seq_bwd")
    if tok:
        TIMER.end(tok)
    return o


def shift_states(Hs, bs, h0=None):
    """Hprev[n=(t,b)] = Hs[(t-1,b)] (h0[b] or 0 at t = 0); fp32 or bf16 rows."""
    lib = _lib.load()
    out = torch.empty_like(Hs)
    if Hs.dtype == torch.bfloat16:
        check(lib.st_shift_states_bf16(ptr(out, Hs.dtype), ptr(Hs, Hs.dtype), ptr(h0, Hs.dtype) if h0 is not None else None,
                                       Hs.shape[1], len(bs), int_array(bs), stream_ptr()), "st_shift_states_bf16")
        return out
    check(lib.st_shift_states(ptr(out, F32), ptr(Hs, F32), ptr(h0, F32), Hs.shape[1], len(bs),
                              int_array(bs), stream_ptr()), "st_shift_states")
    return out


def ce_fwd_bwd(logits, target, grad_scale=None, inplace=False):
    """Returns (loss_sum (1,), lse (N,), dlogits or None).  dlogits = (softmax - onehot) * grad_scale."""
    lib = _lib.load()
    N, V = logits.shape
    dev = logits.device
    loss_sum = torch.empty(1, dtype=F32, device=dev)
    lse = torch.empty(N, dtype=F32, device=dev)
    dl = None
    if grad_scale is not None:
        dl = logits if inplace else torch.empty_like(logits)
    check(lib.st_ce_fwd_bwd(ptr(logits, F32), logits.stride(0), ptr(target, I64), N, V, ptr(loss_sum),
                            ptr(lse), ptr(dl), float(grad_scale or 0.0), stream_ptr()), "st_ce_fwd_bwd")
    return loss_sum, lse, dl


def argmax_rows(X, out=None):
    """Row-wise arg-max (first maximal index).  `out` may be a strided 1-D int64 view (a column of the
    token matrix)."""
    lib = _lib.load()
    idx = torch.empty(X.shape[0], dtype=I64, device=X.device) if out is None else out
    import ctypes as C
    check(lib.st_argmax_rows(ptr(X, F32), X.stride(0), X.shape[0], X.shape[1], C.c_void_p(idx.data_ptr()),
                             idx.stride(0), stream_ptr()), "st_argmax_rows")
    return idx


def gather_rows(dst, table, idx):
    """dst[i, :W] = table[idx[i]] for a (possibly strided) 1-D int64 idx and a row-strided dst view."""
    lib = _lib.load()
    import ctypes as C
    check(lib.st_gather_rows(C.c_void_p(dst.data_ptr()), dst.stride(0), ptr(table, F32), table.shape[1],
                             C.c_void_p(idx.data_ptr()), idx.stride(0), idx.shape[0], stream_ptr()),
          "st_gather_rows")


def topk_rows(X, K):
    lib = _lib.load()
    val = torch.empty(X.shape[0], K, dtype=F32, device=X.device)
    idx = torch.empty(X.shape[0], K, dtype=I32, device=X.device)
    check(lib.st_topk_rows(ptr(X, F32), X.stride(0), X.shape[0], X.shape[1], K, ptr(val), ptr(idx), K,
                           stream_ptr()), "st_topk_rows")
    return val, idx


# ----------------------------------------------------------------------------- bf16 / tensor cores
BF16 = torch.bfloat16


def _raw(t):
    import ctypes as C
    return C.c_void_p(t.data_ptr())


def _rawn(t):
    return _raw(t) if t is not None else None


def gemm_bf16(A, B, *, bias=None, out_dtype=F32, alpha=1.0, beta=0.0, out=None, tag=None, a_t=False, b_t=False):
    """out[M,N] = alpha * A[M,K] . B[N,K]^T + bias on tcgen05 tensor cores.  A, B bf16 with unit inner
    stride and row strides that are multiples of 8.  a_t / b_t: that operand is passed as its transpose -- A as
    (K, M), B as (K, N) -- and consumed in place (MN-major UMMA operand): dW = dY^T X is
    gemm_bf16(dY, X, a_t=True, b_t=True), dX = dY W is gemm_bf16(dY, W, b_t=True)."""
    lib = _lib.load()
    for t in (A, B):
        if t.dim() != 2 or t.dtype != BF16 or t.stride(1) != 1 or not t.is_cuda:
            raise ValueError("gemm_bf16 operands must be 2-D bf16 CUDA tensors with unit inner stride")
    M, K = (A.shape[1], A.shape[0]) if a_t else A.shape
    N, Kb = (B.shape[1], B.shape[0]) if b_t else B.shape
    if K != Kb:
        raise ValueError(f"gemm_bf16: inner dimensions differ ({K} vs {Kb})")
    if out is None:
        out = torch.empty(M, N, dtype=out_dtype, device=A.device)
    elif tuple(out.shape) != (M, N) or out.stride(1) != 1 or out.dtype not in (F32, BF16) or not out.is_cuda:
        raise ValueError(f"gemm_bf16: `out` must be a ({M}, {N}) fp32 / bf16 CUDA tensor with unit inner stride")
    if bias is not None and (bias.numel() != N or bias.dtype != F32):
        raise ValueError("gemm_bf16: bad bias")
    tok = TIMER.begin(tag) if (TIMER is not None and tag) else None
    check(lib.st_gemm_bf16_ex(M, N, K, _raw(A), A.stride(0), int(a_t), _raw(B), B.stride(0), int(b_t), _raw(out),
                              out.stride(0), int(out.dtype == BF16), ptr(bias, F32), float(alpha), float(beta),
                              stream_ptr()), "st_gemm_bf16")
    if tok:
        TIMER.end(tok)
    return out


def split_tf32(x):
    """fp32 (R,C) -> (hi, lo): hi keeps the upper 11 mantissa bits (a tf32), lo = x - hi exactly.  Leading
    dimensions padded to multiples of 4 (TMA operands)."""
    lib = _lib.load()
    if x.dim() != 2 or x.dtype != F32 or x.stride(1) != 1 or not x.is_cuda:
        raise ValueError("split_tf32: 2-D fp32 CUDA tensor with unit inner stride expected")
    R, Cc = x.shape
    ld = (Cc + 3) // 4 * 4
    hi = torch.zeros(R, ld, dtype=F32, device=x.device)[:, :Cc]
    lo = torch.zeros(R, ld, dtype=F32, device=x.device)[:, :Cc]
    check(lib.st_split_tf32(ptr(x, F32), R, Cc, x.stride(0), _raw(hi), _raw(lo), ld, stream_ptr()), "st_split_tf32")
    return hi, lo


def gemm_tf32x3(A, B, *, bias=None, alpha=1.0, beta=0.0, out=None):
    """out[M,N] = alpha * A . B^T + bias to fp32 accuracy on the tensor cores.  A, B: (hi, lo) pairs from
    split_tf32."""
    lib = _lib.load()
    (Ah, Al), (Bh, Bl) = A, B
    M, K = Ah.shape
    N, Kb = Bh.shape
    if K != Kb:
        raise ValueError(f"gemm_tf32x3: inner dimensions differ ({K} vs {Kb})")
    if out is None:
        out = torch.empty(M, N, dtype=F32, device=Ah.device)
    elif tuple(out.shape) != (M, N) or out.stride(1) != 1 or out.dtype != F32 or not out.is_cuda:
        raise ValueError(f"gemm_tf32x3: `out` must be a ({M}, {N}) fp32 CUDA tensor with unit inner stride")
    check(lib.st_gemm_tf32x3(M, N, K, _raw(Ah), _raw(Al), Ah.stride(0), _raw(Bh), _raw(Bl), Bh.stride(0), _raw(out),
                             out.stride(0), ptr(bias, F32), float(alpha), float(beta), stream_ptr()), "st_gemm_tf32x3")
    return out


def gemm_tf32x3_topk(A, B, K_top, *, bias=None, want_tokens=False, want_stats=False):
    """Per-row top-K of A . B^T + bias (fp32-accurate 3xTF32 product) without materialising the product.
    A, B: (hi, lo) pairs from split_tf32.  Returns (val (M,K) f32, idx (M,K) i32[, tok (M,) i64][, row_max, row_sum]);
    row_sum = sum_j exp(x_j - row_max), the soft-max normaliser of the row."""
    lib = _lib.load()
    (Ah, Al), (Bh, Bl) = A, B
    M, K = Ah.shape
    N = Bh.shape[0]
    dev = Ah.device
    parts = lib.st_topk_parts(N)
    cv = torch.empty(M, parts, dtype=F32, device=dev)
    ci = torch.empty(M, parts, dtype=I32, device=dev)
    val = torch.empty(M, K_top, dtype=F32, device=dev)
    idx = torch.empty(M, K_top, dtype=I32, device=dev)
    tok = torch.empty(M, dtype=I64, device=dev) if want_tokens else None
    ps = torch.empty(M, parts // 4, dtype=F32, device=dev) if want_stats else None
    rmax = torch.empty(M, dtype=F32, device=dev) if want_stats else None
    rsum = torch.empty(M, dtype=F32, device=dev) if want_stats else None
    check(lib.st_gemm_tf32x3_topk(M, N, K, _raw(Ah), _raw(Al), Ah.stride(0), _raw(Bh), _raw(Bl), Bh.stride(0), ptr(bias, F32),
                                  int(K_top), ptr(cv), ptr(ci), ptr(val), ptr(idx), K_top, ptr(tok), 1, ptr(ps), ptr(rmax),
                                  ptr(rsum), stream_ptr()),
          "st_gemm_tf32x3_topk")
    out = (val, idx) + ((tok,) if want_tokens else ()) + ((rmax, rsum) if want_stats else ())
    return out


def vocab_topk_screen(h, Wv, bv, K_top, want_tokens=False):
    """Exact per-row top-K of h . Wv^T + bv through the bf16 screening GEMM + fp32 re-scoring (st_vocab_topk_screen).
    h (M, H), Wv (V, H), bv (V) fp32.  Returns (val (M,K) f32, idx (M,K) i32[, tok (M,) i64])."""
    lib = _lib.load()
    M, H = h.shape
    V = Wv.shape[0]
    dev = h.device
    hb, Wb = h.to(torch.bfloat16).contiguous(), Wv.to(torch.bfloat16).contiguous()
    wmax = torch.zeros(1, dtype=F32, device=dev)
    check(lib.st_row_norm_max(ptr(Wv, F32), V, H, ptr(wmax), stream_ptr()), "st_row_norm_max")
    parts = lib.st_topk_parts(V)
    cv = torch.empty(M, parts, dtype=F32, device=dev)
    ci = torch.empty(M, parts, dtype=I32, device=dev)
    val = torch.empty(M, K_top, dtype=F32, device=dev)
    idx = torch.empty(M, K_top, dtype=I32, device=dev)
    tok = torch.empty(M, dtype=I64, device=dev) if want_tokens else None
    check(lib.st_vocab_topk_screen(M, V, H, ptr(h, F32), ptr(hb), ptr(Wv, F32), ptr(Wb), ptr(bv, F32), ptr(wmax), int(K_top),
                                   ptr(cv), ptr(ci), ptr(val), ptr(idx), K_top, ptr(tok), 1, stream_ptr()),
          "st_vocab_topk_screen")
    return (val, idx, tok) if want_tokens else (val, idx)


def cast_bf16(src, want=True, want_t=False):
    """fp32 (R,C) -> (bf16 (R,C) or None, bf16 transpose (C,R) or None).  Leading dimensions are
    padded to multiples of 8 so the results are valid TMA operands; the returned tensors are the
    un-padded views."""
    lib = _lib.load()
    R, Cc = src.shape
    pad = lambda n: (n + 7) // 8 * 8
    d = torch.empty(R, pad(Cc), dtype=BF16, device=src.device)[:, :Cc] if want else None
    dT = torch.empty(Cc, pad(R), dtype=BF16, device=src.device)[:, :R] if want_t else None
    check(lib.st_cast_bf16(_raw(src), R, Cc, src.stride(0), _raw(d) if want else None,
                           d.stride(0) if want else 0, _raw(dT) if want_t else None,
                           dT.stride(0) if want_t else 0, stream_ptr()), "st_cast_bf16")
    return d, dT


# ---- bf16 shadows of the fp32 master parameters.
# The tensor-core kernels read bf16 weights; the parameters stay fp32 nn.Parameters (optimizer / checkpoint
# compatible).  Each parameter gets ONE persistent bf16 copy (stable pointer, so CUDA-graph replays keep reading it)
# that is re-cast only when the parameter changed: torch bumps `Tensor._version` on every in-place update (optimizer
# steps, load_state_dict), and showtell_b200.optim writes the shadow inside its own update kernel.  A transposed
# copy exists only where a kernel needs one as a K-major TMA operand (W_hh^T in the BPTT kernels); the GEMMs take
# operands in either major (gemm_bf16 a_t / b_t).  Writes through `.data` bypass the version counter: call
# invalidate_shadows() after such a write.
class _Shadow:
    __slots__ = ("buf", "version", "param")


_SHADOWS = {}


def _shadow_key(W, transposed):
    return (W.data_ptr(), tuple(W.shape), tuple(W.stride()), bool(transposed))


def _shadow_cast(W, sh, transposed):
    lib = _lib.load()
    R, Cc = W.shape
    check(lib.st_cast_bf16(_raw(W), R, Cc, W.stride(0), None if transposed else _raw(sh.buf),
                           0 if transposed else sh.buf.stride(0), _raw(sh.buf) if transposed else None,
                           sh.buf.stride(0) if transposed else 0, stream_ptr()), "st_cast_bf16")
    sh.version = W._version


def bf16_shadow(W, transposed=False):
    """bf16 copy (or bf16 transpose) of a 2-D fp32 parameter / parameter view, cached until the parameter changes."""
    if W.dim() != 2 or W.dtype != F32 or not W.is_cuda or W.stride(1) != 1:
        raise ValueError("bf16_shadow: 2-D fp32 CUDA tensor with unit inner stride expected")
    key = _shadow_key(W, transposed)
    sh = _SHADOWS.get(key)
    if sh is None:
        if len(_SHADOWS) > 512:
            _SHADOWS.clear()
        R, Cc = (W.shape[1], W.shape[0]) if transposed else W.shape
        sh = _SHADOWS[key] = _Shadow()
        sh.buf = torch.empty(R, (Cc + 7) // 8 * 8, dtype=BF16, device=W.device)[:, :Cc]
        sh.version, sh.param = None, W
    if sh.version != W._version:
        _shadow_cast(W, sh, transposed)
    return sh.buf


def refresh_shadows():
    """Re-cast every stale shadow (called before a captured step is replayed: the graph reads the shadows, the
    casts are deliberately not part of it)."""
    for key, sh in _SHADOWS.items():
        if sh.version != sh.param._version:
            _shadow_cast(sh.param, sh, key[3])


def invalidate_shadows():
    for sh in _SHADOWS.values():
        sh.version = None


def shadows_of(p):
    """(contiguous same-index bf16 shadow or None, [other shadows]) of a parameter: the fused optimizer writes the
    first inside its update kernel and marks the others stale."""
    direct, others = None, []
    for key, sh in _SHADOWS.items():
        if key[0] == p.data_ptr() and sh.param.shape == p.shape and sh.param.stride() == p.stride() and not key[3] \
                and sh.buf.is_contiguous():
            direct = sh
        elif sh.param.untyped_storage().data_ptr() == p.untyped_storage().data_ptr():
            others.append(sh)
    return direct, others


def vocab_ce_fwd(Hs, Wv, bv, target, tag=None):
    """Fused vocabulary projection + cross-entropy forward.  Returns (loss_sum (1,), lse (M,))."""
    lib = _lib.load()
    M, H = Hs.shape
    V = Wv.shape[0]
    dev = Hs.device
    parts = lib.st_vocab_ce_parts(V)
    pm = torch.empty(M, parts, dtype=F32, device=dev)
    ps = torch.empty(M, parts, dtype=F32, device=dev)
    tl = torch.empty(M, dtype=F32, device=dev)
    lse = torch.empty(M, dtype=F32, device=dev)
    loss = torch.empty(1, dtype=F32, device=dev)
    tok = TIMER.begin(tag) if (TIMER is not None and tag) else None
    check(lib.st_vocab_ce_fwd(M, V, H, _raw(Hs), Hs.stride(0), _raw(Wv), Wv.stride(0), ptr(bv, F32),
                              ptr(target, I64), ptr(pm), ptr(ps), ptr(tl), ptr(lse), ptr(loss), stream_ptr()),
          "st_vocab_ce_fwd")
    if tok:
        TIMER.end(tok)
    return loss, lse


def vocab_ce_bwd(Hs, Wv, bv, target, lse, scale, want_t=True, tag=None):
    """dlogits as bf16: P (M,V) and (optionally) its transpose PT (V,M)."""
    lib = _lib.load()
    M, H = Hs.shape
    V = Wv.shape[0]
    dev = Hs.device
    pad = lambda n: (n + 7) // 8 * 8
    P = torch.empty(M, pad(V), dtype=BF16, device=dev)[:, :V]
    PT = torch.empty(V, pad(M), dtype=BF16, device=dev)[:, :M] if want_t else None
    tok = TIMER.begin(tag) if (TIMER is not None and tag) else None
    check(lib.st_vocab_ce_bwd(M, V, H, _raw(Hs), Hs.stride(0), _raw(Wv), Wv.stride(0), ptr(bv, F32),
                              ptr(target, I64), ptr(lse, F32), float(scale), _raw(P), P.stride(0),
                              _raw(PT) if want_t else None, PT.stride(0) if want_t else 0, stream_ptr()),
          "st_vocab_ce_bwd")
    if tok:
        TIMER.end(tok)
    return P, PT


USE_CLUSTER = True      # cluster-resident recurrent kernels where the shape / GPU allow them


def rnn_seq_tc_supported(kind, H):
    return bool(_lib.load().st_rnn_seq_tc_supported(kind, H))


def rnn_seq_tc_fits(kind, H, B):
    """True if the persistent tensor-core recurrent grid for batch B is co-resident on this GPU
    ((H/16) unit tiles x ceil(B/128) batch tiles, one CTA per SM)."""
    if not rnn_seq_tc_supported(kind, H):
        return False
    import ctypes as C
    sms = C.c_int(0)
    check(_lib.load().st_device_info(C.byref(sms), None, None, None), "st_device_info")
    return (H // 16) * ((B + 127) // 128) <= sms.value


def rnn_seq_tc_fwd(kind, Gx, Whh_b, bhh, bs, *, h0=None, h0_b=None, c0=None, save=True, t_range=None, out=None,
                   tag=None):
    """Tensor-core persistent recurrence over steps t_range (default: all).  Returns dict(Hs, Hsb, Cs,
    gates, ghn) or None when the library reports the shape / grid as unsupported (caller falls back to
    rnn_seq_fwd).  `out` = the dict of a previous partial call (same packed buffers)."""
    lib = _lib.load()
    N, H = sum(bs), Whh_b.shape[1]
    dev = Gx.device
    o = out
    if o is None:
        o = {"Hs": torch.empty(N, H, dtype=F32, device=dev), "Hsb": torch.empty(N, H, dtype=BF16, device=dev),
             "Cs": torch.empty(N, H, dtype=F32, device=dev) if kind == _lib.ST_LSTM else None,
             "gates": torch.empty(N, Whh_b.shape[0], dtype=F32, device=dev) if save else None,
             "ghn": torch.empty(N, H, dtype=F32, device=dev) if (save and kind == _lib.ST_GRU) else None,
             "barrier": _barrier(dev)}
    t0, t1 = t_range if t_range is not None else (0, len(bs))
    tok = TIMER.begin(tag) if (TIMER is not None and tag) else None
    if USE_CLUSTER and t1 - t0 > 1 and lib.st_rnn_cluster_supported(kind, H):
        # whole-sequence runs: the cluster-resident kernel (h exchanged through distributed shared memory)
        st = lib.st_rnn_cluster_fwd(kind, H, len(bs), int_array(bs), t0, t1, ptr(Gx, F32), ptr(Whh_b, BF16),
                                    ptr(bhh, F32), ptr(h0, F32), ptr(h0_b, BF16), ptr(c0, F32), ptr(o["Hs"]),
                                    ptr(o["Hsb"]), ptr(o["Cs"]), ptr(o["gates"]), ptr(o["ghn"]), stream_ptr())
        if st == 0:
            if tok:
                TIMER.end(tok)
            return o
        if st != -3:
            check(st, "st_rnn_cluster_fwd")
    st = lib.st_rnn_seq_tc_fwd(kind, H, len(bs), int_array(bs), t0, t1, ptr(Gx, F32), ptr(Whh_b, BF16),
                               ptr(bhh, F32), ptr(h0, F32), ptr(h0_b, BF16), ptr(c0, F32), ptr(o["Hs"]),
                               ptr(o["Hsb"]), ptr(o["Cs"]), ptr(o["gates"]), ptr(o["ghn"]), ptr(o["barrier"]),
                               stream_ptr())
    if st == -3:
        return None
    check(st, "st_rnn_seq_tc_fwd")
    if tok:
        TIMER.end(tok)
    return o


STEP_X = True     # attention loop: fold the context half of W_ih into the recurrent step kernel (rnn_step_x_tc.cu)


def rnn_step_x_tc_fwd(kind, Gx, X_b, Whh_b, Wx_b, bhh, bs, t, *, h0, h0_b, c0=None, save=True, out=None, tag=None):
    """Step t of the attention decoders' recurrence with the projection of X_b (N, EX) bf16 -- embed(ctx) rows --
    accumulated in the same kernel as W_hh h_{t-1}.  Same result dict as rnn_seq_tc_fwd; None when unsupported."""
    lib = _lib.load()
    N, H, EX = sum(bs), Whh_b.shape[1], Wx_b.shape[1]
    if not STEP_X or not lib.st_rnn_step_x_tc_supported(kind, H, EX):
        return None
    dev = Gx.device
    o = out
    if o is None:
        o = {"Hs": torch.empty(N, H, dtype=F32, device=dev), "Hsb": torch.empty(N, H, dtype=BF16, device=dev),
             "Cs": torch.empty(N, H, dtype=F32, device=dev) if kind == _lib.ST_LSTM else None,
             "gates": torch.empty(N, Whh_b.shape[0], dtype=F32, device=dev) if save else None,
             "ghn": torch.empty(N, H, dtype=F32, device=dev) if (save and kind == _lib.ST_GRU) else None,
             "barrier": _barrier(dev)}
    tok = TIMER.begin(tag) if (TIMER is not None and tag) else None
    st = lib.st_rnn_step_x_tc_fwd(kind, H, EX, len(bs), int_array(bs), t, ptr(Gx, F32), _raw(X_b), X_b.stride(0),
                                  _raw(Whh_b), _raw(Wx_b), Wx_b.stride(0), ptr(bhh, F32), ptr(h0, F32), _raw(h0_b),
                                  ptr(c0, F32), ptr(o["Hs"]), ptr(o["Hsb"]), ptr(o["Cs"]), ptr(o["gates"]), ptr(o["ghn"]),
                                  stream_ptr())
    if st == -3:
        return None
    check(st, "st_rnn_step_x_tc_fwd")
    if tok:
        TIMER.end(tok)
    return o


STEP_X_QUERY = True   # ... and the attention-query gradient datt2_{t+1} . W_dec into dh_t


def query_fold_ok(A):
    """Shapes for which rnn_step_x_tc_bwd can take the attention-query gradient (`query=`)."""
    return STEP_X_QUERY and A % 64 == 0 and 64 <= A <= 512


def rnn_step_x_tc_bwd(kind, WhhT_b, WxT_b, bs, t, saved, dHs, dX, *, h0=None, c0=None, out=None, tag=None, query=None):
    """Step t of the reverse pass through the fused step: the gate gradients (same dict as rnn_seq_tc_bwd) and
    rows of step t of dX (N, EX) fp32 = dG_t W_x.  None when the shape is unsupported (needs EX == H, H % 64 == 0).
    `query` = (W_dec^T (H, A) bf16, datt2 (N, A) bf16): also adds datt2[rows of t+1] . W_dec to the carried dh_t."""
    lib = _lib.load()
    N, H, GH = sum(bs), WhhT_b.shape[0], WhhT_b.shape[1]
    if not STEP_X or H % 64 != 0 or H > 512 or WxT_b.shape[0] != H or dX.shape[1] != H:
        return None
    dev = dHs.device
    ldt = (N + 7) // 8 * 8
    o = out
    if o is None:
        dGb = torch.empty(N, GH, dtype=BF16, device=dev)         # row-major only: the GEMMs read them in place (a_t)
        dGhb = torch.empty(N, GH, dtype=BF16, device=dev) if kind == _lib.ST_GRU else dGb
        o = {"dGb": dGb, "dGT_full": None, "dGT": None, "dGhb": dGhb, "dGhT_full": None, "dGhT": None,
             "dbih": None, "dbhh": None, "dstate": torch.zeros(2, bs[0], H, dtype=F32, device=dev), "barrier": _barrier(dev)}
    tok = TIMER.begin(tag) if (TIMER is not None and tag) else None
    st = lib.st_rnn_step_x_tc_bwd(kind, H, len(bs), int_array(bs), t, _raw(WhhT_b), _raw(WxT_b), WxT_b.stride(0),
                                  ptr(h0, F32), ptr(c0, F32), ptr(saved["Hs"], F32), ptr(saved["Cs"]),
                                  ptr(saved["gates"], F32), ptr(saved["ghn"]), ptr(dHs, F32), _raw(o["dGb"]),
                                  _rawn(o["dGT_full"]), _raw(o["dGhb"]), _rawn(o["dGhT_full"]), ldt, ptr(o["dstate"]),
                                  ptr(dX, F32), dX.stride(0), ptr(o["barrier"]),
                                  _raw(query[0]) if query else None, query[0].stride(0) if query else 0,
                                  _raw(query[1]) if query else None, query[1].stride(0) if query else 0,
                                  query[0].shape[1] if query else 0, stream_ptr())
    if st == -3:
        return None
    check(st, "st_rnn_step_x_tc_bwd")
    if tok:
        TIMER.end(tok)
    return o


def rnn_seq_tc_bwd_buffers(kind, WhhT_b, bs, dev, want_bias=True, transposed=False):
    """The output / scratch buffers of rnn_seq_tc_bwd (its `out=`).  Two of them are zero-filled (the incoming state
    gradient and the kernel's hand-off counters): a caller that creates them early on a side stream (ops.fork) keeps
    those two small fills out of the dependency chain  dHs product -> BPTT kernel."""
    N, H, GH = sum(bs), WhhT_b.shape[0], WhhT_b.shape[1]
    ldt = (N + 7) // 8 * 8
    mk = lambda: (torch.empty(N, GH, dtype=BF16, device=dev),
                  torch.empty(GH, ldt, dtype=BF16, device=dev) if transposed else None)
    dGb, dGT = mk()
    dGhb, dGhT = mk() if kind == _lib.ST_GRU else (dGb, dGT)
    want_bias = want_bias and transposed
    return {"dGb": dGb, "dGT_full": dGT, "dGT": dGT[:, :N] if transposed else None, "dGhb": dGhb, "dGhT_full": dGhT,
            "dGhT": dGhT[:, :N] if transposed else None,
            "dbih": torch.empty(GH, dtype=F32, device=dev) if want_bias else None,
            "dbhh": torch.empty(GH, dtype=F32, device=dev) if want_bias else None,
            "dstate": torch.zeros(2, bs[0], H, dtype=F32, device=dev), "barrier": _barrier(dev)}


def rnn_seq_tc_bwd(kind, WhhT_b, bs, saved, dHs, *, h0=None, c0=None, t_range=None, out=None, want_bias=True,
                   tag=None, transposed=False):
    """Returns dict(dGb, dGhb, dstate[, dGT, dGhT, dbih, dbhh]) (bf16 GEMM operands) or None if unsupported.
    transposed=False (default): only the row-major gate gradients are written -- the weight-gradient GEMMs read
    them in place (gemm_bf16 a_t) and the bias gradients are their column sums (colsum)."""
    lib = _lib.load()
    N, H, GH = sum(bs), WhhT_b.shape[0], WhhT_b.shape[1]
    dev = dHs.device
    ldt = (N + 7) // 8 * 8
    o = out
    if o is None:
        o = rnn_seq_tc_bwd_buffers(kind, WhhT_b, bs, dev, want_bias=want_bias, transposed=transposed)
    t_hi, t_lo = t_range if t_range is not None else (len(bs), 0)
    tok = TIMER.begin(tag) if (TIMER is not None and tag) else None
    st = lib.st_rnn_seq_tc_bwd(kind, H, len(bs), int_array(bs), t_hi, t_lo, ptr(WhhT_b, BF16), ptr(h0, F32),
                               ptr(c0, F32), ptr(saved["Hs"], F32), ptr(saved["Cs"]), ptr(saved["gates"], F32),
                               ptr(saved["ghn"]), ptr(dHs, F32), _raw(o["dGb"]), _rawn(o["dGT_full"]),
                               _raw(o["dGhb"]), _rawn(o["dGhT_full"]), ldt, ptr(o["dbih"]), ptr(o["dbhh"]),
                               ptr(o["dstate"]), ptr(o["barrier"]), stream_ptr())
    if st == -3:
        return None
    check(st, "st_rnn_seq_tc_bwd")
    if tok:
        TIMER.end(tok)
    return o


# ----------------------------------------------------------------------------- attention
ACT_LEAKY, ACT_TANH = 0, 1


def attn_relayout(f, bf16=False, want_t=True, out=None):
    """f (B,C,P) fp32 or bf16 channels-first -> F (B*P, C), FT (C, B*P) or None, mean_f (B,C).
    out = (F, mean_f): write into these (a step graph's static operands) instead of fresh tensors."""
    lib = _lib.load()
    B, Cc, Pn = f.shape
    dt = BF16 if bf16 else F32
    if out is not None:
        F, mean_f = out
        if want_t or F.shape != (B * Pn, Cc) or F.dtype != dt or not F.is_contiguous() or mean_f.shape != (B, Cc):
            raise ValueError("attn_relayout: out= does not match the grid")
    else:
        F = torch.empty(B * Pn, Cc, dtype=dt, device=f.device)
        mean_f = torch.empty(B, Cc, dtype=F32, device=f.device)
    ld = (B * Pn + 7) // 8 * 8
    FT = torch.empty(Cc, ld, dtype=dt, device=f.device)[:, :B * Pn] if want_t else None
    fn = lib.st_attn_relayout_bf16in if f.dtype == BF16 else lib.st_attn_relayout
    check(fn(ptr(f, f.dtype), B, Cc, Pn, _raw(F), _raw(FT) if want_t else None, ld, int(bf16),
             ptr(mean_f), stream_ptr()), "st_attn_relayout")
    return F, FT, mean_f


def attn_grid_bpc(f, bf16=False, out=None):
    """Channels-last grid f (B, P, C) fp32 or bf16 -> F (B*P, C) in the compute storage type (the grid itself when the
    types agree: no copy), mean_f (B, C) fp32.  out = (F, mean_f): write into these instead (a step graph's static
    operands: the cast, or one copy when the types agree, lands there directly)."""
    lib = _lib.load()
    B, Pn, Cc = f.shape
    F = f.reshape(B * Pn, Cc)
    dt = BF16 if bf16 else F32
    if out is not None:
        oF, mean_f = out
        if oF.shape != F.shape or oF.dtype != dt or not oF.is_contiguous() or mean_f.shape != (B, Cc):
            raise ValueError("attn_grid_bpc: out= does not match the grid")
        if bf16 and F.dtype != BF16:
            check(lib.st_cast_bf16(_raw(F), B * Pn, Cc, F.stride(0), _raw(oF), oF.stride(0), None, 0, stream_ptr()),
                  "st_cast_bf16")
        else:
            oF.copy_(F)
        F = oF
    else:
        if bf16 and F.dtype != BF16:
            F = cast_bf16(F, True, False)[0]
        elif not bf16 and F.dtype != F32:
            F = F.to(F32)
        mean_f = torch.empty(B, Cc, dtype=F32, device=f.device)
    check(lib.st_grid_mean_bpc(_raw(f), int(f.dtype == BF16), B, Pn, Cc, ptr(mean_f), stream_ptr()), "st_grid_mean_bpc")
    return F, mean_f


def attn_step_fwd(rows, Pn, att1, Fe, att2, wf, bf, b_embed, alphas_t, alpha_stride, S, ctx_out, act=ACT_LEAKY,
                  ctx_bf16=None, tag="attn_fwd"):
    """att1 (B*P, A), Fe (B*P, E) (fp32 or bf16), att2 (rows, A).  alphas_t / ctx_out are (possibly
    strided) views whose first element is row 0; ctx_out row stride = ctx_out.stride(0)."""
    lib = _lib.load()
    A, E = att1.shape[1], Fe.shape[1]
    tok = TIMER.begin(tag) if (TIMER is not None and tag) else None
    check(lib.st_attn_step_fwd(rows, Pn, A, E, _raw(att1), _raw(Fe), int(att1.dtype == BF16), _raw(att2),
                               ptr(wf, F32), ptr(bf, F32), ptr(b_embed, F32), _raw(alphas_t), alpha_stride,
                               ptr(S, F32), _raw(ctx_out), ctx_out.stride(0),
                               _raw(ctx_bf16) if ctx_bf16 is not None else None,
                               ctx_bf16.stride(0) if ctx_bf16 is not None else 0, act, stream_ptr()),
          "st_attn_step_fwd")
    if tok:
        TIMER.end(tok)


def attn_step_bwd(rows, Pn, att1, Fe, att2, wf, alphas_t, alpha_stride, dalpha, dalpha_stride, dctx, de_out,
                  datt2, act=ACT_LEAKY, datt2_bf16=None, gt=None, tag="attn_bwd"):
    lib = _lib.load()
    A, E = att1.shape[1], Fe.shape[1]
    tok = TIMER.begin(tag) if (TIMER is not None and tag) else None
    check(lib.st_attn_step_bwd(rows, Pn, A, E, _raw(att1), _raw(Fe), int(att1.dtype == BF16), _raw(att2),
                               ptr(wf, F32), _raw(alphas_t), alpha_stride,
                               _raw(dalpha) if dalpha is not None else None, dalpha_stride, _raw(dctx),
                               dctx.stride(0), _raw(de_out), _raw(datt2),
                               _raw(datt2_bf16) if datt2_bf16 is not None else None,
                               _raw(gt) if gt is not None else None, act, stream_ptr()),
          "st_attn_step_bwd")
    if tok:
        TIMER.end(tok)


def attn_hoist_bwd(bs, Pn, att1, att2_all, de_all, wf, want_t=True, act=ACT_LEAKY, gt_all=None):
    """Returns (datt1 (B*P, A), datt1T (A, B*P) or None, dwf (A,)) in att1's storage type.  gt_all (N, A): the tensor
    attn_step_bwd(gt=...) filled during the reverse loop (optional; lets the pass skip a third of its work)."""
    lib = _lib.load()
    A = att1.shape[1]
    BP = att1.shape[0]
    dev = att1.device
    datt1 = torch.empty_like(att1)
    ld = (BP + 7) // 8 * 8
    dT = torch.empty(A, ld, dtype=att1.dtype, device=dev)[:, :BP] if want_t else None
    dwf = torch.empty(A, dtype=F32, device=dev)
    isb = int(att1.dtype == BF16)
    check(lib.st_attn_hoist_bwd(len(bs), int_array(bs), Pn, A, _raw(att1), isb, ptr(att2_all, F32),
                                ptr(de_all, F32), ptr(wf, F32), _raw(datt1), _raw(dT) if want_t else None, ld,
                                isb, ptr(dwf), ptr(gt_all, F32) if gt_all is not None else None, act, stream_ptr()),
          "st_attn_hoist_bwd")
    return datt1, dT, dwf


def attn_ctx_all(bs, Pn, F, alphas, want=True, want_t=False, out_dtype=None):
    """ctx (N, C) and/or ctxT (C, N) in F's storage type (or `out_dtype`); alphas (B, Tcap, P) fp32."""
    lib = _lib.load()
    N, Cc = sum(bs), F.shape[1]
    dev = F.device
    dt = out_dtype or F.dtype
    ctx = torch.empty(N, Cc, dtype=dt, device=dev) if want else None
    ld = (N + 7) // 8 * 8
    cT = torch.empty(Cc, ld, dtype=dt, device=dev)[:, :N] if want_t else None
    check(lib.st_attn_ctx_all(len(bs), int_array(bs), Pn, Cc, alphas.shape[1], _raw(F), int(F.dtype == BF16),
                              ptr(alphas, F32), _raw(ctx) if want else None, _raw(cT) if want_t else None, ld,
                              int(dt == BF16), stream_ptr()), "st_attn_ctx_all")
    return ctx, cT


def attn_embed_q(bs, Pn, alphas, dctx):
    """Q (B*P, E) bf16 = sum_t alphas[b,t,p] dctx[(t,b),:]: dW_embed = gemm_bf16(Q, F, a_t=True, b_t=True)."""
    lib = _lib.load()
    B, E = bs[0], dctx.shape[1]
    if E % 8:
        raise ValueError("attn_embed_q: E must be a multiple of 8 (bf16 TMA operand rows)")
    Q = torch.empty(B * Pn, E, dtype=BF16, device=dctx.device)
    check(lib.st_attn_embed_q(len(bs), int_array(bs), Pn, E, alphas.shape[1], ptr(alphas, F32), ptr(dctx, F32),
                              dctx.stride(0), _raw(Q), stream_ptr()), "st_attn_embed_q")
    return Q


def attn_penalty(S, coef):
    lib = _lib.load()
    pen = torch.empty(1, dtype=F32, device=S.device)
    G = torch.empty_like(S)
    check(lib.st_attn_penalty(S.numel(), ptr(S, F32), float(coef), ptr(pen), ptr(G), stream_ptr()),
          "st_attn_penalty")
    return pen, G


def add_rows(dst, src, rows):
    lib = _lib.load()
    check(lib.st_add_rows(_raw(dst), _raw(src), rows, dst.shape[-1], stream_ptr()), "st_add_rows")
