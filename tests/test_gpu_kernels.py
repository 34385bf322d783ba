"""GPU: each C-ABI kernel against a straightforward torch float64 computation of the same
operation (these are unit checks of the plumbing; parity with the reference path is in
test_gpu_base.py / test_gpu_attn.py)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from helpers import rel_err


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda:0")


@pytest.mark.parametrize("ta,tb", [(False, False), (False, True), (True, False), (True, True)])
@pytest.mark.parametrize("M,N,K", [(1, 1, 1), (5, 7, 3), (128, 128, 16), (129, 257, 33), (300, 70, 515)])
def test_sgemm(dev, ta, tb, M, N, K):
    from showtell_b200 import ops
    g = torch.Generator(device="cpu").manual_seed(M * 1000 + N * 10 + K)
    A = torch.randn((K, M) if ta else (M, K), generator=g).to(dev)
    B = torch.randn((N, K) if tb else (K, N), generator=g).to(dev)
    bias = torch.randn(N, generator=g).to(dev)
    C0 = torch.randn(M, N, generator=g).to(dev)
    ref = (A.double().t() if ta else A.double()) @ (B.double().t() if tb else B.double())
    out = ops.sgemm(A, B, transA=ta, transB=tb)
    assert rel_err(out, ref) < 1e-5
    out2 = ops.sgemm(A, B, transA=ta, transB=tb, bias=bias, out=C0.clone(), alpha=0.5, beta=2.0)
    assert rel_err(out2, 0.5 * ref + 2.0 * C0.double() + bias.double()) < 1e-5


def test_sgemm_strided_views(dev):
    from showtell_b200 import ops
    big = torch.randn(40, 100, device=dev)
    A = big[:, 10:42]                       # (40, 32) view, row stride 100
    W = torch.randn(19, 64, device=dev)[:, 32:]
    out = ops.sgemm(A, W, transB=True)
    assert rel_err(out, A.double() @ W.double().t()) < 1e-5


def _caps(B, T, V, lengths, seed):
    g = torch.Generator().manual_seed(seed)
    cap = torch.zeros(B, T, dtype=torch.int64)
    for b, l in enumerate(lengths):
        cap[b, :l] = torch.randint(0, V, (l,), generator=g)
    return cap


def test_pack_inputs_targets_and_bwd(dev):
    from showtell_b200 import _lib, ops
    B, T, V, E = 6, 7, 23, 10
    lengths = [7, 7, 5, 3, 3, 1]
    bs = _lib.batch_sizes(lengths)
    cap = _caps(B, T, V, lengths, 3)
    emb = torch.randn(V, E)
    feat = torch.randn(B, E)
    for with_feature in (True, False):
        X = ops.pack_inputs(emb.to(dev), feat.to(dev), cap.to(dev), bs, with_feature).cpu()
        rows = []
        for t, b_t in enumerate(bs):
            for b in range(b_t):
                if with_feature:
                    rows.append(feat[b] if t == 0 else emb[cap[b, t - 1]])
                else:
                    rows.append(emb[cap[b, t]])
        assert torch.equal(X, torch.stack(rows))
        dX = torch.randn(len(rows), E)
        dEmb = torch.zeros(V, E, device=dev)
        dfeat = torch.zeros(B, E, device=dev)
        ops.pack_inputs_bwd(dX.to(dev), dEmb, dfeat, cap.to(dev), bs, with_feature)
        ref_e, ref_f, n = torch.zeros(V, E, dtype=torch.float64), torch.zeros(B, E, dtype=torch.float64), 0
        for t, b_t in enumerate(bs):
            for b in range(b_t):
                if with_feature and t == 0:
                    ref_f[b] = dX[n]
                else:
                    ref_e[cap[b, t - 1 if with_feature else t]] += dX[n].double()
                n += 1
        assert rel_err(dEmb, ref_e) < 1e-6
        if with_feature:
            assert rel_err(dfeat, ref_f) < 1e-7
    tg = ops.pack_targets(cap.to(dev), bs, V).cpu()
    ref = torch.nn.utils.rnn.pack_padded_sequence(cap, lengths, batch_first=True)[0]
    assert torch.equal(tg, ref)


def test_unsorted_lengths_raise():
    from showtell_b200 import _lib
    with pytest.raises(RuntimeError):
        _lib.batch_sizes([3, 5, 2])
    lib = _lib.load()
    import ctypes as C
    bad = _lib.int_array([2, 3])
    st = lib.st_pack_targets(C.c_void_p(8), C.c_void_p(8), 4, 10, 2, bad, None)
    assert st == -2 and "non-increasing" in _lib.last_error()


def test_bad_token_ids_are_reported_not_dereferenced(dev):
    """An id outside [0, V) (the reference's nn.Embedding / CrossEntropyLoss device-assert on it) is clamped for the
    access and raises IndexError at the next call boundary; the gradient scatter stays inside dEmb."""
    from showtell_b200 import _lib, ops
    from showtell_b200.rnn import RNN
    V, E, B, T = 11, 8, 3, 4
    emb = torch.randn(V, E, device=dev)
    cap = torch.tensor([[1, 5, V + 7, 2], [1, -3, 2, 0], [1, 2, 0, 0]], device=dev)
    bs = _lib.batch_sizes([4, 3, 2])
    _lib.load().st_token_error(None, 1)
    X = ops.pack_inputs(emb, None, cap, bs, False)
    torch.cuda.synchronize()
    assert torch.equal(X[bs[0] + bs[1]], emb[V - 1]) and torch.equal(X[bs[0] + 1], emb[0])     # clamped rows
    with pytest.raises(IndexError, match="outside"):
        ops.pack_inputs(emb, None, cap, bs, False)
    tg = ops.pack_targets(cap, bs, V)
    torch.cuda.synchronize()
    assert int(tg.min()) >= 0 and int(tg.max()) < V
    dEmb = torch.zeros(V + 64, E, device=dev)[:V]                   # guard rows behind the table stay zero
    ops.pack_inputs_bwd(torch.ones(sum(bs), E, device=dev), dEmb, None, cap, bs, False)
    torch.cuda.synchronize()
    assert float(dEmb.sum()) == sum(bs) * E
    with pytest.raises(IndexError):
        _lib.raise_token_error()
    _lib.raise_token_error()                                         # cleared
    # through the module: the step after the bad one raises
    m = RNN(E, 8, V, 1).to(dev)
    feat = torch.randn(B, E, device=dev)
    m.forward_loss(feat, cap, [4, 3, 2]).backward()
    torch.cuda.synchronize()
    with pytest.raises(IndexError):
        m.forward_loss(feat, cap.clamp(0, V - 1), [4, 3, 2])
    m.forward_loss(feat, cap.clamp(0, V - 1), [4, 3, 2]).backward()


@pytest.mark.parametrize("rows,cols", [(1, 1), (300, 37), (1000, 130)])
def test_colsum(dev, rows, cols):
    from showtell_b200 import ops
    M = torch.randn(rows, cols, device=dev)
    assert rel_err(ops.colsum(M), M.double().sum(0)) < 1e-5
    acc = torch.ones(cols, device=dev)
    ops.colsum(M, out=acc, accumulate=True)
    assert rel_err(acc, M.double().sum(0) + 1) < 1e-5


@pytest.mark.parametrize("rows,cols,ld", [(5120, 10000, 10000), (300, 2048, 2048), (1000, 777, 784), (33, 64, 72), (9, 130, 131)])
def test_colsum_bf16(dev, rows, cols, ld):
    """bf16 column sums (db_v from the dlogits, db_ih / db_hh from the gate gradients): vector path and fallback."""
    from showtell_b200 import ops
    M = torch.randn(rows, ld, device=dev).bfloat16()[:, :cols]
    assert rel_err(ops.colsum(M), M.double().sum(0)) < 2e-5
    acc = torch.ones(cols, device=dev)
    ops.colsum(M, out=acc, accumulate=True)
    assert rel_err(acc, M.double().sum(0) + 1) < 2e-5


def test_bf16_shadows_follow_the_parameter(dev):
    """ops.bf16_shadow: one persistent bf16 copy per parameter, re-cast when torch's version counter moves (optimizer
    steps, load_state_dict), rewritten in place by showtell_b200.optim's own update kernel."""
    from showtell_b200 import ops, optim
    w = torch.nn.Parameter(torch.randn(48, 64, device=dev))
    a = ops.bf16_shadow(w)
    at = ops.bf16_shadow(w, transposed=True)
    assert torch.equal(a, w.detach().bfloat16()) and torch.equal(at, w.detach().t().bfloat16())
    assert ops.bf16_shadow(w.detach()).data_ptr() == a.data_ptr()            # detached alias: same shadow
    with torch.no_grad():
        w.mul_(2.0)                                                          # torch-side in-place update
    b = ops.bf16_shadow(w)
    assert b.data_ptr() == a.data_ptr() and torch.equal(b, w.detach().bfloat16())
    w.grad = torch.randn_like(w)
    for opt in (optim.SGD([w], lr=0.1, momentum=0.9), optim.Adam([w], lr=0.1)):
        opt.step()                                                           # raw-pointer update: shadow written in the kernel
        torch.cuda.synchronize()
        assert torch.equal(a, w.detach().bfloat16())
        assert torch.equal(ops.bf16_shadow(w, transposed=True), w.detach().t().bfloat16())   # stale-marked, re-cast here
    v = w[:, 16:48]                                                          # a view (attention: halves of W_ih)
    sv = ops.bf16_shadow(v.detach())
    assert torch.equal(sv, v.detach().bfloat16())
    opt.step()
    torch.cuda.synchronize()
    assert torch.equal(ops.bf16_shadow(v.detach()), v.detach().bfloat16())


@pytest.mark.parametrize("N,V", [(3, 5), (40, 1000), (7, 10000)])
def test_ce(dev, N, V):
    from showtell_b200 import ops
    logits = (torch.randn(N, V) * 3).to(dev)
    tgt = torch.randint(0, V, (N,)).to(dev)
    ls, lse, dl = ops.ce_fwd_bwd(logits, tgt, grad_scale=1.0 / N)
    ld = logits.double().requires_grad_(True)
    ref = torch.nn.functional.cross_entropy(ld, tgt, reduction="sum")
    ref.backward()
    assert abs(float(ls) - float(ref)) / float(ref) < 1e-5
    assert rel_err(lse, torch.logsumexp(logits.double(), 1)) < 1e-6
    assert rel_err(dl, ld.grad / N) < 1e-5
    ls2, _, dl2 = ops.ce_fwd_bwd(logits.clone(), tgt, grad_scale=1.0 / N, inplace=True)
    assert torch.equal(dl2, dl) and float(ls2) == pytest.approx(float(ls), rel=1e-6)


def test_argmax_topk(dev):
    from showtell_b200 import ops
    X = torch.randn(17, 10000, device=dev)
    X[3, 77] = X[3, 5000] = 50.0                 # tie -> lowest index (Tensor.max(1)[1] on CPU)
    idx = ops.argmax_rows(X)
    ref = X.cpu().max(1)[1]
    assert torch.equal(idx.cpu(), ref) and int(idx[3]) == 77
    val, ti = ops.topk_rows(X, 5)
    rv, ri = X.topk(5, dim=1)
    assert torch.equal(val, rv)
    assert torch.equal(ti[:3].long(), ri[:3]) and torch.equal(ti[4:].long(), ri[4:])
    assert ti[3, :2].tolist() == [77, 5000]


def _ref_seq(kind, Gx, Whh, bhh, bs, H, h0=None, c0=None):
    """float64 torch loop of the recurrence over a packed sequence."""
    B0 = bs[0]
    h = torch.zeros(B0, H, dtype=torch.float64) if h0 is None else h0.double()
    c = torch.zeros(B0, H, dtype=torch.float64) if c0 is None else c0.double()
    outs, cs, off = [], [], 0
    for b in bs:
        gx = Gx[off:off + b]
        off += b
        h, c = h[:b], c[:b]
        gh = h @ Whh.t() + bhh
        if kind == "lstm":
            a = gx + gh
            i, f, g, o = [a[:, k * H:(k + 1) * H] for k in range(4)]
            c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(g)
            h = torch.sigmoid(o) * torch.tanh(c)
        else:
            r = torch.sigmoid(gx[:, :H] + gh[:, :H])
            z = torch.sigmoid(gx[:, H:2 * H] + gh[:, H:2 * H])
            n = torch.tanh(gx[:, 2 * H:] + r * gh[:, 2 * H:])
            h = (1 - z) * n + z * h
        outs.append(h)
        cs.append(c)
    return torch.cat(outs), torch.cat(cs)


@pytest.mark.parametrize("kind", ["gru", "lstm"])
@pytest.mark.parametrize("H,lengths,init", [
    (8, [3, 2, 2, 1], False),
    (36, [5] * 7, True),
    (64, sorted([9, 9, 8, 6, 5, 3] * 30, reverse=True), False),          # 180 rows -> two batch tiles, ragged
    (512, [20] * 40 + [13] * 24, True),            # full-size hidden state
])
def test_rnn_seq_fwd_bwd(dev, kind, H, lengths, init):
    from showtell_b200 import _lib, ops
    k = _lib.ST_LSTM if kind == "lstm" else _lib.ST_GRU
    G = 4 if kind == "lstm" else 3
    bs = _lib.batch_sizes(lengths)
    N, B0 = sum(bs), bs[0]
    g = torch.Generator().manual_seed(H + len(lengths))
    s = 1.0 / H ** 0.5
    Gx = torch.randn(N, G * H, generator=g, dtype=torch.float64)
    Whh = (torch.rand(G * H, H, generator=g, dtype=torch.float64) * 2 - 1) * s * 2
    bhh = (torch.rand(G * H, generator=g, dtype=torch.float64) * 2 - 1) * s
    h0 = torch.randn(B0, H, generator=g, dtype=torch.float64) * 0.5 if init else None
    c0 = torch.randn(B0, H, generator=g, dtype=torch.float64) * 0.5 if (init and kind == "lstm") else None
    dHs = torch.randn(N, H, generator=g, dtype=torch.float64)

    leaves = [t.clone().requires_grad_(True) for t in (Gx, Whh, bhh)]
    h0r = h0.clone().requires_grad_(True) if h0 is not None else None
    c0r = c0.clone().requires_grad_(True) if c0 is not None else None
    Hs_ref, Cs_ref = _ref_seq(kind, leaves[0], leaves[1], leaves[2], bs, H, h0r, c0r)
    (Hs_ref * dHs).sum().backward()

    f = lambda t: None if t is None else t.float().to(dev)
    out = ops.rnn_seq_fwd(k, f(Gx), f(Whh), f(bhh), bs, h0=f(h0), c0=f(c0))
    assert rel_err(out["Hs"], Hs_ref) < 2e-5
    if kind == "lstm":
        assert rel_err(out["Cs"], Cs_ref) < 2e-5
    b = ops.rnn_seq_bwd(k, f(Whh), bs, out, f(dHs), h0=f(h0), c0=f(c0))
    assert rel_err(b["dG"], leaves[0].grad) < 1e-4
    Hprev = ops.shift_states(out["Hs"], bs, f(h0))
    dWhh = ops.sgemm(b["dGh"], Hprev, transA=True)
    assert rel_err(dWhh, leaves[1].grad) < 1e-4
    assert rel_err(ops.colsum(b["dGh"]), leaves[2].grad) < 1e-4
    if h0 is not None:
        assert rel_err(b["dstate"][0], h0r.grad) < 1e-4
    if c0 is not None:
        assert rel_err(b["dstate"][1], c0r.grad) < 1e-4


def test_rnn_seq_stepwise_equals_whole(dev):
    """Single-step calls (decode / attention use) must reproduce the persistent whole-sequence run."""
    from showtell_b200 import _lib, ops
    H, lengths = 64, [6, 6, 4, 2]
    bs = _lib.batch_sizes(lengths)
    N = sum(bs)
    for k, G in ((_lib.ST_GRU, 3), (_lib.ST_LSTM, 4)):
        Gx = torch.randn(N, G * H, device=dev)
        Whh = torch.randn(G * H, H, device=dev) * 0.1
        bhh = torch.randn(G * H, device=dev) * 0.1
        whole = ops.rnn_seq_fwd(k, Gx, Whh, bhh, bs)
        st = None
        for t in range(len(bs)):
            st = ops.rnn_seq_fwd(k, Gx, Whh, bhh, bs, t_range=(t, t + 1), out=st)
        assert torch.equal(st["Hs"], whole["Hs"])
        dHs = torch.randn(N, H, device=dev)
        bw = ops.rnn_seq_bwd(k, Whh, bs, whole, dHs)
        bst = None
        for t in reversed(range(len(bs))):
            bst = ops.rnn_seq_bwd(k, Whh, bs, whole, dHs, t_range=(t + 1, t), out=bst)
        assert torch.equal(bst["dG"], bw["dG"]) and torch.equal(bst["dstate"], bw["dstate"])


def test_scale_multi(dev):
    """The chain rule of forward_loss.backward(): [t * g] for a list of gradient tensors in one launch."""
    from showtell_b200 import _lib, ops
    gen = torch.Generator().manual_seed(5)
    shapes = [(10000, 512), (2048,), (1, 512), (37, 3), (5,), (4097,), (0,)]
    ts = [torch.randn(*s, generator=gen).to(dev) for s in shapes]
    ts.append(torch.randn(1001, generator=gen).to(dev)[1:])          # not 16-byte aligned
    g = torch.tensor(0.37, device=dev)
    lib = _lib.load()
    l0 = lib.st_launch_count()
    out = ops.scale_multi(ts, g)
    assert lib.st_launch_count() - l0 == 1
    for o, t in zip(out, ts):
        assert o.shape == t.shape and torch.equal(o, t * g)
    many = [torch.randn(17, generator=gen).to(dev) for _ in range(40)]  # more than one table
    for o, t in zip(ops.scale_multi(many, g), many):
        assert torch.equal(o, t * g)


@pytest.mark.parametrize("lengths,P,E", [([5, 5, 4, 2, 1], 30, 64), ([20] * 9, 196, 512), ([18, 7, 3], 49, 72)])
def test_attn_embed_q(dev, lengths, P, E):
    """Q[(b,p), e] = sum_t alphas[b,t,p] dctx[(t,b), e] over each row's live steps (the A operand of dW_embed = Q^T F)."""
    from showtell_b200 import _lib, ops
    bs = _lib.batch_sizes(lengths)
    B, T, N = len(lengths), len(bs), sum(bs)
    g = torch.Generator().manual_seed(P + E)
    alphas = torch.rand(B, T + 2, P, generator=g)
    dctx = torch.randn(N, E, generator=g)
    ref = torch.zeros(B, P, E, dtype=torch.float64)
    off = 0
    for t, bt in enumerate(bs):
        ref[:bt] += alphas[:bt, t, :, None].double() * dctx[off:off + bt, None, :].double()
        off += bt
    Q = ops.attn_embed_q(bs, P, alphas.to(dev), dctx.to(dev))
    assert Q.shape == (B * P, E) and Q.dtype == torch.bfloat16
    assert rel_err(Q.float(), ref.reshape(B * P, E)) < 5e-3
