"""GPU: tcgen05/TMEM/TMA GEMM and its fused cross-entropy epilogues against torch float64 on the
same bf16-rounded operands (so the only difference is fp32 accumulation order)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from helpers import rel_err

DEV = "cuda:0"


@pytest.fixture
def gemm_variant():
    """Pins the kernel behind st_gemm_bf16 / st_vocab_ce_* (st_debug_gemm_variant) and restores the choice."""
    from showtell_b200 import _lib
    lib = _lib.load()
    yield lambda v: lib.st_debug_gemm_variant(v)
    lib.st_debug_gemm_variant(0)


# variant: 0 = the library's choice, 128 / 256 = single-CTA kernel, 2 = CTA-pair kernel (cta_group::2);
# | 0x4000 = B-multicast CTA pairs wherever the shape allows them, | 0x8000 = never
MC, NOMC = 0x4000, 0x8000


@pytest.mark.parametrize("variant", [0, 128, 256, 2, MC | 128, MC | 256, NOMC])
@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (128, 128, 512), (256, 384, 128), (5120, 2048, 512),
                                   (100, 50, 40), (129, 257, 72), (640, 10000, 512), (37, 1000, 2048),
                                   (5120, 512, 10000),      # dX of the vocabulary projection: stream-K, 40 tiles
                                   (10000, 512, 5120),      # dW of the vocabulary projection: stream-K, ragged M
                                   (2048, 512, 5120),       # dW_hh: 16 tiles cut into 74 ranges
                                   (300, 700, 4000)])       # stream-K with ragged M, N and K
def test_gemm_bf16(M, N, K, variant, gemm_variant):
    from showtell_b200 import ops
    gemm_variant(variant)
    g = torch.Generator().manual_seed(M + N + K)
    Kp = (K + 7) // 8 * 8
    A = torch.randn(M, Kp, generator=g).to(DEV).bfloat16()[:, :K]
    B = torch.randn(N, Kp, generator=g).to(DEV).bfloat16()[:, :K]
    bias = torch.randn(N, generator=g).to(DEV)
    ref = A.double() @ B.double().t()
    tol = 1e-5 if K <= 2048 else 4e-5                  # fp32 accumulation over K terms
    out = ops.gemm_bf16(A, B)
    assert out.dtype == torch.float32 and rel_err(out, ref) < tol
    out2 = ops.gemm_bf16(A, B, bias=bias, alpha=0.5)
    assert rel_err(out2, 0.5 * ref + bias.double()) < tol
    out3 = ops.gemm_bf16(A, B, bias=bias, out_dtype=torch.bfloat16)
    assert out3.dtype == torch.bfloat16 and rel_err(out3, ref + bias.double()) < 1e-2
    acc = torch.randn(M, N, generator=g).to(DEV)
    out4 = ops.gemm_bf16(A, B, beta=1.0, out=acc.clone())
    assert rel_err(out4, ref + acc.double()) < tol
    # a view into a wider buffer (ldc > N): columns outside the view are untouched (stream-K zeroes C itself)
    wide = torch.full((M, N + 12), 7.0, device=DEV)
    ops.gemm_bf16(A, B, bias=bias, out=wide[:, 4:4 + N])
    assert rel_err(wide[:, 4:4 + N], ref + bias.double()) < tol
    assert bool((wide[:, :4] == 7).all()) and bool((wide[:, 4 + N:] == 7).all())


@pytest.mark.parametrize("variant", [0, 128, 256, 2, MC | 128, MC | 256, NOMC])
@pytest.mark.parametrize("a_t,b_t", [(True, True), (False, True), (True, False)])
@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (128, 256, 128), (256, 384, 192), (100, 50, 40), (129, 257, 72),
                                   (10000, 512, 5120),      # dW of the vocabulary projection: P^T . Hs
                                   (5120, 512, 10000),      # dX of the vocabulary projection: P . W_v
                                   (2048, 512, 5120),       # dW_hh / dW_ih
                                   (300, 700, 4000)])
def test_gemm_bf16_mn_major_operands(M, N, K, a_t, b_t, variant, gemm_variant):
    """Operands passed as their transposes ((K, M) / (K, N) row-major) and consumed in place as MN-major UMMA
    operands: the backward products dW = dY^T X and dX = dY W without transposed copies."""
    from showtell_b200 import ops
    gemm_variant(variant)
    g = torch.Generator().manual_seed(M + 2 * N + 3 * K)
    pad = lambda n: (n + 7) // 8 * 8
    A = torch.randn(M, K, generator=g).to(DEV).bfloat16()
    B = torch.randn(N, K, generator=g).to(DEV).bfloat16()
    ref = A.double() @ B.double().t()
    Ain = torch.zeros(K, pad(M), device=DEV, dtype=torch.bfloat16)[:, :M].copy_(A.t()) if a_t else \
        torch.zeros(M, pad(K), device=DEV, dtype=torch.bfloat16)[:, :K].copy_(A)
    Bin = torch.zeros(K, pad(N), device=DEV, dtype=torch.bfloat16)[:, :N].copy_(B.t()) if b_t else \
        torch.zeros(N, pad(K), device=DEV, dtype=torch.bfloat16)[:, :K].copy_(B)
    tol = 1e-5 if K <= 2048 else 4e-5
    out = ops.gemm_bf16(Ain, Bin, a_t=a_t, b_t=b_t)
    assert out.shape == (M, N) and rel_err(out, ref) < tol, rel_err(out, ref)
    bias = torch.randn(N, generator=g).to(DEV)
    acc = torch.randn(M, N, generator=g).to(DEV)
    out2 = ops.gemm_bf16(Ain, Bin, a_t=a_t, b_t=b_t, bias=bias, alpha=0.5, beta=1.0, out=acc.clone())
    assert rel_err(out2, 0.5 * ref + bias.double() + acc.double()) < tol
    out3 = ops.gemm_bf16(Ain, Bin, a_t=a_t, b_t=b_t, out_dtype=torch.bfloat16)
    assert rel_err(out3, ref) < 1e-2
    with pytest.raises(ValueError):
        ops.gemm_bf16(Ain, Bin, a_t=a_t, b_t=b_t, out=torch.empty(M + 1, N, device=DEV))


@pytest.mark.parametrize("M,N,K", [(128, 128, 32), (100, 50, 40), (129, 257, 72), (640, 10000, 512), (4096, 1536, 512),
                                   (37, 1000, 2048), (4096, 10000, 512)])
def test_gemm_tf32x3_is_fp32_accurate(M, N, K):
    """The 3xTF32 tensor-core GEMM of the decoding loops against float64: fp32-level accuracy.  The operand
    split is exact to ~2^-22; what remains is the tensor core's accumulator, which truncates instead of rounding
    (K/8 accumulation steps): measured 3e-6 of max|C| at K = 512 against 6e-7 for the CUDA-core fp32 GEMM and
    1.5e-3 for a plain tf32 / bf16 product."""
    from showtell_b200 import ops
    g = torch.Generator().manual_seed(M + N + K)
    A = torch.randn(M, K, generator=g).to(DEV)
    B = (torch.randn(N, K, generator=g) / K ** 0.5).to(DEV)
    bias = torch.randn(N, generator=g).to(DEV)
    ref = A.double() @ B.double().t() + bias.double()
    hiA, loA = ops.split_tf32(A)
    assert torch.equal(hiA + loA, A) and bool(((hiA.view(torch.int32) & 0x1fff) == 0).all())
    out = ops.gemm_tf32x3((hiA, loA), ops.split_tf32(B), bias=bias)
    e_tc = rel_err(out, ref)
    e_sg = rel_err(ops.sgemm(A, B, transB=True, bias=bias), ref)
    print(f"tf32x3 {M}x{N}x{K}: max-norm error {e_tc:.2e} (CUDA-core fp32 GEMM {e_sg:.2e})")
    assert e_tc < 1e-6 * max(K / 64, 1.0) ** 0.75 and e_tc < 10 * e_sg + 2e-7, (e_tc, e_sg)
    acc = torch.randn(M, N, generator=g).to(DEV)
    out2 = ops.gemm_tf32x3((hiA, loA), ops.split_tf32(B), alpha=0.5, beta=1.0, out=acc.clone())
    assert rel_err(out2, 0.5 * (ref - bias.double()) + acc.double()) < 1e-5


def test_cast_bf16():
    from showtell_b200 import ops
    x = torch.randn(70, 45, device=DEV)
    d, dT = ops.cast_bf16(x, True, True)
    assert torch.equal(d, x.bfloat16()) and torch.equal(dT, x.bfloat16().t())
    assert d.stride(0) % 8 == 0 and dT.stride(0) % 8 == 0


@pytest.mark.parametrize("variant", [0, 128, 256, 2, MC | 256, NOMC])
@pytest.mark.parametrize("M,V,H", [(64, 128, 64), (300, 1000, 128), (640, 10000, 512), (5120, 10000, 512),
                                   (2555, 9999, 512)])
def test_vocab_ce_fused(M, V, H, variant, gemm_variant):
    from showtell_b200 import ops
    gemm_variant(variant)
    g = torch.Generator().manual_seed(M + V)
    Hs = (torch.randn(M, H, generator=g) * 0.5).to(DEV).bfloat16()
    Wv = (torch.randn(V, H, generator=g) * 0.1).to(DEV).bfloat16()
    bv = (torch.randn(V, generator=g) * 0.1).to(DEV)
    tgt = torch.randint(0, V, (M,), generator=g).to(DEV)
    logits = (Hs.double() @ Wv.double().t() + bv.double()).requires_grad_(True)
    ref = torch.nn.functional.cross_entropy(logits, tgt, reduction="sum")
    ref.backward()
    loss, lse = ops.vocab_ce_fwd(Hs, Wv, bv, tgt)
    assert abs(float(loss) - float(ref)) < 1e-5 * abs(float(ref))
    assert rel_err(lse, torch.logsumexp(logits.detach(), 1)) < 1e-5
    scale = 1.0 / M
    P, PT = ops.vocab_ce_bwd(Hs, Wv, bv, tgt, lse, scale)
    dref = logits.grad * scale
    assert rel_err(P, dref) < 1e-2                     # bf16 storage of dlogits
    assert torch.equal(PT, P.t())
    # fp32-level check of the recompute: row sums of dlogits vanish
    assert float(P.float().sum(1).abs().max()) < 2e-2 * scale


@pytest.fixture
def bwd_ks():
    """st_debug_set_bwd_ks: pin the BPTT kernel's K split (1 = none, 4 = clusters of 4 unit tiles); restored after."""
    from showtell_b200 import _lib
    lib = _lib.load()
    yield lambda ks: lib.st_debug_set_bwd_ks(int(ks))
    lib.st_debug_set_bwd_ks(0)


@pytest.mark.parametrize("ks", [1, 4 | (64 << 8), 4 | (128 << 8)])     # low byte: K split; bits 8..: batch-tile height
@pytest.mark.parametrize("kind", ["gru", "lstm"])
@pytest.mark.parametrize("H,lengths,init", [
    (64, [5, 5, 4, 2], False),
    (128, sorted([9, 9, 8, 6, 5, 3] * 30, reverse=True), True),    # 180 rows: two batch tiles, ragged
    (256, sorted([6, 5, 5, 3] * 20, reverse=True), True),          # GRU: 3 k-blocks per CTA of a K-split cluster
    (512, [20] * 200 + [13] * 56, False),                          # config-2 shape
])
def test_rnn_seq_tensor_core_vs_cuda_core(kind, H, lengths, init, ks, bwd_ks):
    """The tcgen05 persistent recurrence (bf16 operands, fp32 state) against the fp32 CUDA-core
    kernels on the same bf16-rounded W_hh: forward states, gate gradients, dh0/dc0, bias grads.
    ks: the BPTT kernel with every CTA streaming all of K (1) / K split over clusters of 4 unit tiles (4)."""
    from showtell_b200 import _lib, ops
    bwd_ks(ks)
    k = _lib.ST_LSTM if kind == "lstm" else _lib.ST_GRU
    G = 4 if kind == "lstm" else 3
    bs = _lib.batch_sizes(lengths)
    N, B0 = sum(bs), bs[0]
    g = torch.Generator().manual_seed(H)
    s = 1.0 / H ** 0.5
    Gx = torch.randn(N, G * H, generator=g).to(DEV)
    Whh = ((torch.rand(G * H, H, generator=g) * 2 - 1) * s).to(DEV).bfloat16().float()
    bhh = ((torch.rand(G * H, generator=g) * 2 - 1) * s).to(DEV)
    h0 = (torch.randn(B0, H, generator=g) * 0.5).to(DEV).bfloat16().float() if init else None
    c0 = (torch.randn(B0, H, generator=g) * 0.5).to(DEV) if (init and kind == "lstm") else None
    dHs = torch.randn(N, H, generator=g).to(DEV)
    ref = ops.rnn_seq_fwd(k, Gx, Whh, bhh, bs, h0=h0, c0=c0)
    Wb, WT = ops.cast_bf16(Whh, True, True)
    out = ops.rnn_seq_tc_fwd(k, Gx, Wb, bhh, bs, h0=h0, h0_b=None if h0 is None else h0.bfloat16(), c0=c0)
    assert out is not None, "tensor-core recurrent kernel reported unsupported"
    assert rel_err(out["Hs"], ref["Hs"]) < 1e-2
    assert torch.equal(out["Hsb"], out["Hs"].bfloat16())
    if kind == "lstm":
        assert rel_err(out["Cs"], ref["Cs"]) < 1e-2
    assert rel_err(out["gates"], ref["gates"]) < 1e-2
    # backward on the tensor-core forward's own saved state
    rb = ops.rnn_seq_bwd(k, Whh, bs, out, dHs, h0=h0, c0=c0)
    tb = ops.rnn_seq_tc_bwd(k, WT, bs, out, dHs, h0=h0, c0=c0, transposed=True)
    assert tb is not None
    tn = ops.rnn_seq_tc_bwd(k, WT, bs, out, dHs, h0=h0, c0=c0)          # default: row-major gate gradients only
    assert tn["dGT"] is None and torch.equal(tn["dGb"], tb["dGb"]) and torch.equal(tn["dstate"], tb["dstate"])
    assert rel_err(ops.colsum(tn["dGb"]), ops.colsum(rb["dG"])) < 2e-2
    assert rel_err(tb["dGb"], rb["dG"]) < 2e-2
    assert rel_err(tb["dGhb"], rb["dGh"]) < 2e-2
    assert torch.equal(tb["dGT"], tb["dGb"].t()) and torch.equal(tb["dGhT"], tb["dGhb"].t())
    assert rel_err(tb["dstate"][0], rb["dstate"][0]) < 2e-2
    if kind == "lstm":
        assert rel_err(tb["dstate"][1], rb["dstate"][1]) < 2e-2
    assert rel_err(tb["dbih"], ops.colsum(rb["dG"])) < 2e-2
    assert rel_err(tb["dbhh"], ops.colsum(rb["dGh"])) < 2e-2


@pytest.mark.parametrize("ks", [1, 4 | (64 << 8), 4 | (128 << 8)])
@pytest.mark.parametrize("kind", ["gru", "lstm"])
def test_rnn_seq_tensor_core_stepwise_equals_whole(kind, ks, bwd_ks, monkeypatch):
    """Partial step ranges (used by the attention loop) resume from the packed state rows and must
    reproduce the whole-sequence persistent run bit for bit."""
    from showtell_b200 import _lib, ops
    bwd_ks(ks)
    k = _lib.ST_LSTM if kind == "lstm" else _lib.ST_GRU
    G = 4 if kind == "lstm" else 3
    H, lengths = 128, [7, 7, 6, 4, 4, 2]
    bs = _lib.batch_sizes(lengths)
    N, B0 = sum(bs), bs[0]
    g = torch.Generator().manual_seed(5)
    Gx = torch.randn(N, G * H, generator=g).to(DEV)
    Whh = (torch.randn(G * H, H, generator=g) * 0.08).to(DEV)
    bhh = (torch.randn(G * H, generator=g) * 0.1).to(DEV)
    h0 = (torch.randn(B0, H, generator=g) * 0.5).to(DEV)
    c0 = (torch.randn(B0, H, generator=g) * 0.5).to(DEV) if kind == "lstm" else None
    dHs = torch.randn(N, H, generator=g).to(DEV)
    Wb, WT = ops.cast_bf16(Whh, True, True)
    h0b = h0.bfloat16()
    monkeypatch.setattr(ops, "USE_CLUSTER", False)       # the device-wide tc kernel on both sides
    whole = ops.rnn_seq_tc_fwd(k, Gx, Wb, bhh, bs, h0=h0, h0_b=h0b, c0=c0)
    st = None
    for t in range(len(bs)):
        st = ops.rnn_seq_tc_fwd(k, Gx, Wb, bhh, bs, h0=h0, h0_b=h0b, c0=c0, t_range=(t, t + 1), out=st)
    assert torch.equal(st["Hs"], whole["Hs"]) and torch.equal(st["Hsb"], whole["Hsb"])
    assert torch.equal(st["gates"], whole["gates"])
    bw = ops.rnn_seq_tc_bwd(k, WT, bs, whole, dHs, h0=h0, c0=c0, transposed=True)
    bst = None
    for t in reversed(range(len(bs))):
        bst = ops.rnn_seq_tc_bwd(k, WT, bs, whole, dHs, h0=h0, c0=c0, t_range=(t + 1, t), out=bst, want_bias=False,
                                 transposed=True)
    assert torch.equal(bst["dGb"], bw["dGb"]) and torch.equal(bst["dGT"], bw["dGT"])
    assert torch.equal(bst["dstate"], bw["dstate"])


@pytest.mark.parametrize("kind", ["gru", "lstm"])
@pytest.mark.parametrize("H,lengths,init", [
    (64, [5, 5, 4, 2], True),                                      # cluster of 2, one 16-row slice
    (256, sorted([9, 9, 8, 6, 5, 3] * 30, reverse=True), True),    # cluster of 8, 180 ragged rows
    (512, [20] * 200 + [13] * 56, False),                          # config-2 shape: cluster of 16, 48-row slices
    (512, [20] * 40, True),                                        # 32-row slices
])
def test_rnn_cluster_resident_forward(kind, H, lengths, init, monkeypatch):
    """The cluster-resident recurrence (h exchanged through distributed shared memory) against the
    device-wide tcgen05 kernel (same bf16 operands, fp32 state); split step ranges resume exactly."""
    from showtell_b200 import _lib, ops
    k = _lib.ST_LSTM if kind == "lstm" else _lib.ST_GRU
    G = 4 if kind == "lstm" else 3
    bs = _lib.batch_sizes(lengths)
    N, B0, T = sum(bs), bs[0], len(bs)
    g = torch.Generator().manual_seed(H + 1)
    s = 1.0 / H ** 0.5
    Gx = torch.randn(N, G * H, generator=g).to(DEV)
    Whh = ((torch.rand(G * H, H, generator=g) * 2 - 1) * s).to(DEV)
    bhh = ((torch.rand(G * H, generator=g) * 2 - 1) * s).to(DEV)
    h0 = (torch.randn(B0, H, generator=g) * 0.5).to(DEV).bfloat16().float() if init else None
    c0 = (torch.randn(B0, H, generator=g) * 0.5).to(DEV) if (init and kind == "lstm") else None
    h0b = None if h0 is None else h0.bfloat16()
    Wb, _ = ops.cast_bf16(Whh, True, False)
    lib = _lib.load()
    assert lib.st_rnn_cluster_supported(k, H)
    monkeypatch.setattr(ops, "USE_CLUSTER", False)
    ref = ops.rnn_seq_tc_fwd(k, Gx, Wb, bhh, bs, h0=h0, h0_b=h0b, c0=c0)
    monkeypatch.setattr(ops, "USE_CLUSTER", True)
    l0 = lib.st_launch_count()
    out = ops.rnn_seq_tc_fwd(k, Gx, Wb, bhh, bs, h0=h0, h0_b=h0b, c0=c0)
    assert lib.st_launch_count() - l0 == 1
    torch.cuda.synchronize()
    for key in ("Hs", "Cs", "gates", "ghn"):
        if ref[key] is not None:
            assert rel_err(out[key], ref[key]) < 2e-3, key
    assert torch.equal(out["Hsb"], out["Hs"].bfloat16())
    # split ranges: [0, 3) then [3, T) resumes from the packed state rows
    if T > 4:
        part = ops.rnn_seq_tc_fwd(k, Gx, Wb, bhh, bs, h0=h0, h0_b=h0b, c0=c0, t_range=(0, 3))
        part = ops.rnn_seq_tc_fwd(k, Gx, Wb, bhh, bs, h0=h0, h0_b=h0b, c0=c0, t_range=(3, T), out=part)
        assert torch.equal(part["Hs"], out["Hs"]) and torch.equal(part["gates"], out["gates"])


@pytest.mark.parametrize("kind", ["gru", "lstm"])
@pytest.mark.parametrize("H,EX,lengths", [
    (128, 64, [7, 7, 6, 4, 4, 2]),                                 # one 64-row tile, ragged
    (512, 512, [6] * 100 + [4] * 28),                              # config-3 shape: 2 x 64-row tiles, ring reused
    (512, 512, [3] * 300 + [2] * 212),                             # config-4 shape: 4 x 128-row tiles
    (96, 40, [5, 3, 3, 1]),                                        # k tails (H, EX not multiples of 64)
])
def test_rnn_step_with_folded_context_projection(kind, H, EX, lengths):
    """rnn_attn.py:70: the step kernel that accumulates W_ih[:, E:] embed(ctx) together with W_hh h (one
    K-concatenated tensor-core product) against {small GEMM into the pre-activations; single-step kernel}."""
    from showtell_b200 import _lib, ops
    k = _lib.ST_LSTM if kind == "lstm" else _lib.ST_GRU
    G = 4 if kind == "lstm" else 3
    bs = _lib.batch_sizes(lengths)
    N, B0 = sum(bs), bs[0]
    g = torch.Generator().manual_seed(9)
    Gx = torch.randn(N, G * H, generator=g).to(DEV)
    Whh = (torch.randn(G * H, H, generator=g) * 0.08).to(DEV)
    Wx = (torch.randn(G * H, EX, generator=g) * 0.08).to(DEV)
    bhh = (torch.randn(G * H, generator=g) * 0.1).to(DEV)
    h0 = (torch.randn(B0, H, generator=g) * 0.5).to(DEV)
    c0 = (torch.randn(B0, H, generator=g) * 0.5).to(DEV) if kind == "lstm" else None
    X = torch.randn(N, EX, generator=g).to(DEV)
    Wb, _ = ops.cast_bf16(Whh, True, False)
    Wxb, _ = ops.cast_bf16(Wx, True, False)
    Xb, _ = ops.cast_bf16(X, True, False)
    h0b, _ = ops.cast_bf16(h0, True, False)
    ref, fused, Gref = None, None, Gx.clone()
    off = 0
    for t, bt in enumerate(bs):
        ops.gemm_bf16(Xb[off:off + bt], Wxb, out=Gref[off:off + bt], beta=1.0)
        ref = ops.rnn_seq_tc_fwd(k, Gref, Wb, bhh, bs, h0=h0, h0_b=h0b, c0=c0, t_range=(t, t + 1), out=ref)
        fused = ops.rnn_step_x_tc_fwd(k, Gx, Xb, Wb, Wxb, bhh, bs, t, h0=h0, h0_b=h0b, c0=c0, out=fused)
        assert ref is not None and fused is not None
        off += bt
    torch.cuda.synchronize()
    # same bf16 operands, fp32 accumulation in a different order: agreement far inside the bf16-mode bar
    for key in ("Hs", "gates") + (("Cs",) if kind == "lstm" else ("ghn",)):
        err = float((fused[key] - ref[key]).abs().max())
        assert err < 5e-3, (key, err)
    assert float((fused["Hsb"].float() - ref["Hsb"].float()).abs().max()) < 2e-2


@pytest.mark.parametrize("kind", ["gru", "lstm"])
@pytest.mark.parametrize("H,lengths", [
    (128, [7, 7, 6, 4, 4, 2]),                                     # one 64-row tile, ragged
    (512, [6] * 100 + [4] * 28),                                   # config-3 shape: 2 x 64-row tiles
    (512, [3] * 300 + [2] * 212),                                  # config-4 shape: 4 x 128-row tiles
])
def test_rnn_step_backward_with_folded_context_gradient(kind, H, lengths, monkeypatch):
    """The reverse step that forms d embed(ctx_t) = dG_t W_ih[:, E:] from the same stream of the gate-gradient
    tile as dh_{t-1}, against {single-step BPTT kernel; small GEMM}, over a whole reverse pass."""
    from showtell_b200 import _lib, ops
    k = _lib.ST_LSTM if kind == "lstm" else _lib.ST_GRU
    G = 4 if kind == "lstm" else 3
    bs = _lib.batch_sizes(lengths)
    N, B0 = sum(bs), bs[0]
    g = torch.Generator().manual_seed(13)
    Gx = torch.randn(N, G * H, generator=g).to(DEV)
    Whh = (torch.randn(G * H, H, generator=g) * 0.08).to(DEV)
    Wx = (torch.randn(G * H, H, generator=g) * 0.08).to(DEV)
    bhh = (torch.randn(G * H, generator=g) * 0.1).to(DEV)
    h0 = (torch.randn(B0, H, generator=g) * 0.5).to(DEV)
    c0 = (torch.randn(B0, H, generator=g) * 0.5).to(DEV) if kind == "lstm" else None
    dHs = torch.randn(N, H, generator=g).to(DEV)
    Wb, WT = ops.cast_bf16(Whh, True, True)
    _, WxT = ops.cast_bf16(Wx, False, True)
    h0b, _ = ops.cast_bf16(h0, True, False)
    monkeypatch.setattr(ops, "USE_CLUSTER", False)
    saved = ops.rnn_seq_tc_fwd(k, Gx, Wb, bhh, bs, h0=h0, h0_b=h0b, c0=c0)
    ref, fused, fq = None, None, None
    dX_ref = torch.zeros(N, H, device=DEV)
    dX = torch.zeros(N, H, device=DEV)
    dXq = torch.zeros(N, H, device=DEV)
    offs = [sum(bs[:t]) for t in range(len(bs))]
    # attention-query gradients (one row per packed token) and decoder_att^T, for the `query=` variant
    A = 128
    dq = (torch.randn(N, A, generator=g) * 0.3).to(DEV)
    dqb, _ = ops.cast_bf16(dq, True, False)
    _, WdT = ops.cast_bf16((torch.randn(A, H, generator=g) * 0.08).to(DEV), False, True)       # (H, A)
    for t in reversed(range(len(bs))):
        o0, o1 = offs[t], offs[t] + bs[t]
        ref = ops.rnn_seq_tc_bwd(k, WT, bs, saved, dHs, h0=h0, c0=c0, t_range=(t + 1, t), out=ref, want_bias=False)
        ops.gemm_bf16(ref["dGb"][o0:o1], WxT, out=dX_ref[o0:o1])
        fused = ops.rnn_step_x_tc_bwd(k, WT, WxT, bs, t, saved, dHs, dX, h0=h0, c0=c0, out=fused)
        assert ref is not None and fused is not None
        # the carried gradient is the only coupling between steps: compare it step by step
        e = float((fused["dstate"] - ref["dstate"]).abs().max() / ref["dstate"].abs().max())
        assert e < 1e-3, (t, e)
        # query variant: the kernel of step t adds dq[rows of t+1] . W_dec itself; the reference adds it after step t+1
        fq = ops.rnn_step_x_tc_bwd(k, WT, WxT, bs, t, saved, dHs, dXq, h0=h0, c0=c0, out=fq, query=(WdT, dqb))
        assert fq is not None
    # full-chain reference for the query variant
    refc = None
    for t in reversed(range(len(bs))):
        o0, o1 = offs[t], offs[t] + bs[t]
        refc = ops.rnn_seq_tc_bwd(k, WT, bs, saved, dHs, h0=h0, c0=c0, t_range=(t + 1, t), out=refc, want_bias=False)
        if t > 0:
            ops.gemm_bf16(dqb[o0:o1], WdT, out=refc["dstate"][0][:bs[t]], beta=1.0)
    torch.cuda.synchronize()
    eq = float((fq["dstate"] - refc["dstate"]).abs().max() / refc["dstate"].abs().max())
    assert eq < 2e-3, eq
    assert float((fq["dGb"].float() - refc["dGb"].float()).abs().max()) <= 2e-2 * float(refc["dGb"].float().abs().max())
    torch.cuda.synchronize()
    assert float((fused["dGb"].float() - ref["dGb"].float()).abs().max()) <= 2e-2 * float(ref["dGb"].float().abs().max())
    assert fused["dGT"] is None                      # row-major gate gradients only: the GEMMs read them in place
    # two tensor-core kernels with different summation orders (the reference kernel splits K over a cluster): carried
    # gradients that differ in the last fp32 bits round some bf16 gate gradients the other way (2^-9 each)
    assert float((dX - dX_ref).abs().max() / dX_ref.abs().max()) < 2e-3


@pytest.mark.parametrize("M,N,K,topk", [(100, 50, 40, 3), (129, 257, 72, 5), (640, 10000, 512, 3), (4096, 10000, 512, 1),
                                        (37, 1000, 2048, 8), (300, 10000, 512, 5)])
def test_gemm_tf32x3_fused_topk(M, N, K, topk):
    """Vocabulary projection fused with arg-max / top-K (rnn.py:51,63,90-91): same values and indices as the
    materialised 3xTF32 product followed by torch.topk; exact ties resolve to the lower index (torch.max's rule)."""
    from showtell_b200 import ops
    g = torch.Generator().manual_seed(M + N + K + topk)
    A = torch.randn(M, K, generator=g).to(DEV)
    B = (torch.randn(N, K, generator=g) / K ** 0.5).to(DEV)
    B[N // 2] = B[N // 3]                                   # two identical columns: an exact tie in every row
    bias = torch.randn(N, generator=g).to(DEV)
    bias[N // 2] = bias[N // 3]
    sa, sb = ops.split_tf32(A), ops.split_tf32(B)
    full = ops.gemm_tf32x3(sa, sb, bias=bias)
    val, idx, tok, rmax, rsum = ops.gemm_tf32x3_topk(sa, sb, topk, bias=bias, want_tokens=True, want_stats=True)
    rv, ri = torch.sort(full, dim=1, descending=True, stable=True)   # stable: lower index first among equal values
    assert torch.equal(val, rv[:, :topk]), float((val - rv[:, :topk]).abs().max())
    assert torch.equal(idx.long(), ri[:, :topk])
    assert torch.equal(tok, ri[:, 0])
    assert torch.equal(rmax, full.max(1)[0])                        # soft-max normaliser of the same logits
    lse = rmax.double() + rsum.double().log()
    assert float((lse - torch.logsumexp(full.double(), 1)).abs().max()) < 1e-5


@pytest.mark.parametrize("M,V,H,topk,cluster", [(64, 1000, 64, 3, 0), (300, 10000, 512, 5, 0), (4096, 10000, 512, 3, 0),
                                                (257, 10000, 512, 1, 0), (128, 10000, 512, 3, 12), (100, 300, 128, 8, 20)])
def test_vocab_topk_screen_is_exact(M, V, H, topk, cluster):
    """bf16 screening + fp32 re-scoring (st_vocab_topk_screen) returns the exact top-K of the fp32 logits: compared with
    float64 logits wherever the decisive margins exceed fp32 rounding noise.  `cluster` columns of ONE 128-column part
    are near-copies of the row's best column (relative perturbation 1e-4, far inside the bf16 error band): more
    survivors than a part keeps, which forces the re-score-the-whole-part path."""
    from showtell_b200 import ops
    g = torch.Generator().manual_seed(M + V + H + topk)
    h = torch.tanh(torch.randn(M, H, generator=g)).to(DEV)
    W = ((torch.rand(V, H, generator=g) * 2 - 1) / H ** 0.5).to(DEV)
    b = ((torch.rand(V, generator=g) * 2 - 1) / H ** 0.5).to(DEV)
    if cluster:
        base = int(torch.argmax((h[:1].double() @ W.double().t() + b.double())[0]))
        lo = (base // 128) * 128
        cols = [c for c in range(lo, min(V, lo + 128)) if c != base][:cluster]
        for j, c in enumerate(cols):
            W[c] = W[base] * (1.0 + 1e-4 * (j + 1) * (-1) ** j)
            b[c] = b[base]
        h[1:] = h[:1] + 1e-3 * torch.randn(M - 1, H, generator=g).to(DEV)     # every row sees the same crowd at the top
    val, idx, tok = ops.vocab_topk_screen(h, W, b, topk, want_tokens=True)
    ref = h.double() @ W.double().t() + b.double()
    rv, ri = torch.sort(ref, dim=1, descending=True, stable=True)
    # rows whose first topk + 1 reference values are separated by more than fp32 noise must match index for index
    noise = 2e-6 * ref.abs().max()
    sep = ((rv[:, :topk] - rv[:, 1:topk + 1]) > noise).all(dim=1)
    assert int(sep.sum()) >= 0.8 * M, int(sep.sum())
    assert torch.equal(idx.long()[sep], ri[:, :topk][sep])
    assert torch.equal(tok[sep], ri[:, 0][sep])
    assert float((val.double() - torch.gather(ref, 1, idx.long())).abs().max()) < noise   # the values are the logits of the returned columns
    # and on every row the returned values are the K largest up to that noise
    assert float((val.double() - rv[:, :topk]).abs().max()) < 2 * noise
