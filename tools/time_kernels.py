"""Dev micro-timing of the per-step kernels of the attention loop (CUDA events, 20 back-to-back calls
like the training loop, so L2 residency matches what the step sees)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from showtell_b200 import ops

dev = torch.device("cuda:0")
F32, BF16 = torch.float32, torch.bfloat16


def timeit(fn, n=20, reps=5):
    """n calls captured into one CUDA graph (host launch overhead excluded, as in the training step)."""
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        with torch.cuda.graph(g, stream=st):
            for _ in range(n):
                fn()
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / n)
    return best * 1e3


def attn(B, P, dt, A=512, E=512):
    att1 = torch.randn(B * P, A, device=dev).to(dt)
    Fe = torch.randn(B * P, E, device=dev).to(dt)
    att2 = torch.randn(B, A, device=dev)
    wf = torch.randn(A, device=dev) * 0.05
    bf = torch.zeros(1, device=dev)
    be = torch.zeros(E, device=dev)
    alphas = torch.zeros(B, 1, P, device=dev)
    S = torch.zeros(B, P, device=dev)
    ctx = torch.empty(B, E, device=dev)
    ctxb = torch.empty(B, E, device=dev, dtype=BF16)
    dctx = torch.randn(B, E, device=dev)
    de = torch.empty(B, P, device=dev)
    datt2 = torch.empty(B, A, device=dev)
    datt2b = torch.empty(B, A, device=dev, dtype=BF16)
    f = lambda: ops.attn_step_fwd(B, P, att1, Fe, att2, wf, bf, be, alphas[:, 0, :], P, S, ctx, ctx_bf16=ctxb, tag=None)
    b = lambda: ops.attn_step_bwd(B, P, att1, Fe, att2, wf, alphas[:, 0, :], P, None, P, dctx, de, datt2,
                                  datt2_bf16=datt2b, tag=None)
    byt = B * P * (A + E) * att1.element_size()
    tf, tb = timeit(f), timeit(b)
    print(f"attn_step B={B} P={P} {dt}: fwd {tf:.1f} us ({byt/tf/1e3:.0f} GB/s)  bwd {tb:.1f} us ({byt/tb/1e3:.0f} GB/s)",
          flush=True)


def gemm(M, N, K, beta=0.0):
    A = torch.randn(M, K, device=dev).to(BF16)
    B = torch.randn(N, K, device=dev).to(BF16)
    out = torch.zeros(M, N, device=dev)
    t = timeit(lambda: ops.gemm_bf16(A, B, out=out, beta=beta))
    print(f"gemm_bf16 {M}x{N}x{K} beta={beta}: {t:.1f} us ({2*M*N*K/t/1e6:.1f} TFLOP/s)", flush=True)


def rnn(kind, B, T=20, H=512):
    from showtell_b200 import _lib
    k = _lib.ST_LSTM if kind == "lstm" else _lib.ST_GRU
    G = 4 if kind == "lstm" else 3
    bs = [B] * T
    N = B * T
    Gx = torch.randn(N, G * H, device=dev)
    Whh = torch.randn(G * H, H, device=dev) * 0.04
    bhh = torch.zeros(G * H, device=dev)
    Wb, WT = ops.cast_bf16(Whh, True, True)
    dHs = torch.randn(N, H, device=dev)
    res = {}
    for name, flag in (("device-wide tc", False), ("cluster", True)):
        ops.USE_CLUSTER = flag
        st = {}
        def f():
            st["o"] = ops.rnn_seq_tc_fwd(k, Gx, Wb, bhh, bs, out=st.get("o"))
        tf = timeit(f, n=4)
        bst = {}
        def b():
            bst["o"] = ops.rnn_seq_tc_bwd(k, WT, bs, st["o"], dHs, out=bst.get("o"))
        tb = timeit(b, n=4)
        print(f"rnn {kind} B={B} T={T} H={H} [{name}]: fwd {tf:.1f} us ({tf/T:.2f} us/step)  bwd {tb:.1f} us ({tb/T:.2f} us/step)",
              flush=True)
    ops.USE_CLUSTER = True


def gemm_variants():
    """The GEMM shapes of the training steps under each kernel variant (0 = library's choice)."""
    from showtell_b200 import _lib
    lib = _lib.load()
    shapes = [(5120, 2048, 512, "ih_fwd cfg2"), (5120, 10000, 512, "vocab store"), (10000, 512, 5120, "vocab_dw"),
              (5120, 512, 10000, "vocab_dx"), (2048, 512, 5120, "hh_dw"), (5120, 512, 2048, "ih_dx"),
              (25088, 512, 2048, "att1 cfg3"), (512, 2048, 25088, "att1_dw cfg3"), (100352, 512, 2048, "att1 cfg4"),
              (10240, 10000, 512, "vocab store cfg4"), (512, 2048, 512, "step cfg4")]
    for M, N, K, name in shapes:
        A = torch.randn(M, K, device=dev).bfloat16()
        B = torch.randn(N, K, device=dev).bfloat16()
        out = torch.empty(M, N, device=dev)
        outb = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        row = []
        for v in (0, 128, 256, 256 | 0x1000, 2 | 0x1000):
            lib.st_debug_gemm_variant(v)
            t = timeit(lambda: ops.gemm_bf16(A, B, out=out), n=4)
            tb = timeit(lambda: ops.gemm_bf16(A, B, out=outb, out_dtype=torch.bfloat16), n=4)
            row.append(f"v{v & 0xfff}{'-nosk' if v & 0x1000 else ''}: {t:.1f} us {2*M*N*K/t/1e6:.0f} TF/s (bf16 out {tb:.1f} us {2*M*N*K/tb/1e6:.0f})")
        lib.st_debug_gemm_variant(0)
        print(f"gemm {name} {M}x{N}x{K}: " + " | ".join(row), flush=True)
    # fused cross-entropy epilogues
    for M, V, H in ((5120, 10000, 512), (2560, 10000, 512), (10240, 10000, 512)):
        Hs = (torch.randn(M, H, device=dev) * 0.5).bfloat16()
        Wv = (torch.randn(V, H, device=dev) * 0.1).bfloat16()
        bv = torch.randn(V, device=dev) * 0.1
        tgt = torch.randint(0, V, (M,), device=dev)
        row = []
        for v in (0, 128, 256, 2):
            lib.st_debug_gemm_variant(v)
            loss, lse = ops.vocab_ce_fwd(Hs, Wv, bv, tgt)
            t1 = timeit(lambda: ops.vocab_ce_fwd(Hs, Wv, bv, tgt), n=4)
            t2 = timeit(lambda: ops.vocab_ce_bwd(Hs, Wv, bv, tgt, lse, 1.0 / M), n=4)
            row.append(f"v{v}: fwd {t1:.1f} us {2*M*V*H/t1/1e6:.0f} TF/s, dlogits {t2:.1f} us {2*M*V*H/t2/1e6:.0f} TF/s")
        lib.st_debug_gemm_variant(0)
        print(f"vocab_ce {M}x{V}x{H}: " + " | ".join(row), flush=True)


if __name__ == "__main__":
    which = sys.argv[1:] or ["attn", "gemm"]
    if "attn" in which:
        attn(128, 196, BF16); attn(128, 49, BF16); attn(512, 196, BF16); attn(128, 196, F32)
    if "gemmv" in which:
        gemm_variants()
    if "rnn" in which:
        rnn("lstm", 256); rnn("gru", 128); rnn("lstm", 512); rnn("gru", 32)
    if "gemm" in which:
        for M in (128, 512):
            gemm(M, 512, 512); gemm(M, 1536, 512, 1.0); gemm(M, 512, 1536); gemm(M, 2048, 512); gemm(M, 512, 2048)
        gemm(5120, 2048, 512); gemm(25088, 512, 2048); gemm(5120, 10000, 512)
