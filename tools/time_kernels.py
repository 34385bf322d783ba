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


if __name__ == "__main__":
    which = sys.argv[1:] or ["attn", "gemm"]
    if "attn" in which:
        attn(128, 196, BF16); attn(128, 49, BF16); attn(512, 196, BF16); attn(128, 196, F32)
    if "gemm" in which:
        for M in (128, 512):
            gemm(M, 512, 512); gemm(M, 1536, 512, 1.0); gemm(M, 512, 1536); gemm(M, 2048, 512); gemm(M, 512, 2048)
        gemm(5120, 2048, 512); gemm(25088, 512, 2048); gemm(5120, 10000, 512)
