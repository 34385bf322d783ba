"""Dev timing of the tensor-core GEMM at the step's shapes (CUDA events, L2 flushed between runs).
    python tools/time_gemms.py            # default variants; SHOWTELL_GEMM_VARIANT=0x2000 -> MN-major through 2-D boxes
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from showtell_b200 import _lib, ops

dev = "cuda:0"
lib = _lib.load()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
BF = torch.bfloat16


def bench(name, fn, flops, iters=10):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(iters):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    t = sorted(ts)[len(ts) // 2]
    print(f"{name:46s} {t * 1e3:8.1f} us  {flops / t / 1e9:8.1f} TFLOP/s")


def main():
    for var in ([0, 0x2000] if len(sys.argv) < 2 else [int(sys.argv[1], 0)]):
        lib.st_debug_gemm_variant(var)
        print(f"--- st_debug_gemm_variant({var:#x})")
        BP, C, A, E, N, V, H = 128 * 196, 2048, 512, 512, 2560, 10000, 512
        F = torch.randn(BP, C, device=dev).to(BF)
        W = torch.randn(A, C, device=dev).to(BF)
        d1 = torch.randn(BP, A, device=dev).to(BF)
        bench("att1_fwd  F W^T        (25088,512,2048) K-major", lambda: ops.gemm_bf16(F, W, out_dtype=BF), 2.0 * BP * A * C)
        Hb2 = torch.randn(5120, 512, device=dev).to(BF)
        Wv2 = torch.randn(10000, 512, device=dev).to(BF)
        bench("vocab-like Hs Wv^T     (5120,10000,512) K-major", lambda: ops.gemm_bf16(Hb2, Wv2, out_dtype=BF), 2.0 * 5120 * 10000 * 512)
        bench("att1_dw   d1^T F       (512,2048,25088) MN/MN", lambda: ops.gemm_bf16(d1, F, a_t=True, b_t=True), 2.0 * BP * A * C)
        for n_tok, tag in ((2560, "cfg3"), (5120, "cfg2")):
            Pm = torch.randn(n_tok, V, device=dev).to(BF)
            Hb = torch.randn(n_tok, H, device=dev).to(BF)
            Wv = torch.randn(V, H, device=dev).to(BF)
            fl = 2.0 * n_tok * V * H
            bench(f"vocab_dw  P^T Hs  {tag}  (10000,512,{n_tok}) MN/MN", lambda: ops.gemm_bf16(Pm, Hb, a_t=True, b_t=True), fl)
            bench(f"vocab_dx  P Wv    {tag}  ({n_tok},512,10000) K/MN", lambda: ops.gemm_bf16(Pm, Wv, b_t=True), fl)
        dG = torch.randn(5120, 2048, device=dev).to(BF)
        Hp = torch.randn(5120, 512, device=dev).to(BF)
        Wih = torch.randn(2048, 512, device=dev).to(BF)
        bench("hh_dw     dG^T Hprev   (2048,512,5120) MN/MN", lambda: ops.gemm_bf16(dG, Hp, a_t=True, b_t=True), 2.0 * 5120 * 2048 * 512)
        bench("ih_dx     dG Wih       (5120,512,2048) K/MN", lambda: ops.gemm_bf16(dG, Wih, b_t=True), 2.0 * 5120 * 2048 * 512)
    lib.st_debug_gemm_variant(0)


if __name__ == "__main__":
    main()
