"""Dev timing of the persistent recurrent kernels alone (CUDA events, L2-warm): forward + BPTT of one layer."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from showtell_b200 import _lib, ops

DEV = "cuda:0"


def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


def main():
    lib = _lib.load()
    for kind, G, B in (("lstm", 4, 256), ("gru", 3, 256), ("lstm", 4, 128), ("gru", 3, 32)):
        k = _lib.ST_LSTM if kind == "lstm" else _lib.ST_GRU
        H, T = 512, 20
        bs = [B] * T
        N = B * T
        g = torch.Generator().manual_seed(1)
        Gx = torch.randn(N, G * H, generator=g).to(DEV)
        Whh = (torch.randn(G * H, H, generator=g) * 0.04).to(DEV)
        bhh = torch.zeros(G * H, device=DEV)
        dHs = (torch.randn(N, H, generator=g) * 0.01).to(DEV)
        Wb, WT = ops.cast_bf16(Whh, True, True)
        out = ops.rnn_seq_tc_fwd(k, Gx, Wb, bhh, bs)
        t_f = timeit(lambda: ops.rnn_seq_tc_fwd(k, Gx, Wb, bhh, bs, out=out))
        line = f"{kind} B={B}: fwd {t_f:.1f} us ({t_f / T:.2f}/step)"
        ref = None
        for kp in (1, 2, 4, 0):
            lib.st_debug_set_bwd_kp(kp)
            bo = ops.rnn_seq_tc_bwd(k, WT, bs, out, dHs)
            t_b = timeit(lambda: ops.rnn_seq_tc_bwd(k, WT, bs, out, dHs, out=bo))
            if ref is None:
                ref = (bo["dGb"].clone(), bo["dstate"].clone())
            same = torch.equal(ref[0], bo["dGb"]) and torch.equal(ref[1], bo["dstate"])
            line += f" | bwd kp={kp}: {t_b:.1f} us ({t_b / T:.2f}/step){'' if same else ' MISMATCH'}"
        print(line, flush=True)


if __name__ == "__main__":
    main()
