"""Dev tool: per-kernel totals of one batched beam-search call (torch profiler / CUPTI).
    python tools/trace_beam.py [K] [n_img] [screen]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile


def main():
    K = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
    from showtell_b200 import _lib
    from showtell_b200.rnn import RNN
    if len(sys.argv) > 3:
        _lib.load().st_debug_decode_screen(int(sys.argv[3]))
    dev = torch.device("cuda:0")
    torch.manual_seed(1)
    m = RNN(512, 512, 10000, 1).to(dev)
    m.decode_gemm = "tf32x3"
    feat = torch.randn(n, 512, device=dev)
    for _ in range(2):
        m.sentence_index(feat, beam_size=K, max_len=20)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as p:
        m.sentence_index(feat, beam_size=K, max_len=20)
        torch.cuda.synchronize()
    tot = {}
    for e in p.events():
        if e.device_type != torch.autograd.DeviceType.CUDA:
            continue
        nm = e.name.replace("(anonymous namespace)::", "").replace("void ", "").split("(")[0][:70]
        a = tot.setdefault(nm, [0, 0.0])
        a[0] += 1
        a[1] += e.time_range.end - e.time_range.start
    print(f"# beam-{K}, {n} images: per kernel name: count, total us, mean us")
    for nm, (c, d) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        print(f"{c:5d} {d:10.1f} {d / c:8.1f}  {nm}")


if __name__ == "__main__":
    main()
