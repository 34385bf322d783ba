"""Dev tool: kernel-by-kernel timeline of ONE replayed training step (torch profiler / CUPTI):
start offset, duration and gap to the previous kernel's end on the device, plus per-kernel-name totals.
    python tools/trace_step.py attn_gru 128 196 bf16     |  python tools/trace_step.py lstm 256 0 bf16
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile


def main():
    kind, B, P, dtype = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
    T = int(sys.argv[5]) if len(sys.argv) > 5 else 20
    rank = int(os.environ.get("LOCAL_RANK", "0"))
    dev = torch.device(f"cuda:{rank}")
    torch.cuda.set_device(dev)
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(1)
    if kind.startswith("attn"):
        from showtell_b200.rnn_attn import RNN_Attn as G
        from showtell_b200.rnn_attn_LSTM import RNN_Attn as Lm
        m = (G if kind == "attn_gru" else Lm)(512, 2048, 512, 512, 10000, 1, dtype=dtype).to(dev)
        feat = torch.relu(torch.randn(B, 2048, P, device=dev))
    else:
        from showtell_b200.rnn import RNN as G
        from showtell_b200.rnn_lstm import RNN as Lm
        m = (G if kind == "gru" else Lm)(512, 512, 10000, 1, dtype=dtype).to(dev)
        feat = torch.randn(B, 512, device=dev)
    cap = torch.randint(4, 10000, (B, T), device=dev)
    lengths = [T] * B
    world = int(os.environ.get("WORLD_SIZE", "1"))
    kw = {}
    if world > 1:      # under torchrun: the data-parallel step (rank 0 prints its own timeline)
        from showtell_b200 import parallel
        m.grad_reducer = parallel.GradReducer()
        kw = {"global_tokens": B * T * world}
        if kind.startswith("attn"):
            kw["global_batch"] = B * world

    def step():
        m.forward_backward(feat, cap, lengths, **kw)

    for _ in range(6):
        step()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as p:
        step()
        torch.cuda.synchronize()
    evs = [e for e in p.events() if e.device_type == torch.autograd.DeviceType.CUDA and "memcpy" not in e.name.lower()
           and "memset" not in e.name.lower()]
    evs.sort(key=lambda e: e.time_range.start)
    t0 = evs[0].time_range.start
    end_prev = t0
    tot = {}
    lines = []
    for e in evs:
        s, d = e.time_range.start - t0, e.time_range.end - e.time_range.start
        gap = e.time_range.start - end_prev
        end_prev = max(end_prev, e.time_range.end)
        nm = e.name.replace("(anonymous namespace)::", "").replace("void ", "").split("(")[0][:70]
        lines.append(f"{s:9.1f} {d:8.1f} {gap:8.1f}  {nm}")
        a = tot.setdefault(nm, [0, 0.0, 0.0])
        a[0] += 1; a[1] += d; a[2] += max(gap, 0.0)
    span = end_prev - t0
    if rank != 0:
        os._exit(0)
    print(f"# {kind} B={B} P={P} {dtype} T={T}: {len(evs)} kernels, span {span:.1f} us")
    print("# per kernel name: count, total us, total gap-before us")
    for nm, (c, d, g) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        print(f"{c:5d} {d:9.1f} {g:9.1f}  {nm}")
    print("# timeline: start us, dur us, gap us, name")
    print("\n".join(lines))
    sys.stdout.flush()
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        os._exit(0)


if __name__ == "__main__":
    main()
