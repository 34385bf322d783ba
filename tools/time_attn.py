"""Dev timing of the attention decoder training step (CUDA events + torch profiler)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from showtell_b200.rnn_attn import RNN_Attn as GRU
from showtell_b200.rnn_attn_LSTM import RNN_Attn as LSTM

def run(kind, B, P, dtype, T=20, iters=5, prof=False):
    dev = torch.device("cuda:0")
    torch.manual_seed(1)
    m = (GRU if kind == "gru" else LSTM)(512, 2048, 512, 512, 10000, 1, dtype=dtype).to(dev)
    feat = torch.relu(torch.randn(B, 2048, P, device=dev))
    cap = torch.randint(4, 10000, (B, T), device=dev)
    lengths = [T] * B
    def step():
        m.zero_grad(); loss, _ = m.forward_loss(feat, cap, lengths); loss.backward(); return loss
    for _ in range(5): step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): loss = step()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    print(f"attn-{kind} {dtype} B={B} P={P}: {ms:.3f} ms/iter  {B*T/ms*1e3:.0f} tok/s  loss={float(loss.detach()):.4f}", flush=True)
    if prof:
        from torch.profiler import profile, ProfilerActivity
        with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as p:
            step(); torch.cuda.synchronize()
        print(p.key_averages().table(sort_by="cuda_time_total", row_limit=22, max_name_column_width=60))

if __name__ == "__main__":
    if len(sys.argv) > 1:      # e.g.  time_attn.py gru 128 196 bf16 [prof]
        run(sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), sys.argv[4], prof="prof" in sys.argv[5:])
    else:
        run("gru", 128, 196, "fp32")
        run("gru", 128, 196, "bf16", prof=True)
        run("gru", 128, 49, "bf16")
        run("lstm", 512, 196, "bf16")
