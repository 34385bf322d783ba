"""Dev timing of the base decoder training step (CUDA events)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from showtell_b200.rnn import RNN as GRU
from showtell_b200.rnn_lstm import RNN as LSTM

def run(kind, B, T=20, E=512, H=512, V=10000, L=1, iters=5, dtype="bf16"):
    dev = torch.device("cuda:0")
    torch.manual_seed(1)
    m = (GRU if kind == "gru" else LSTM)(E, H, V, L, dtype=dtype).to(dev)
    feat = torch.randn(B, E, device=dev)
    cap = torch.randint(4, V, (B, T), device=dev)
    lengths = [T] * B
    for _ in range(3):
        m.zero_grad(); loss = m.forward_loss(feat, cap, lengths); loss.backward()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        m.zero_grad(); loss = m.forward_loss(feat, cap, lengths); loss.backward()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    print(f"{kind} {dtype} B={B} L={L}: {ms:.3f} ms/iter  {B*T/ms*1e3:.0f} tok/s  loss={float(loss):.4f}", flush=True)

if __name__ == "__main__":
    run("gru", 32); run("lstm", 256); run("gru", 128); run("lstm", 512)
    from torch.profiler import profile, ProfilerActivity
    dev = torch.device("cuda:0")
    m = LSTM(512, 512, 10000, 1, dtype="bf16").to(dev)
    feat = torch.randn(256, 512, device=dev); cap = torch.randint(4, 10000, (256, 20), device=dev)
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(3):
            m.zero_grad(); m.forward_loss(feat, cap, [20] * 256).backward()
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=20))
