"""Dev tool: how many columns survive the screening bound of st_vocab_topk_screen, and how often a part overflows.
    python tools/screen_stats.py [scale_h]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from showtell_b200 import ops

dev = "cuda:0"
torch.manual_seed(0)
M, V, H, K = 4096, 10000, 512, 3
from showtell_b200.rnn import RNN
m = RNN(512, 512, 10000, 1).to(dev)
W, b = m.linear.weight.detach(), m.linear.bias.detach()
feat = torch.randn(M, 512, device=dev)
# hidden state after a few greedy steps of the random-init decoder (what the bench feeds the projection)
with torch.no_grad():
    gru = torch.nn.GRU(512, 512, 1).to(dev)
    gru.load_state_dict({k.replace("unit.", ""): v for k, v in m.state_dict().items() if k.startswith("unit.")})
    out, hN = gru(feat[None])
    for _ in range(4):
        tok = (out[0] @ W.t() + b).argmax(1)
        out, hN = gru(m.embeddings.weight[tok][None], hN)
h = out[0].contiguous()
print("|h| mean", float(h.norm(dim=1).mean()), "wmax", float(W.norm(dim=1).max()))
approx = (h.bfloat16() @ W.bfloat16().t()).float() + b
eps = 0.0041874 * h.norm(dim=1) * W.norm(dim=1).max()
kth = approx.topk(K, dim=1).values[:, -1]
tau = kth - 2 * eps
surv = approx >= tau[:, None]
print("survivors per row: mean", float(surv.sum(1).float().mean()), "max", int(surv.sum(1).max()))
pad = (128 - V % 128) % 128
sp = torch.nn.functional.pad(surv, (0, pad)).view(M, -1, 128).sum(2)
print("parts with >= 3 survivors: rows", int((sp >= 3).any(1).sum()), "parts", int((sp >= 3).sum()))
print("logit std", float(approx.std(dim=1).mean()), "2eps mean", float(2 * eps.mean()))
for _ in range(3):
    ops.vocab_topk_screen(h, W, b, K)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    ops.vocab_topk_screen(h, W, b, K)
e1.record(); torch.cuda.synchronize()
print("vocab_topk_screen incl. casts:", e0.elapsed_time(e1) / 10 * 1e3, "us")
