"""Dev tool: one batched beam-search call (for ncu).   python tools/run_beam_once.py [K] [n_img]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

K = int(sys.argv[1]) if len(sys.argv) > 1 else 3
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
from showtell_b200.rnn import RNN

dev = torch.device("cuda:0")
torch.manual_seed(1)
m = RNN(512, 512, 10000, 1).to(dev)
m.decode_gemm = "tf32x3"
feat = torch.randn(n, 512, device=dev)
for _ in range(2):
    out = m.sentence_index(feat, beam_size=K, max_len=20)
torch.cuda.synchronize()
print("ok", tuple(out.shape) if hasattr(out, "shape") else len(out))
