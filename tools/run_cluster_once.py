"""Dev: a few launches of the cluster-resident recurrent forward (LSTM, config-2 shape) for ncu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from showtell_b200 import _lib, ops
dev = "cuda:0"
H, B, T = 512, int(sys.argv[1]) if len(sys.argv) > 1 else 256, 20
k, G = _lib.ST_LSTM, 4
bs = [B] * T
Gx = torch.randn(B * T, G * H, device=dev)
Whh = torch.randn(G * H, H, device=dev) * 0.04
bhh = torch.zeros(G * H, device=dev)
Wb, WT = ops.cast_bf16(Whh, True, True)
o = None
for _ in range(3):
    o = ops.rnn_seq_tc_fwd(k, Gx, Wb, bhh, bs, out=o)
torch.cuda.synchronize()
print("ok", float(o["Hs"].abs().mean()))
