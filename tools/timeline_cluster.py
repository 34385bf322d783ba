import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from showtell_b200 import _lib, ops
lib = _lib.load()
lib.st_debug_set_timeline.argtypes = [ctypes.c_void_p]
dev = "cuda:0"
H, B, T = 512, int(sys.argv[1]) if len(sys.argv) > 1 else 256, 20
k, G = _lib.ST_LSTM, 4
bs = [B] * T
N = B * T
Gx = torch.randn(N, G * H, device=dev)
Whh = (torch.randn(G * H, H, device=dev) * 0.04)
bhh = torch.zeros(G * H, device=dev)
Wb, WT = ops.cast_bf16(Whh, True, True)
SAVE = 'nosave' not in sys.argv
for _ in range(2):
    out = ops.rnn_seq_tc_fwd(k, Gx, Wb, bhh, bs, save=SAVE)
tl = torch.zeros(T * 16, dtype=torch.int64, device=dev)
lib.st_debug_set_timeline(ctypes.c_void_p(tl.data_ptr()))
out = ops.rnn_seq_tc_fwd(k, Gx, Wb, bhh, bs, save=SAVE)
torch.cuda.synchronize()
lib.st_debug_set_timeline(None)
a = tl.cpu().view(T, 16)
names = "slots: 0 mma:wait_h  1 h_ready  2 mma_issued | 4 epi:acc_seen  5 tmem_read  6 stage1_done  8 stage2_done  10 fbar_ok  11 copies_issued"
print(names)
for t in range(1, T - 1):
    base = int(a[t, 0])
    print(t, [int(x) - base for x in a[t, :13]], " step:", int(a[t + 1, 1]) - int(a[t, 1]))
