import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from showtell_b200 import _lib, ops
lib = _lib.load()
lib.st_debug_set_timeline.argtypes = [ctypes.c_void_p]
lib.st_debug_set_bwd_ks(int(os.environ.get("ST_BWD_KS", "0")))
dev = "cuda:0"
H, B, T = 512, 256, 20
k, G = _lib.ST_LSTM, 4
bs = [B] * T
N = B * T
Gx = torch.randn(N, G * H, device=dev)
Whh = (torch.randn(G * H, H, device=dev) * 0.04)
bhh = torch.zeros(G * H, device=dev)
Wb, WT = ops.cast_bf16(Whh, True, True)
dHs = torch.randn(N, H, device=dev)
for _ in range(2):
    out = ops.rnn_seq_tc_fwd(k, Gx, Wb, bhh, bs)
    tb = ops.rnn_seq_tc_bwd(k, WT, bs, out, dHs)
tl = torch.zeros(T * 8, dtype=torch.int64, device=dev)
lib.st_debug_set_timeline(ctypes.c_void_p(tl.data_ptr()))
out = ops.rnn_seq_tc_fwd(k, Gx, Wb, bhh, bs)
torch.cuda.synchronize()
a = tl.cpu().view(T, 8)
print("FWD per step (ns rel. to step's first stamp): [wait_begin, barrier_passed, first_kblock, last_kblock, acc_ready, tmem_read, epi_done, arrived]")
for t in range(1, T):
    base = int(a[t, 0]); prev = int(a[t - 1, 7])
    print(t, [int(x) - base for x in a[t]], "since prev arrive:", base - prev, " step total:", int(a[t, 7]) - int(a[t - 1, 7]))
tl.zero_()
tb = ops.rnn_seq_tc_bwd(k, WT, bs, out, dHs)
torch.cuda.synchronize()
a = tl.cpu().view(T, 8)
print("BWD per step: [phase1_begin, phase1_done, barrier_passed, first_kblock, last_kblock, acc_ready]")
for t in range(T - 2, -1, -1):
    base = int(a[t, 0])
    print(t, [int(x) - base for x in a[t, :6]], " step total:", int(a[t, 0]) - int(a[t + 1, 0]))
lib.st_debug_set_timeline(None)
