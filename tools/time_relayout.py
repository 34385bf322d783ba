import torch, sys
sys.path.insert(0, '/root/repo')
from showtell_b200 import ops, _lib
lib = _lib.load()
for dt in (torch.float32, torch.bfloat16):
    f = torch.randn(128, 2048, 196, device='cuda').to(dt)
    for legacy in (1, 0):
        lib.st_debug_relayout_legacy(legacy)
        for _ in range(3): ops.attn_relayout(f, bf16=True, want_t=False)
        big = torch.empty(512 << 20, dtype=torch.uint8, device='cuda')
        ts = []
        for _ in range(10):
            big.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); ops.attn_relayout(f, bf16=True, want_t=False); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b) * 1e3)
        ts.sort()
        print(dt, 'legacy' if legacy else 'bulk', 'median us %.1f min %.1f' % (ts[5], ts[0]))
lib.st_debug_relayout_legacy(0)
