"""torchrun worker of tests/test_gpu_parallel.py (world_size >= 2, one process per GPU): the data-parallel
training step -- batch shards, global denominators, gradients exchanged by the library's symmetric-memory
all-reduce kernel on a side stream -- reproduces the single-GPU gradients of the whole batch, eagerly and
as a replayed CUDA graph.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tests/dp_worker.py
"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from showtell_b200 import parallel  # noqa: E402
from showtell_b200.rnn_attn import RNN_Attn as AttnGru  # noqa: E402
from showtell_b200.rnn_lstm import RNN as LstmRNN  # noqa: E402


def rel(a, b):
    return float((a - b).abs().max() / max(float(b.abs().max()), 1e-30))


def run_case(name, make, feat, cap, lengths, dev, tol, backend):
    rank, world = dist.get_rank(), dist.get_world_size()
    B = feat.shape[0]
    attn = feat.dim() == 3
    torch.manual_seed(5)
    full = make().to(dev)
    torch.manual_seed(5)
    part = make().to(dev)
    part.grad_reducer = parallel.GradReducer(backend=backend)
    n_glob = sum(lengths)
    lo, hi = parallel.shard_rows(B, rank, world)

    def step(m, f, c, l, **kw):
        for p in m.parameters():
            p.grad = None
        out = m.forward_loss(f, c, l, **kw)
        loss = out[0] if attn else out
        loss.backward()
        return loss.detach().clone()

    kw_full = dict(alpha_c=1.0) if attn else {}
    loss_full = step(full, feat.to(dev), cap.to(dev), lengths, **kw_full)
    ref = {n: p.grad.clone() for n, p in full.named_parameters()}
    kw = dict(alpha_c=1.0, global_tokens=n_glob, global_batch=B) if attn else dict(global_tokens=n_glob)
    f, c, l = feat[lo:hi].to(dev), cap[lo:hi].to(dev), lengths[lo:hi]
    worst = 0.0
    for it in range(5):                       # 2 eager steps, capture, 2 replays
        loss = step(part, f, c, l, **kw)
        dist.all_reduce(loss)
        assert abs(float(loss) - float(loss_full)) <= tol * abs(float(loss_full)), (name, it, float(loss), float(loss_full))
        for n, p in part.named_parameters():
            if n == "attn.full_att.bias":     # identically zero up to rounding noise (softmax shift invariance)
                continue
            e = rel(p.grad, ref[n])
            worst = max(worst, e)
            assert e <= tol, (name, backend, it, n, e)
        # every rank holds bit-identical reduced gradients
        probe = torch.stack([p.grad.double().sum() for p in part.parameters()])
        gathered = [torch.empty_like(probe) for _ in range(world)]
        dist.all_gather(gathered, probe)
        assert all(torch.equal(gathered[0], g) for g in gathered), (name, backend, it)
    used = part.grad_reducer.backend
    if rank == 0:
        print(f"dp_worker {name:10s} backend {used:5s} world {world}: loss + grads match the single-GPU step "
              f"(worst rel err {worst:.2e}, tol {tol:g})", flush=True)


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    dist.init_process_group("nccl", device_id=dev)
    g = torch.Generator().manual_seed(11)
    E, H, V, B, T = 64, 128, 503, 48, 9
    lengths = sorted([int(x) for x in torch.randint(3, T + 1, (B,), generator=g)], reverse=True)
    lengths[0] = T
    cap = torch.randint(4, V, (B, T), generator=g)
    feat = torch.randn(B, E, generator=g)
    C, A, Pn = 96, 64, 12
    featg = torch.relu(torch.randn(B, C, Pn, generator=g))
    for backend in ("symm", "nccl"):
        for dtype, tol in (("fp32", 2e-5), ("bf16", 2e-3)):
            run_case(f"lstm/{dtype}", lambda: LstmRNN(E, H, V, 2, dtype=dtype), feat, cap, lengths, dev, tol, backend)
            run_case(f"attn/{dtype}", lambda: AttnGru(E, C, A, H, V, 1, dtype=dtype), featg, cap, lengths, dev, tol, backend)
    dist.barrier()
    if dist.get_rank() == 0:
        print("dp_worker OK", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
