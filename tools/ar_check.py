"""torchrun tool: correctness + timing of the library's symmetric-memory all-reduce kernel
(csrc/allreduce.cu) against torch.distributed's NCCL all-reduce.

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/ar_check.py
"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from showtell_b200 import parallel  # noqa: E402


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    dist.init_process_group("nccl", device_id=dev)
    rank, world = dist.get_rank(), dist.get_world_size()
    force_p2p = os.environ.get("AR_FORCE_P2P", "0") == "1"
    pair = os.environ.get("AR_PAIR_P2P", "1") == "1"            # world 2: peer loads / stores (default) or multicast
    from showtell_b200 import _lib
    _lib.load().st_debug_allreduce_pair_p2p(int(pair))
    for numel in (2048 + 4, 5_120_000 + 10_000, 7_225_344):
        shapes = [(numel,)]
        b = parallel._SymBucket(shapes, None, dev)
        if force_p2p:
            b.multicast = None
        g = torch.Generator(device=dev).manual_seed(100 + rank)
        x = torch.randn(numel, device=dev, generator=g)
        ref = x.clone()
        dist.all_reduce(ref)
        for nblocks in (8, 16, 24, 32, 48, 64):
            b.views[0].copy_(x)
            torch.cuda.synchronize()
            dist.barrier()
            b.allreduce(nblocks)
            torch.cuda.synchronize()
            out = b.views[0].clone()
            err = float((out - ref).abs().max() / ref.abs().max())
            gathered = [torch.empty_like(out) for _ in range(world)]
            dist.all_gather(gathered, out)
            same = all(torch.equal(gathered[0], t) for t in gathered)
            assert err < 1e-5 and same, (numel, nblocks, err, same)
            # timing: back-to-back launches (result values grow, irrelevant)
            b.views[0].copy_(x * 1e-3)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            dist.barrier()
            torch.cuda.synchronize()
            e0.record()
            for _ in range(20):
                b.allreduce(nblocks)
            e1.record()
            torch.cuda.synchronize()
            t_sym = e0.elapsed_time(e1) / 20
            if rank == 0:
                print(f"numel {numel:9d} ({numel * 4 / 1e6:6.2f} MB) nblocks {nblocks:2d} "
                      f"{'pair-p2p' if (world == 2 and pair) else 'multicast' if b.multicast else 'p2p':9s} err {err:.1e} identical {same} "
                      f"{t_sym * 1e3:8.1f} us  busbw {2 * (world - 1) / world * numel * 4 / t_sym / 1e6:7.1f} GB/s",
                      flush=True)
        y = x * 1e-3
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(3):
            dist.all_reduce(y)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(20):
            dist.all_reduce(y)
        e1.record()
        torch.cuda.synchronize()
        t_nccl = e0.elapsed_time(e1) / 20
        # CUDA-graph capture + replay of the kernel
        b.views[0].copy_(x)
        torch.cuda.synchronize()
        dist.barrier()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            b.allreduce(24)
        b.views[0].copy_(x)
        torch.cuda.synchronize()
        dist.barrier()
        gr.replay()
        torch.cuda.synchronize()
        gerr = float((b.views[0] - ref).abs().max() / ref.abs().max())
        assert gerr < 1e-5, gerr
        if rank == 0:
            print(f"numel {numel:9d} NCCL all_reduce {t_nccl * 1e3:8.1f} us; graph replay err {gerr:.1e}", flush=True)
        del gr
    dist.barrier()
    if rank == 0:
        print("ar_check OK", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
