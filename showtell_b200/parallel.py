"""Batch-sharded data parallelism for the decoder training step (one process per GPU).

The reference is single-process (no DataParallel / torch.distributed anywhere);
samples are independent, so the path shards by batch rows with one exchange per iteration: a sum
all-reduce of the fp32 gradients over NCCL (NVLink 5 / NVSwitch).  The loss is a mean over the
GLOBAL token count (and the attention penalty over the global batch), so every rank scales its
local sums by the global denominators before the reduction and the reduced gradients equal the
single-GPU gradients of the concatenated batch.

Overlap: forward_loss produces gradients in backward order -- the vocabulary projection first,
then the recurrent / attention weights, the embedding last.  `GradReducer.reduce(...)` is called
by the engine as each group becomes final; each call flattens the group into one bucket and issues
the all-reduce on a side stream, so the 20 MB vocabulary bucket travels while the BPTT kernels run.
`finish()` makes the compute stream wait for the side stream.
"""
import torch
import torch.distributed as dist


def shard_rows(n_rows, rank, world):
    """Contiguous row range of `rank` (images / batch rows are independent; keeps the
    length-descending order inside each shard)."""
    base, rem = divmod(n_rows, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def global_counts(local_tokens, local_batch, group=None, device=None):
    """(global token count, global batch) -- the denominators of the mean losses."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return int(local_tokens), int(local_batch)
    t = torch.tensor([float(local_tokens), float(local_batch)], dtype=torch.float64, device=device)
    dist.all_reduce(t, group=group)
    return int(round(float(t[0]))), int(round(float(t[1])))


class GradReducer:
    """Sum all-reduce of gradient groups as flat fp32 buckets, on a side stream when on CUDA."""

    def __init__(self, group=None):
        self.group = group
        self.world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
        self._stream = None
        self._pending = []

    def _side_stream(self, device):
        if self._stream is None:
            self._stream = torch.cuda.Stream(device=device)
        return self._stream

    def reduce(self, tensors, ready=None):
        """tensors: list of gradient tensors that are final.  Reduced in place (asynchronously on
        CUDA: call finish() before reading them).  `ready`: event after which the tensors are final
        when their producers ran on another stream (ops.fork); default: the current stream's tail."""
        tensors = [t for t in tensors if t is not None]
        if self.world == 1 or not tensors:
            return
        if tensors[0].is_cuda:
            side = self._side_stream(tensors[0].device)
            if ready is None:
                ready = torch.cuda.Event()
                ready.record()                             # producers ran on the current stream
            with torch.cuda.stream(side):
                side.wait_event(ready)
                flat = torch.cat([t.reshape(-1) for t in tensors])
                dist.all_reduce(flat, group=self.group)
                off = 0
                for t in tensors:
                    t.copy_(flat[off:off + t.numel()].view_as(t))
                    off += t.numel()
                    t.record_stream(side)
                done = torch.cuda.Event()
                done.record()
            self._pending.append(done)
        else:
            flat = torch.cat([t.reshape(-1) for t in tensors])
            dist.all_reduce(flat, group=self.group)
            off = 0
            for t in tensors:
                t.copy_(flat[off:off + t.numel()].view_as(t))
                off += t.numel()

    def finish(self):
        for ev in self._pending:
            torch.cuda.current_stream().wait_event(ev)
        self._pending = []
