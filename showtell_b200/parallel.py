"""Batch-sharded data parallelism for the decoder training step (one process per GPU).

The reference is single-process (no DataParallel / torch.distributed anywhere);
samples are independent, so the path shards by batch rows with one exchange per iteration: a sum
all-reduce of the fp32 gradients over NVLink 5 / NVSwitch.  The loss is a mean over the
GLOBAL token count (and the attention penalty over the global batch), so every rank scales its
local sums by the global denominators before the reduction and the reduced gradients equal the
single-GPU gradients of the concatenated batch.

Overlap: forward_loss produces gradients in backward order -- the vocabulary projection first,
then the recurrent / attention weights, the embedding last.  `GradReducer.reduce(...)` is called
by the engine as each group becomes final; each call gathers the group into one symmetric bucket
and issues the all-reduce on a side stream, so the 20 MB vocabulary bucket travels while the BPTT
kernels run.  `finish()` makes the compute stream wait for the side stream.  The exchange itself is
the library's own kernel (csrc/allreduce.cu: two-shot over NVLink peer mappings, in-switch
multimem reduction when the NVSwitch multicast mapping exists), not an NCCL call.
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


class _SymBucket:
    """One symmetric gradient bucket: `numel` fp32 slots followed by the all-reduce kernel's flag words,
    allocated identically on every rank and mapped into every peer (torch symmetric memory does the
    allocation + handle exchange; the arithmetic is st_allreduce_sum_f32, csrc/allreduce.cu)."""

    def __init__(self, shapes, group, device):
        import ctypes as C
        import torch.distributed._symmetric_memory as symm
        from . import _lib
        lib = _lib.load()
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        pad = lambda n: (n + 63) // 64 * 64            # 256-byte aligned slots
        self.offsets, off = [], 0
        for shp in shapes:
            self.offsets.append(off)
            off += pad(int(torch.Size(shp).numel()))
        self.numel, self.shapes = off, [tuple(s) for s in shapes]
        fw = lib.st_allreduce_flag_words(world)
        self.buf = symm.empty(self.numel + fw, dtype=torch.float32, device=device)
        self.buf.zero_()
        hdl = symm.rendezvous(self.buf, group if group is not None else dist.group.WORLD)
        base = [int(p) for p in hdl.buffer_ptrs]
        delta = self.buf.data_ptr() - base[rank]       # the tensor's offset inside the symmetric allocation
        self.peers = (C.c_void_p * world)(*[b + delta for b in base])
        self.flags = (C.c_void_p * world)(*[b + delta + 4 * self.numel for b in base])
        mc = int(hdl.multicast_ptr or 0)                # 0: no NVSwitch multicast on this node
        self.multicast = C.c_void_p(mc + delta) if mc else None
        self.hdl, self.rank, self.world = hdl, rank, world
        self.ones = torch.ones((), dtype=torch.float32, device=device)
        torch.cuda.synchronize(device)
        dist.barrier(group)                             # every rank's flags are zero before anyone signals
        self.views = [self.buf[o:o + int(torch.Size(s).numel())].view(s) for o, s in zip(self.offsets, self.shapes)]

    def allreduce(self, nblocks):
        from . import _lib
        _lib.check(_lib.load().st_allreduce_sum_f32(self.peers, self.multicast, self.flags, self.rank, self.world,
                                                    self.numel, nblocks, _lib.stream_ptr()), "st_allreduce_sum_f32")


class GradReducer:
    """Sum all-reduce of gradient groups on a side stream.

    CUDA: each group is gathered into a symmetric bucket (one multi-tensor copy) and reduced in place
    by the library's own two-shot NVLink / NVSwitch kernel (`backend="symm"`, default); reduce()
    returns views of the bucket, valid until the same group is reduced again.  `backend="nccl"` keeps
    torch.distributed's NCCL all-reduce (A/B reference; also used, with a warning, if the node cannot
    set up peer mappings).  CPU tensors (the gloo tests) always go through torch.distributed."""

    def __init__(self, group=None, backend=None, nblocks=None):
        import os
        self.group = group
        self.world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
        self.backend = backend or os.environ.get("SHOWTELL_ALLREDUCE", "symm")
        # CTAs of the exchange kernel: each rank moves 1/world of a bucket, so small worlds need more loads in flight
        # per rank (measured, 20.5 MB: 2 GPUs 113 us with 8 CTAs, 76 us with 16; 8 GPUs 68 us with 8, 69 us with 16)
        self.nblocks = int(nblocks or os.environ.get("SHOWTELL_AR_BLOCKS", "16" if self.world <= 4 else "8"))
        self._stream = None
        self._pending = []
        self._buckets = {}
        self._call = 0
        # development aid: exchanges (by position in the step) whose kernel is not launched -- timing breakdowns only
        self._skip = {int(x) for x in os.environ.get("SHOWTELL_AR_SKIP", "").split(",") if x.strip()}

    def _side_stream(self, device):
        if self._stream is None:
            self._stream = torch.cuda.Stream(device=device)
        return self._stream

    def slots(self, shapes):
        """Views of the symmetric bucket the NEXT reduce() call of this step will use, if that call has
        been seen before with tensors of exactly `shapes` (else None): producers that write their
        gradients straight into them skip the gather copy."""
        if self.world == 1 or self.backend != "symm":
            return None
        b = self._buckets.get((self._call, tuple(tuple(s) for s in shapes)))
        return b.views if b is not None else None

    def _bucket(self, tensors):
        key = (self._call, tuple(tuple(t.shape) for t in tensors))
        b = self._buckets.get(key)
        if b is None:
            if torch.cuda.is_current_stream_capturing():
                raise RuntimeError("GradReducer: a symmetric bucket cannot be created during CUDA-graph capture "
                                   "(run the step eagerly once first)")
            # every rank must run the same exchanges with the same number of CTAs, or the barrier epochs desynchronise
            mine = (self.nblocks, tuple(sorted(self._skip)), self._call, tuple(tuple(t.shape) for t in tensors))
            seen = [None] * self.world
            dist.all_gather_object(seen, mine, group=self.group)
            if any(x != mine for x in seen):
                raise RuntimeError(f"GradReducer: ranks disagree on the exchange setup (CTAs, skipped exchanges, position, "
                                   f"shapes): {seen}")
            try:
                b = _SymBucket([t.shape for t in tensors], self.group, tensors[0].device)
            except Exception as exc:                    # no peer access on this node: NCCL does the exchange
                import warnings
                warnings.warn(f"showtell_b200: symmetric-memory all-reduce unavailable ({exc}); using NCCL")
                self.backend, b = "nccl", None
            self._buckets[key] = b
        return b

    def reduce(self, tensors, ready=None):
        """tensors: list of gradient tensors that are final.  Returns the list of reduced tensors
        (asynchronously on CUDA: call finish() before reading them).  `ready`: event after which the
        inputs are final when their producers ran on another stream (ops.fork); default: the current
        stream's tail."""
        tensors = [t for t in tensors if t is not None]
        if self.world == 1 or not tensors or self.backend == "none":     # "none": timing floor without the exchange
            return tensors
        if not tensors[0].is_cuda:
            flat = torch.cat([t.reshape(-1) for t in tensors])
            dist.all_reduce(flat, group=self.group)
            off = 0
            for t in tensors:
                t.copy_(flat[off:off + t.numel()].view_as(t))
                off += t.numel()
            return tensors
        side = self._side_stream(tensors[0].device)
        bucket = self._bucket(tensors) if self.backend == "symm" else None
        self._call += 1
        if ready is None:
            ready = torch.cuda.Event()
            ready.record()                                 # producers ran on the current stream
        with torch.cuda.stream(side):
            side.wait_event(ready)
            if bucket is not None:
                from . import ops
                todo = [(t, v) for t, v in zip(tensors, bucket.views) if t.data_ptr() != v.data_ptr()]
                if todo:                                   # gather what was not produced in place: one launch
                    ops.scale_multi([t for t, _ in todo], bucket.ones, outs=[v for _, v in todo])
                if self._call - 1 not in self._skip:
                    bucket.allreduce(self.nblocks)
                out = bucket.views
                for t, _ in todo:
                    t.record_stream(side)
            else:
                flat = torch.cat([t.reshape(-1) for t in tensors])
                dist.all_reduce(flat, group=self.group)
                off = 0
                for t in tensors:
                    t.copy_(flat[off:off + t.numel()].view_as(t))
                    off += t.numel()
                    t.record_stream(side)
                out = tensors
            done = torch.cuda.Event()
            done.record()
        self._pending.append(done)
        return out

    def finish(self):
        for ev in self._pending:
            torch.cuda.current_stream().wait_event(ev)
        self._pending = []
        self._call = 0
        if self.backend == "symm" and self._buckets:
            from . import _lib
            err = _lib.load().st_allreduce_error()          # a mapped host word: no synchronisation
            if err:
                raise RuntimeError(f"showtell_b200: a gradient exchange timed out waiting for rank {err - 1} "
                                   "(st_allreduce_set_timeout_ms; the gradients of that step are undefined)")
