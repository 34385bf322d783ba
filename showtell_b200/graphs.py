"""CUDA-graph replay of a whole training step (forward + loss + backward kernels).

A step is 20-230 kernel launches issued from Python; at the benchmark shapes the GPU work (~1 ms for
the LSTM step) is no longer than the host time to issue it.  After a step signature -- module,
parameter storage, shapes, the tuple of batch_sizes -- has been seen twice eagerly, the third call
captures the identical sequence of C-ABI calls on a capture stream into a `torch.cuda.CUDAGraph`
and later calls replay it: inputs are copied into static buffers, one graph launch runs every
kernel (cooperative launches and TMA descriptors included: all buffers come from the graph's
private pool, so the raw pointers baked into kernel parameters stay valid), outputs are the
graph's static loss / gradient tensors.

A gradient reducer (parallel.GradReducer: NCCL all-reduces on a side stream) is captured with the
step -- the side stream forks from and joins the capture stream through events, NCCL collectives
are graph-capturable -- so data-parallel ranks replay {kernels + all-reduces} as one graph each.
Not used while `ops.TIMER` records per-kernel events or when `module.use_cuda_graphs` is False.

One loss per module may be outstanding: the gradients a replayed step (or a data-parallel step, whose
reduced gradients are views of the symmetric bucket) hands to autograd live in STATIC storage that the
next forward_loss of the same module overwrites.  Every forward_loss takes a ticket (`ticket(mod)`), and
backward() checks it (`check_ticket`): using a loss whose gradients have since been overwritten -- gradient
accumulation over two losses, a validation forward_loss between forward and backward -- raises instead of
silently returning the newer step's gradients.  ("One loss at a time".)
"""
import torch

from . import ops

_EAGER_CALLS_BEFORE_CAPTURE = 2


class _Entry:
    __slots__ = ("seen", "graph", "static_in", "out", "failed")

    def __init__(self):
        self.seen, self.graph, self.static_in, self.out, self.failed = 0, None, None, None, False


def ticket(mod):
    """Called by forward_loss after its gradients exist: invalidates the tickets of earlier losses.  Returns
    (generation, static): static = the gradients live in storage the next step reuses (graph replay, or the
    symmetric buckets of a gradient reducer)."""
    static = bool(mod.__dict__.get("_last_run_static", False)) or getattr(mod, "grad_reducer", None) is not None
    gen = mod.__dict__.get("_loss_generation", 0) + (1 if static else 0)     # eager single-GPU steps own fresh tensors
    mod.__dict__["_loss_generation"] = gen
    return gen, static


def check_ticket(mod, tk):
    """Called by backward(): raises if a later forward_loss of `mod` has overwritten the static gradients."""
    gen, static = tk
    if static and mod.__dict__.get("_loss_generation", 0) != gen:
        raise RuntimeError(
            "showtell_b200: backward() of a forward_loss result whose gradients were overwritten by a later "
            "forward_loss call on the same module (the fused step keeps ONE set of static gradients per module: "
            "call backward() before the next forward_loss, or use forward() + nn.CrossEntropyLoss for "
            "several outstanding losses)")


def _copy_prologue(inputs, out):
    """Default prologue: the operands are the inputs themselves (copied into the static operands on replay)."""
    if out is None:
        return list(inputs)
    for dst, src in zip(out, inputs):
        dst.copy_(src, non_blocking=True)
    return out


def enabled(mod):
    return getattr(mod, "use_cuda_graphs", True) and ops.TIMER is None


def run(mod, key, body, inputs, prologue=None):
    """body(*inputs) -> pytree of tensors (dict / tuple / tensor / None).  Runs it eagerly, or via a
    captured graph once `key` has been seen often enough.
    prologue(inputs, out) -> operands: an eager pass in front of the graph that turns the caller's inputs into the
    body's operands -- body(*operands) -- writing them into `out` (the graph's static operands) when it is given.
    A step whose first kernel re-lays a large input (the attention decoders' 200 MB grid) thereby reads the
    caller's tensor once, instead of a device copy into a static input followed by the re-layout of that copy."""
    mod.__dict__["_last_run_static"] = False
    if prologue is None:
        prologue = _copy_prologue
    if not enabled(mod):
        return body(*prologue(inputs, None))
    cache = mod.__dict__.setdefault("_step_graphs", {})
    e = cache.get(key)
    if e is None:
        if len(cache) > 16:                      # varying shapes: do not hoard graph pools
            cache.clear()
        e = cache[key] = _Entry()
    if e.graph is not None:
        ops.refresh_shadows()                     # bf16 weight shadows live outside the graph (ops.bf16_shadow)
        prologue(inputs, e.static_in)
        e.graph.replay()
        mod.__dict__["_last_run_static"] = True
        return e.out
    e.seen += 1
    if e.failed or e.seen <= _EAGER_CALLS_BEFORE_CAPTURE:
        return body(*prologue(inputs, None))
    try:
        static_in = [t.clone() if t.data_ptr() in {i.data_ptr() for i in inputs} else t for t in prologue(inputs, None)]
        ops.refresh_shadows()                     # fresh before capture: no cast becomes part of the graph
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, capture_error_mode="relaxed"):
            out = body(*static_in)
        e.graph, e.static_in, e.out = graph, static_in, out
        graph.replay()                            # results of THIS call
        mod.__dict__["_last_run_static"] = True
        return e.out
    except Exception as exc:                      # capture unsupported for this sequence: stay eager
        e.failed = True
        import warnings
        warnings.warn(f"showtell_b200: CUDA-graph capture failed ({exc}); continuing eagerly")
        torch.cuda.synchronize()
        return body(*prologue(inputs, None))
