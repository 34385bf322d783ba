"""Optimizer step of the reference training loops (main.py:97-100,152; LSTM/main_lstm.py:88-91;
Attention/main_attn.py:91-94,134): `torch.optim.SGD(params, lr=, momentum=)` and `torch.optim.Adam(params, lr=)`.

Drop-ins with the same constructor arguments, `step()` / `zero_grad()` and -- because they subclass
`torch.optim.Optimizer` and keep torch's state keys (`momentum_buffer`; `step`, `exp_avg`, `exp_avg_sq`) -- the
same `state_dict()` layout, so checkpoints written by `utils.create_checkpoint` (utils.py:125-145) load either
way.  The arithmetic is ONE multi-tensor kernel launch per step (csrc/optim.cu: st_sgd_step / st_adam_step)
instead of torch's per-operation foreach launches.  fp32 CUDA parameters only; anything else raises.
"""
import ctypes as C

import torch

from . import _lib

_MAX = 32      # ST_OPT_MAX tensors per launch


def _check(p):
    if not (p.is_cuda and p.dtype == torch.float32 and p.is_contiguous()):
        raise RuntimeError("showtell_b200.optim: parameters must be contiguous fp32 CUDA tensors (no CPU fallback)")
    g = p.grad
    if g.is_sparse:
        raise RuntimeError("showtell_b200.optim does not support sparse gradients")
    if not (g.is_cuda and g.dtype == torch.float32):
        raise RuntimeError("showtell_b200.optim: gradients must be fp32 CUDA tensors")
    return g if g.is_contiguous() else g.contiguous()


def _ptrs(tensors):
    return (C.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])


def _counts(tensors):
    return (C.c_int64 * len(tensors))(*[t.numel() for t in tensors])


def _shadow_ptrs(params):
    """bf16 shadows (ops.bf16_shadow) the update kernel rewrites in the same pass; other copies of a parameter
    (transposed, sliced) are marked stale and re-cast on their next use.  Returns (pointer table or None, finish())."""
    from . import ops
    table, fresh, any_shadow = [], [], False
    for p in params:
        direct, others = ops.shadows_of(p)
        for sh in others:
            sh.version = None
        table.append(direct.buf.data_ptr() if direct is not None else None)
        if direct is not None:
            fresh.append((direct, p))
            any_shadow = True

    def finish():
        for sh, p in fresh:
            sh.version = p._version           # the raw-pointer update does not bump the counter: the shadow is current
    return ((C.c_void_p * len(params))(*table) if any_shadow else None), finish


class SGD(torch.optim.Optimizer):
    """torch.optim.SGD(params, lr, momentum) as main.py:98 constructs it (dampening 0, no weight decay / nesterov)."""

    def __init__(self, params, lr=1e-3, momentum=0.0):
        if lr < 0.0:
            raise ValueError(f"Invalid learning rate: {lr}")
        if momentum < 0.0:
            raise ValueError(f"Invalid momentum value: {momentum}")
        # the same group keys as torch.optim.SGD, so param_groups of a checkpoint are interchangeable
        super().__init__(params, dict(lr=lr, momentum=momentum, dampening=0, weight_decay=0, nesterov=False,
                                      maximize=False, foreach=None, differentiable=False, fused=None))

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = _lib.load()
        for group in self.param_groups:
            if group["dampening"] != 0 or group["weight_decay"] != 0 or group["nesterov"] or group["maximize"]:
                raise ValueError("showtell_b200.optim.SGD implements the reference configuration only "
                                 "(dampening 0, weight_decay 0, no nesterov, no maximize)")
            lr, mom = float(group["lr"]), float(group["momentum"])
            fresh, warm = [], []
            for p in group["params"]:
                if p.grad is None:
                    continue
                g = _check(p)
                if mom == 0.0:                                  # torch keeps no state without momentum
                    warm.append((p, g, None))
                    continue
                st = self.state[p]
                if st.get("momentum_buffer") is None:
                    st["momentum_buffer"] = torch.empty_like(p, memory_format=torch.contiguous_format)
                    fresh.append((p, g, st["momentum_buffer"]))
                else:
                    warm.append((p, g, st["momentum_buffer"]))
            for items, first in ((fresh, 1), (warm, 0)):
                for i in range(0, len(items), _MAX):
                    part = items[i:i + _MAX]
                    ps, gs = [a for a, _, _ in part], [b for _, b, _ in part]
                    ms = _ptrs([c for _, _, c in part]) if mom != 0.0 else None
                    sh, done = _shadow_ptrs(ps)
                    _lib.check(lib.st_sgd_step(len(part), _ptrs(ps), _ptrs(gs), ms, sh, _counts(ps), lr, mom, first, None,
                                               _lib.stream_ptr()), "st_sgd_step")
                    done()
        return loss


class Adam(torch.optim.Optimizer):
    """torch.optim.Adam(params, lr) as main.py:100 constructs it (betas (0.9, 0.999), eps 1e-8, no weight decay /
    amsgrad)."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8):
        if lr < 0.0:
            raise ValueError(f"Invalid learning rate: {lr}")
        if not (0.0 <= betas[0] < 1.0 and 0.0 <= betas[1] < 1.0):
            raise ValueError(f"Invalid beta parameters: {betas}")
        if eps < 0.0:
            raise ValueError(f"Invalid epsilon value: {eps}")
        super().__init__(params, dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=0, amsgrad=False, maximize=False,
                                      foreach=None, capturable=False, differentiable=False, fused=None,
                                      decoupled_weight_decay=False))

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = _lib.load()
        for group in self.param_groups:
            if group["weight_decay"] != 0 or group["amsgrad"] or group["maximize"]:
                raise ValueError("showtell_b200.optim.Adam implements the reference configuration only "
                                 "(weight_decay 0, no amsgrad, no maximize)")
            lr, (b1, b2), eps = float(group["lr"]), group["betas"], float(group["eps"])
            by_step = {}
            for p in group["params"]:
                if p.grad is None:
                    continue
                g = _check(p)
                st = self.state[p]
                if len(st) == 0:
                    st["step"] = torch.tensor(0.0, dtype=torch.float32)          # host scalar, as torch keeps it
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                st["step"] += 1
                by_step.setdefault(int(st["step"]), []).append((p, g, st["exp_avg"], st["exp_avg_sq"]))
            for t, items in by_step.items():
                for i in range(0, len(items), _MAX):
                    part = items[i:i + _MAX]
                    ps = [a for a, _, _, _ in part]
                    sh, done = _shadow_ptrs(ps)
                    _lib.check(lib.st_adam_step(len(part), _ptrs(ps), _ptrs([b for _, b, _, _ in part]),
                                                _ptrs([c for _, _, c, _ in part]), _ptrs([d for _, _, _, d in part]),
                                                sh, _counts(ps), lr, float(b1), float(b2), eps, t, None,
                                                _lib.stream_ptr()), "st_adam_step")
                    done()
        return loss
