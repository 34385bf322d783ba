"""Encoder head of the base captioning models (cnn.py:37-38,49): the two layers of `ResNet` that main.py:96
trains, `last_layer(linear_secondlast_layer(x))` = BatchNorm1d(embed_dim, momentum=0.01) of Linear(2048, embed_dim)
applied to the pooled, detached trunk output (cnn.py:46-47).  It is the step immediately before the decoder
path (its output is `cnn_feature`, its incoming gradient is the decoder's d cnn_feature); the ResNet trunk itself
stays a torch-side feature producer.

`EncoderHead` keeps the reference's sub-module names, so the corresponding entries of a `ResNet.state_dict()`
(`linear_secondlast_layer.*`, `last_layer.*` incl. running statistics) load unchanged; the arithmetic runs on
the library's GEMMs + fused batch-norm kernels (csrc/head.cu).  CUDA only.
"""
import torch
import torch.nn as nn

from . import _lib, ops
from .engine import weight_grad

F32 = torch.float32


def _bn_fwd(Y, gamma, beta, eps, momentum, use_running, rmean, rvar, save):
    lib = _lib.load()
    B, E = Y.shape
    out = torch.empty_like(Y)
    sm = torch.empty(E, dtype=F32, device=Y.device) if save else None
    si = torch.empty(E, dtype=F32, device=Y.device) if save else None
    _lib.check(lib.st_bn1d_fwd(_lib.ptr(Y, F32), Y.stride(0), B, E, _lib.ptr(gamma, F32), _lib.ptr(beta, F32), float(eps),
                               float(momentum), int(use_running), _lib.ptr(rmean), _lib.ptr(rvar), _lib.ptr(sm), _lib.ptr(si),
                               _lib.ptr(out, F32), out.stride(0), _lib.stream_ptr()), "st_bn1d_fwd")
    return out, sm, si


class _HeadFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mod, x, W, b, gamma, beta):
        if not x.is_cuda:
            raise RuntimeError("showtell_b200 runs on CUDA tensors only (no CPU fallback)")
        if x.dim() != 2 or x.shape[1] != W.shape[1]:
            raise ValueError(f"EncoderHead expects (B, {W.shape[1]}) pooled features, got {tuple(x.shape)}")
        x = x.detach().contiguous().to(F32)
        mode = mod.compute_dtype
        if mode == "bf16":
            xb, _ = ops.cast_bf16(x, True, False)
            Wb, _ = ops.cast_bf16(W.detach(), True, False)
            Y = ops.gemm_bf16(xb, Wb, bias=b.detach(), tag="head_fwd")
        else:
            Y = ops.sgemm(x, W.detach(), transB=True, bias=b.detach(), tag="head_fwd")          # cnn.py:37
        bn = mod.last_layer
        train = mod.training or not bn.track_running_stats
        if train and x.shape[0] < 2:
            raise ValueError("Expected more than 1 value per channel when training")           # as nn.BatchNorm1d
        out, sm, si = _bn_fwd(Y, gamma.detach(), beta.detach(), bn.eps, bn.momentum, not train,
                              bn.running_mean if bn.track_running_stats else None,
                              bn.running_var if bn.track_running_stats else None, train)       # cnn.py:38
        if train and bn.track_running_stats:
            bn.num_batches_tracked += 1
        ctx.mod, ctx.train = mod, train
        ctx.save_for_backward(x, W, gamma, Y, sm, si)
        return out

    @staticmethod
    def backward(ctx, dOut):
        if not ctx.train:
            raise NotImplementedError("EncoderHead: backward through eval-mode batch norm (the reference never does)")
        x, W, gamma, Y, sm, si = ctx.saved_tensors
        lib = _lib.load()
        B, E = Y.shape
        dOut = dOut.contiguous().to(F32)
        dY = torch.empty_like(Y)
        dgamma = torch.empty(E, dtype=F32, device=Y.device)
        dbeta = torch.empty(E, dtype=F32, device=Y.device)
        _lib.check(lib.st_bn1d_bwd(_lib.ptr(Y, F32), Y.stride(0), _lib.ptr(dOut, F32), dOut.stride(0), B, E,
                                   _lib.ptr(gamma.detach(), F32), _lib.ptr(sm, F32), _lib.ptr(si, F32), _lib.ptr(dgamma),
                                   _lib.ptr(dbeta), _lib.ptr(dY), dY.stride(0), _lib.stream_ptr()), "st_bn1d_bwd")
        mode = ctx.mod.compute_dtype
        dW = weight_grad(mode, dY, x, "head_dw")                                               # dY^T x
        db = ops.colsum(dY)
        dx = None
        if ctx.needs_input_grad[1]:
            dx = ops.sgemm(dY, W.detach(), tag="head_dx")
        return None, dx, dW, db, dgamma, dbeta


class EncoderHead(nn.Module):
    """`ResNet.linear_secondlast_layer` + `ResNet.last_layer` (cnn.py:37-42): same names, shapes and initialisation."""

    def __init__(self, in_features=2048, embed_dim=256, *, dtype="fp32"):
        super().__init__()
        if dtype not in ("fp32", "bf16"):
            raise ValueError('dtype must be "fp32" or "bf16"')
        self.compute_dtype = dtype
        self.linear_secondlast_layer = nn.Linear(in_features, embed_dim)     # parameter containers only
        self.last_layer = nn.BatchNorm1d(embed_dim, momentum=0.01)
        self.linear_secondlast_layer.weight.data.normal_(0, 0.05)            # cnn.py:41
        self.last_layer.bias.data.fill_(0)                                   # cnn.py:42

    def forward(self, x):
        """x: (B, in_features) pooled trunk output (detached, cnn.py:46-47) -> cnn_feature (B, embed_dim)."""
        lin, bn = self.linear_secondlast_layer, self.last_layer
        return _HeadFn.apply(self, x.reshape(x.shape[0], -1), lin.weight, lin.bias, bn.weight, bn.bias)
