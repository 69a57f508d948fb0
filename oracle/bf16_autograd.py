"""bf16 operand rounding for BOTH passes ("bf16 training oracle").  TEST INFRASTRUCTURE ONLY.

SURVEY.md Appendix A.5 defines the bf16 oracle of the forward: the two operands of every F.linear /
torch.matmul are rounded to bf16, accumulation and everything else stay fp32.  A bf16 training step
rounds the operands of the backward contractions as well (dX = dY W, dW = dY^T X, dA = dC B^T,
dB = A^T dC all read bf16 dY / dC and the bf16 copies of the saved operands); these two autograd
Functions state exactly that, so that gradients of the CUDA path can be compared with a reference
whose rounding points are the same in the backward too.  They are installed over F.linear /
torch.matmul of the UNMODIFIED reference by oracle/ref_shim.bf16_train_shim() and used directly by the
oracle ports (mm_dtype=torch.bfloat16, round_bwd=True).  Bias gradients stay fp32 sums of the
incoming gradient.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

_lin = F.linear            # captured before any shim patches them
_mm = torch.matmul


def _r(x):
    return x.bfloat16().to(x.dtype)


class LinearBf16(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b):
        xr, wr = _r(x), _r(w)
        ctx.save_for_backward(xr, wr)
        ctx.has_bias = b is not None
        return _lin(xr, wr, b)

    @staticmethod
    def backward(ctx, g):
        xr, wr = ctx.saved_tensors
        gr = _r(g)
        gx = _mm(gr, wr)
        gw = _mm(gr.reshape(-1, gr.shape[-1]).t(), xr.reshape(-1, xr.shape[-1]))
        gb = g.reshape(-1, g.shape[-1]).sum(0) if ctx.has_bias else None
        return gx, gw, gb


class MatmulBf16(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b):
        ar, br = _r(a), _r(b)
        ctx.save_for_backward(ar, br)
        ctx.shapes = (a.shape, b.shape)
        return _mm(ar, br)

    @staticmethod
    def backward(ctx, g):
        ar, br = ctx.saved_tensors
        gr = _r(g)
        ga = _mm(gr, br.transpose(-1, -2))
        gb = _mm(ar.transpose(-1, -2), gr)
        sa, sb = ctx.shapes
        if ga.shape != sa:
            ga = ga.sum_to_size(sa)
        if gb.shape != sb:
            gb = gb.sum_to_size(sb)
        return ga, gb


def linear_bf16(x, w, b=None):
    if not x.is_floating_point():
        return _lin(x, w, b)
    return LinearBf16.apply(x, w, b)


def matmul_bf16(a, b):
    if not a.is_floating_point() or a.dim() < 2 or b.dim() < 2:
        return _mm(a, b)
    return MatmulBf16.apply(a, b)
