"""More per-kernel GPU parity: fused GeLU + derivative epilogue pair, staged epilogue edge cases
(ragged M / N tails, strided outputs), generic-N LayerNorm."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

from clg_vqa_b200 import ops  # noqa: E402

DEV = "cuda"


def _rand(*shape, dtype=torch.float32, scale=1.0, seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(DEV).to(dtype)


def _rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


@pytest.mark.parametrize("backend", [ops.GEMM_TC, ops.GEMM_SIMT])
@pytest.mark.parametrize("cdt", [torch.bfloat16, torch.float32])
def test_gelu_grad_pair(backend, cdt):
    M, N, K = 700, 3072, 256
    a = _rand(M, K, dtype=torch.bfloat16, seed=1)
    w = _rand(N, K, dtype=torch.bfloat16, scale=0.08, seed=2)
    bias = _rand(N, seed=3)
    u = (a.double() @ w.double().t() + bias.double())
    h = torch.empty(M, N, device=DEV, dtype=cdt)
    gp = torch.empty(M, N, device=DEV, dtype=cdt)
    ops.gemm(a, w, h, bias=bias, act=ops.ACT_GELU_GRAD, aux_out=gp, backend=backend)
    ref_h = torch.nn.functional.gelu(u)
    ref_g = 0.5 * (1 + torch.erf(u / math.sqrt(2))) + u * torch.exp(-0.5 * u * u) / math.sqrt(2 * math.pi)
    tol = 6e-3 if cdt == torch.bfloat16 else 2e-6
    assert float((h.double() - ref_h).abs().max()) <= tol * float(ref_h.abs().max())
    assert float((gp.double() - ref_g).abs().max()) <= tol * float(ref_g.abs().max())
    # backward side: v *= aux
    dy = _rand(M, K, dtype=torch.bfloat16, seed=4)
    out = torch.empty(M, N, device=DEV, dtype=cdt)
    ops.gemm(dy, w, out, act=ops.ACT_MUL_AUX, aux_in=gp, backend=backend)
    ref = (dy.double() @ w.double().t()) * gp.double()
    assert _rel(out, ref) <= (6e-3 if cdt == torch.bfloat16 else 2e-5)


@pytest.mark.parametrize("M,N", [(130, 776), (257, 1848), (128, 96), (1000, 72)])
def test_staged_epilogue_tails_and_strides(M, N):
    K = 192
    a = _rand(M, K, dtype=torch.bfloat16, seed=5)
    w = _rand(N, K, dtype=torch.bfloat16, scale=0.05, seed=6)
    bias = _rand(N, seed=7)
    big = torch.zeros(M, N + 40, device=DEV)
    res_big = _rand(M, N + 8, seed=8)
    c = big[:, 8:8 + N]                      # strided output view (ldc = N + 40, 32 B offset)
    res = res_big[:, :N]
    ops.gemm(a, w, c, bias=bias, residual=res, backend=ops.GEMM_TC)
    ref = a.double() @ w.double().t() + bias.double() + res.double()
    assert _rel(c, ref) < 2e-5
    assert float(big[:, :8].abs().max()) == 0.0 and float(big[:, 8 + N:].abs().max()) == 0.0
    cb = torch.empty(M, N, device=DEV, dtype=torch.bfloat16)
    ops.gemm(a, w, cb, bias=bias, backend=ops.GEMM_TC)
    assert _rel(cb, (a.double() @ w.double().t() + bias.double())) < 1e-2
    acc = _rand(M, N, seed=9)
    acc0 = acc.clone()
    ops.gemm(a, w, acc, accumulate=True, backend=ops.GEMM_TC)
    assert _rel(acc, acc0.double() + a.double() @ w.double().t()) < 2e-5


@pytest.mark.parametrize("N", [200, 1000])
def test_layernorm_generic_width(N):
    rows = 333
    x = _rand(rows, N, seed=10)
    g = 1 + 0.1 * _rand(N, seed=11)
    b = 0.1 * _rand(N, seed=12)
    dy = _rand(rows, N, seed=13)
    y = torch.empty_like(x); mean = torch.empty(rows, device=DEV); rstd = torch.empty(rows, device=DEV)
    ops.layernorm_fwd(x, g, b, 1e-5, y=y, mean=mean, rstd=rstd)
    xr = x.double().requires_grad_(True); gr = g.double().requires_grad_(True); br = b.double().requires_grad_(True)
    ref = torch.nn.functional.layer_norm(xr, (N,), gr, br, 1e-5)
    assert _rel(y, ref) < 2e-6
    ref.backward(dy.double())
    dx = torch.empty_like(x); dg = torch.zeros(N, device=DEV); db = torch.zeros(N, device=DEV)
    ops.layernorm_bwd(dy, x, g, mean, rstd, dx=dx, dgamma=dg, dbeta=db)
    assert _rel(dx, xr.grad) < 1e-5 and _rel(dg, gr.grad) < 1e-5 and _rel(db, br.grad) < 1e-5
