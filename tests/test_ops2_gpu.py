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


@pytest.mark.parametrize("wd", [1e-4, 5e-2])
def test_adamw_cold_blocks_bit_identical_to_dense(wd):
    """Blocks of 256 elements that have never received a gradient (word-embedding rows no batch has touched) take the
    shortcut p -= lr*wd*p; the result must be bit-identical to the dense update, and the norm pass must mark exactly
    the blocks with non-zero gradients."""
    import math
    from clg_vqa_b200 import _lib as L
    rows, dim = 4000, 768
    k = rows * dim
    n = (k + L.CHUNK - 1) // L.CHUNK * L.CHUNK + 2 * L.CHUNK
    k2 = 1000                                    # a second, non-decayed tensor that never gets gradients
    start2 = n - L.CHUNK
    torch.manual_seed(3)
    p0 = torch.zeros(n, device=DEV); p0[:k] = torch.randn(k, device=DEV) * 0.05; p0[start2:start2 + k2] = 1.0
    chunk_tensor = torch.full((n // L.CHUNK,), -1, dtype=torch.int32)
    chunk_tensor[:(k + L.CHUNK - 1) // L.CHUNK] = 0
    chunk_tensor[start2 // L.CHUNK] = 1
    metas = (L.TensorMeta * 2)()
    # wd = 1e-4: lr*wd = 4e-9 < 2^-25, the decay of a cold block rounds to the identity and the block is skipped entirely;
    # wd = 5e-2: lr*wd = 2e-6, cold blocks take the read-decay-write shortcut.  Both must equal the dense update bit for bit.
    metas[0] = L.TensorMeta(4e-5, wd, 0, -1, 0, k)
    metas[1] = L.TensorMeta(4e-5, 0.0, -1, -1, start2, k2)
    chunk_tensor = chunk_tensor.to(DEV)
    tens = torch.frombuffer(bytearray(bytes(metas)), dtype=torch.uint8).to(DEV)
    state = {}
    for mode in ("dense", "hot", "hot+zero", "dense+zero"):
        p = p0.clone(); m = torch.zeros(n, device=DEV); v = torch.zeros(n, device=DEV)
        # the kernel's precondition (Engine._refresh keeps it): the shadow already holds bf16(p) - blocks the step does not
        # touch (cold, decay below half an ulp) keep their shadow as it is
        shadow = p0.to(torch.bfloat16)
        hot = torch.zeros(n // 256, device=DEV, dtype=torch.uint8)
        partials = torch.zeros(1024, device=DEV); ss = torch.zeros(1, device=DEV)
        gen = torch.Generator(device="cpu").manual_seed(11)
        touched = torch.zeros(rows, dtype=torch.bool)
        for step in range(1, 5):
            g = torch.zeros(n, device=DEV)
            ids = torch.randperm(rows, generator=gen)[:150]          # unique: index_put with duplicates is not deterministic
            touched[ids] = True
            g.view(-1)[:k].view(rows, dim)[ids.to(DEV)] = torch.randn(150, dim, generator=gen).to(DEV)
            ss.zero_()
            g[n - 2 * L.CHUNK + 5] = 3.0                  # a chunk without an optimiser entry (chunk_tensor == -1)
            use_hot = mode.startswith("hot")
            ops.sumsq(g[:n - 2 * L.CHUNK], partials, ss, hot=hot if use_hot else None)
            ops.adamw_step(p, g, m, v, chunk_tensor, tens, sumsq=ss, max_norm=1.0, eps=1e-6, bias_c1=1 - 0.9 ** step,
                           bias_c2=1 - 0.999 ** step, shadow=shadow, hot=hot if use_hot else None,
                           zero_grad=mode.endswith("zero"))
            # zero_grad: the step leaves the whole gradient arena cleared (also where no optimiser entry exists)
            assert (int(torch.count_nonzero(g)) == 0) == mode.endswith("zero")
        state[mode] = (p, m, v, shadow, hot, touched)
    for other in ("hot", "hot+zero", "dense+zero"):
        for a, b in zip(state["dense"][:4], state[other][:4]):
            assert torch.equal(a, b)
    hot, touched = state["hot"][4], state["hot"][5]
    assert torch.equal(hot[:k // 256].view(rows, 3).bool().cpu(), touched[:, None].expand(rows, 3))
    assert int(hot[k // 256:].sum()) == 0
    assert 0.05 < float(touched.float().mean()) < 0.5           # the test exercises both paths
    assert torch.equal(state["hot"][0][start2:start2 + k2], torch.ones(k2, device=DEV))


def test_segment_sum_rows_deterministic_reduction():
    """data-parallel word-gradient reduction: local pre-reduction (duplicates summed into the first row, other copies renamed
    to the padding id) followed by the table reduction equals index_add; two runs give bit-identical tables"""
    n, d, V, pad = 5000, 768, 300, 1
    g = torch.Generator(device="cpu").manual_seed(5)
    ids = torch.randint(0, V, (n,), generator=g)
    ids[::7] = 0; ids[3::11] = 2                                 # long runs (<s>, </s>)
    rows = torch.randn(n, d, generator=g)
    ref = torch.zeros(V, d, dtype=torch.float64).index_add_(0, ids, rows.double())
    ref[pad] = 0
    outs = []
    for _ in range(2):
        ids_d, rows_d = ids.to(DEV), rows.to(DEV).clone()
        ids_out = ids_d.clone()
        s_ids, order = torch.sort(ids_d, stable=True)
        ops.segment_sum_rows(rows_d, order, s_ids, None, pad, ids_out=ids_out)
        assert int((ids_out != pad).sum()) == len(set(ids.tolist()) - {pad})          # one surviving row per distinct id
        s2, o2 = torch.sort(ids_out, stable=True)
        table = torch.zeros(V, d, device=DEV)
        ops.segment_sum_rows(rows_d, o2, s2, table, pad)
        outs.append(table)
        assert float((table.double().cpu() - ref).abs().max()) < 1e-3
    assert torch.equal(outs[0], outs[1])
    # direct (no pre-reduction) gives the same values up to summation order
    table2 = torch.zeros(V, d, device=DEV)
    s_ids, order = torch.sort(ids.to(DEV), stable=True)
    ops.segment_sum_rows(rows.to(DEV), order, s_ids, table2, pad)
    assert float((table2.double().cpu() - ref).abs().max()) < 1e-3


def test_sumsq_sparse_reads_only_marked_blocks():
    """clg_sumsq_sparse + clg_mark_rows: the norm of an arena whose word-table range is zero except for the rows of the
    batch's token ids equals the dense norm; unmarked blocks are really skipped (a poisoned unmarked block is not seen)."""
    rows, dim = 5000, 768
    first = 4096                                      # table starts at element 4096 (block 16)
    n = first + rows * dim + 4096
    n = n // 256 * 256
    torch.manual_seed(5)
    g = torch.zeros(n, device=DEV)
    g[:first] = torch.randn(first, device=DEV)
    g[first + rows * dim:] = torch.randn(n - first - rows * dim, device=DEV)
    ids = torch.randint(0, rows, (64, 40), device=DEV)
    ids[:, 30:] = 1                                   # padding id: never receives a gradient
    tab = g[first:first + rows * dim].view(rows, dim)
    uniq = torch.unique(ids[ids != 1])
    tab[uniq] = torch.randn(uniq.numel(), dim, device=DEV)
    touched = torch.zeros(n // 256, dtype=torch.uint8, device=DEV)
    ops.mark_rows(ids, dim, first, rows, 1, touched)
    lo, hi = first // 256, (first + rows * dim) // 256
    want = torch.zeros(rows, dtype=torch.bool, device=DEV); want[uniq] = True
    assert torch.equal(touched[lo:hi].view(rows, 3).bool(), want[:, None].expand(rows, 3))
    assert int(touched[:lo].sum()) == 0 and int(touched[hi:].sum()) == 0
    partials = torch.zeros(1024, device=DEV)
    dense = torch.zeros(1, device=DEV); sparse = torch.zeros(1, device=DEV)
    hot_d = torch.zeros(n // 256, dtype=torch.uint8, device=DEV); hot_s = torch.zeros_like(hot_d)
    ops.sumsq(g, partials, dense, hot=hot_d)
    ops.sumsq(g, partials, sparse, hot=hot_s, sparse=(lo, hi, touched))
    ref = float(g.double().pow(2).sum())
    assert abs(float(dense) - ref) <= 1e-5 * ref and abs(float(sparse) - ref) <= 1e-5 * ref
    assert torch.equal(hot_d, hot_s)
    # an unmarked block is not read at all
    free = int(torch.nonzero(~want)[0])
    tab[free] = float("nan")
    sparse.zero_()
    ops.sumsq(g, partials, sparse, sparse=(lo, hi, touched))
    assert abs(float(sparse) - ref) <= 1e-5 * ref


def test_fused_adamw_matches_reference_tree_trajectories():
    """the fused multi-tensor AdamW kernel against tests/golden/adamw_ref_step.npz: trajectories produced by the step body
    vendored in the reference tree (features_extraction/mmf/mmf/modules/optimizers.py:L15-84, run unmodified by
    oracle/make_golden_adamw.py).  fp32 with another operation order (fma contraction, rsqrt-free sqrt): 2e-6 relative."""
    import os
    import numpy as np
    from clg_vqa_b200 import _lib as L
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "adamw_ref_step.npz"), allow_pickle=False)
    names = sorted({k.split("::")[0] for k in g.files if "::" in k})
    for name in names:
        lr, b1, b2, eps, wd, cb = [float(x) for x in g[name + "::hyper"]]
        p0 = torch.from_numpy(g[name + "::p0"]).reshape(-1)
        k = p0.numel()
        n = (k + L.CHUNK - 1) // L.CHUNK * L.CHUNK
        p = torch.zeros(n, device=DEV); p[:k] = p0.to(DEV)
        m = torch.zeros(n, device=DEV); v = torch.zeros(n, device=DEV)
        chunk_tensor = torch.zeros(n // L.CHUNK, dtype=torch.int32, device=DEV)
        metas = (L.TensorMeta * 1)()
        metas[0] = L.TensorMeta(lr, wd, -1, -1, 0, k)
        tens = torch.frombuffer(bytearray(bytes(metas)), dtype=torch.uint8).to(DEV)
        grads = torch.from_numpy(g[name + "::grads"])
        want = torch.from_numpy(g[name + "::params"])
        for t in range(grads.shape[0]):
            gr = torch.zeros(n, device=DEV); gr[:k] = grads[t].reshape(-1).to(DEV)
            c1 = 1 - b1 ** (t + 1) if cb else 1.0
            c2 = 1 - b2 ** (t + 1) if cb else 1.0
            ops.adamw_step(p, gr, m, v, chunk_tensor, tens, sumsq=None, max_norm=0.0, beta1=b1, beta2=b2, eps=eps, bias_c1=c1, bias_c2=c2)
            ref = want[t].reshape(-1).to(DEV)
            err = float((p[:k] - ref).abs().max() / ref.abs().max())
            assert err <= 2e-6, (name, t, err)
        assert float((m[:k].cpu() - torch.from_numpy(g[name + "::exp_avg"]).reshape(-1)).abs().max()) <= 1e-6 * float(np.abs(g[name + "::exp_avg"]).max())
