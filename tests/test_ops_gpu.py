"""Per-kernel parity on the GPU: every C-ABI op against a plain PyTorch fp32/fp64 reference of the
same op on the same seeded inputs (bit-exact for the integer / mask work, tight fp tolerances
otherwise).  Model-level parity against the oracle and the golden fixtures is in
test_model_gpu.py."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

from clg_vqa_b200 import ops  # noqa: E402
from clg_vqa_b200.ops import DropSite  # noqa: E402

DEV = "cuda"
PQ = round(0.1 * 65536) / 65536.0      # dropout probability as quantised by the kernels
KEEP = 1.0 - PQ


def _rand(*shape, dtype=torch.float32, scale=1.0, seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(DEV).to(dtype)


def _rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


# ------------------------------------------------------------------------------------------------
# GEMM
# ------------------------------------------------------------------------------------------------
def _gemm_ref(a, b, a_k, b_k):
    A = a.double() if a_k else a.double().t()
    B = b.double() if b_k else b.double().t()
    return A @ B.t()


@pytest.mark.parametrize("backend", [ops.GEMM_TC, ops.GEMM_SIMT])
@pytest.mark.parametrize("a_k,b_k", [(True, True), (True, False), (False, False), (False, True)])
@pytest.mark.parametrize("M,N,K", [(256, 256, 128), (608, 768, 768), (1000, 2304, 200), (768, 768, 4864), (128, 64, 64)])
def test_gemm_plain(backend, a_k, b_k, M, N, K):
    a = _rand(M, K, dtype=torch.bfloat16, seed=1) if a_k else _rand(K, M, dtype=torch.bfloat16, seed=1)
    b = _rand(N, K, dtype=torch.bfloat16, seed=2) if b_k else _rand(K, N, dtype=torch.bfloat16, seed=2)
    c = torch.full((M, N), float("nan"), device=DEV, dtype=torch.float32)
    ops.gemm(a, b, c, a_kmajor=a_k, b_kmajor=b_k, backend=backend)
    ref = _gemm_ref(a, b, a_k, b_k)
    err = _rel(c, ref)
    assert err < 2e-5, "rel err %g" % err


@pytest.mark.parametrize("backend", [ops.GEMM_TC, ops.GEMM_SIMT])
def test_gemm_epilogues(backend):
    M, N, K = 640, 768, 256
    a = _rand(M, K, dtype=torch.bfloat16, seed=3)
    w = _rand(N, K, dtype=torch.bfloat16, scale=0.05, seed=4)
    bias = _rand(N, seed=5)
    res = _rand(M, N, seed=6)
    lin = a.double() @ w.double().t() + bias.double()
    # bias -> bf16
    c = torch.empty(M, N, device=DEV, dtype=torch.bfloat16)
    ops.gemm(a, w, c, bias=bias, backend=backend)
    assert _rel(c, lin.bfloat16()) < 1e-2
    assert float((c.double() - lin).abs().max()) < 0.05
    # bias + gelu, pre-activation saved
    u = torch.empty(M, N, device=DEV, dtype=torch.float32)
    h = torch.empty(M, N, device=DEV, dtype=torch.float32)
    ops.gemm(a, w, h, bias=bias, act=ops.ACT_GELU, aux_out=u, backend=backend)
    assert _rel(u, lin) < 2e-5
    assert _rel(h, torch.nn.functional.gelu(lin)) < 2e-5
    # bias + residual -> fp32
    y = torch.empty(M, N, device=DEV, dtype=torch.float32)
    ops.gemm(a, w, y, bias=bias, residual=res, backend=backend)
    assert _rel(y, lin + res.double()) < 2e-5
    # dgelu with bf16 aux
    ub = u.bfloat16()
    d = torch.empty(M, N, device=DEV, dtype=torch.bfloat16)
    ops.gemm(a, w, d, act=ops.ACT_DGELU, aux_in=ub, backend=backend)
    x = ub.double()
    dg = 0.5 * (1 + torch.erf(x / math.sqrt(2))) + x * torch.exp(-0.5 * x * x) / math.sqrt(2 * math.pi)
    refd = (a.double() @ w.double().t()) * dg
    assert float((d.double() - refd).abs().max() / refd.abs().max()) < 1e-2
    # relu / drelu / tanh
    r = torch.empty(M, N, device=DEV, dtype=torch.float32)
    ops.gemm(a, w, r, bias=bias, act=ops.ACT_RELU, backend=backend)
    assert _rel(r, torch.relu(lin)) < 2e-5
    ops.gemm(a, w, r, bias=bias, act=ops.ACT_TANH, backend=backend)
    assert _rel(r, torch.tanh(lin)) < 2e-5
    ops.gemm(a, w, r, act=ops.ACT_DRELU, aux_in=res, backend=backend)
    assert _rel(r, (a.double() @ w.double().t()) * (res > 0)) < 2e-5
    # accumulate + alpha
    acc0 = _rand(M, N, seed=8)
    acc = acc0.clone()
    ops.gemm(a, w, acc, alpha=0.5, accumulate=True, backend=backend)
    assert _rel(acc, acc0.double() + 0.5 * (a.double() @ w.double().t())) < 2e-5


def test_gemm_mask_bits_and_splitk():
    # wgrad shape: dW[N_w, K_w] = dY^T X with tokens as the reduction dim, masked, split-K
    T, Nw, Kw = 4864, 768, 768
    dy = _rand(T, Nw, dtype=torch.bfloat16, seed=9)
    x = _rand(T, Kw, dtype=torch.bfloat16, seed=10)
    mask = (torch.rand(Nw, Kw, device=DEV) > 0.4).float()
    bits = torch.zeros((Nw * Kw + 31) // 32, device=DEV, dtype=torch.int32)
    ops.pack_mask(mask, bits)
    ref = (dy.double().t() @ x.double()) * mask.double()
    for backend in (ops.GEMM_TC, ops.GEMM_SIMT):
        dw = torch.full((Nw, Kw), float("nan"), device=DEV, dtype=torch.float32)
        ops.gemm(dy, x, dw, a_kmajor=False, b_kmajor=False, mask_bits=bits, backend=backend)
        assert _rel(dw, ref) < 3e-5
        assert torch.equal(dw[mask == 0], torch.zeros_like(dw[mask == 0]))   # exactly zero where masked
        dw2 = dw.clone()
        ops.gemm(dy, x, dw2, a_kmajor=False, b_kmajor=False, mask_bits=bits, accumulate=True, backend=backend)
        assert _rel(dw2, 2 * ref) < 3e-5
    dw = torch.empty(Nw, Kw, device=DEV, dtype=torch.float32)
    ops.gemm(dy, x, dw, a_kmajor=False, b_kmajor=False, backend=ops.GEMM_TC, split_k=7)
    assert _rel(dw, dy.double().t() @ x.double()) < 3e-5


def test_gemm_fp32_simt_and_strided():
    M, N, K = 200, 1842, 768
    a_full = _rand(M, 3 * K, seed=11)
    a = a_full[:, K:2 * K]                       # strided A (lda = 3K)
    w = _rand(N, K, scale=0.05, seed=12)
    bias = _rand(N, seed=13)
    c = torch.empty(M, N, device=DEV, dtype=torch.float32)
    ops.gemm(a, w, c, bias=bias)
    assert _rel(c, a.double() @ w.double().t() + bias.double()) < 2e-6
    # K = 7 location projection
    loc = _rand(500, 7, seed=14)
    wl = _rand(768, 7, seed=15)
    o = torch.empty(500, 768, device=DEV)
    ops.gemm(loc, wl, o)
    assert _rel(o, loc.double() @ wl.double().t()) < 2e-6


def test_gemm_dropout_consistency():
    M, N, K = 384, 768, 128
    a = _rand(M, K, dtype=torch.bfloat16, seed=16)
    w = _rand(N, K, dtype=torch.bfloat16, seed=17)
    site = DropSite(0.1, seed=1234, offset=1 << 20)
    y0 = torch.empty(M, N, device=DEV)
    y1 = torch.empty(M, N, device=DEV)
    ops.gemm(a, w, y0, backend=ops.GEMM_TC)
    ops.gemm(a, w, y1, dropout=site, backend=ops.GEMM_TC)
    kept = y1 != 0
    frac = float(kept.float().mean())
    assert abs(frac - 0.9) < 0.01
    assert _rel(y1[kept], y0[kept] / KEEP) < 1e-5
    y2 = torch.empty(M, N, device=DEV)
    ops.gemm(a, w, y2, dropout=site, backend=ops.GEMM_SIMT)       # same mask from both backends
    assert torch.equal(y2 != 0, kept)


# ------------------------------------------------------------------------------------------------
# LayerNorm
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("N", [768, 1536, 128])
def test_layernorm_fwd_bwd(N):
    rows = 1000
    x = _rand(rows, N, seed=20)
    g = 1 + 0.1 * _rand(N, seed=21)
    b = 0.1 * _rand(N, seed=22)
    dy = _rand(rows, N, seed=23)
    y = torch.empty_like(x)
    yl = torch.empty(rows, N, device=DEV, dtype=torch.bfloat16)
    mean = torch.empty(rows, device=DEV)
    rstd = torch.empty(rows, device=DEV)
    ops.layernorm_fwd(x, g, b, 1e-5, y=y, y_lowp=yl, mean=mean, rstd=rstd)
    xr = x.double().requires_grad_(True)
    gr = g.double().requires_grad_(True)
    br = b.double().requires_grad_(True)
    ref = torch.nn.functional.layer_norm(xr, (N,), gr, br, 1e-5)
    assert _rel(y, ref) < 2e-6
    assert torch.equal(yl, y.bfloat16())
    ref.backward(dy.double())
    dx = torch.empty_like(x)
    dxl = torch.empty(rows, N, device=DEV, dtype=torch.bfloat16)
    dg = torch.zeros(N, device=DEV)
    db = torch.zeros(N, device=DEV)
    dbias = torch.zeros(N, device=DEV)
    ops.layernorm_bwd(dy, x, g, mean, rstd, dx=dx, dx_lowp=dxl, dgamma=dg, dbeta=db, dbias=dbias)
    assert _rel(dx, xr.grad) < 1e-5
    assert _rel(dg, gr.grad) < 1e-5
    assert _rel(db, br.grad) < 1e-5
    assert _rel(dbias, xr.grad.sum(0)) < 1e-4
    assert torch.equal(dxl, dx.bfloat16())


def test_layernorm_rowmap_add_valid():
    B, T, R, N = 6, 40, 36, 768
    S = T + R
    x = _rand(B * R, N, seed=30)
    g = 1 + 0.1 * _rand(N, seed=31)
    b = 0.1 * _rand(N, seed=32)
    addm = _rand(B * R, N, seed=33)
    addv = _rand(N, seed=34)
    joint = torch.zeros(B * S, N, device=DEV)
    ops.layernorm_fwd(x, g, b, 1e-5, y=joint, out_map=ops.rowmap(R, T, S), add_mat=addm, add_vec=addv)
    ref = torch.nn.functional.layer_norm(x.double(), (N,), g.double(), b.double(), 1e-5) + addm.double() + addv.double()
    got = joint.view(B, S, N)[:, T:, :].reshape(B * R, N)
    assert _rel(got, ref) < 2e-6
    assert float(joint.view(B, S, N)[:, :T, :].abs().max()) == 0.0
    # M3P-style row validity (pre and post) and backward through the map
    valid = (torch.rand(B * R, device=DEV) > 0.3).to(torch.uint8)
    y = torch.empty(B * R, N, device=DEV)
    mean = torch.empty(B * R, device=DEV)
    rstd = torch.empty(B * R, device=DEV)
    ops.layernorm_fwd(x, g, b, 1e-12, y=y, mean=mean, rstd=rstd, row_valid=valid, valid_pre=True, valid_post=True)
    v = valid.double().unsqueeze(1)
    xr = x.double().requires_grad_(True)
    ref = torch.nn.functional.layer_norm(xr * v, (N,), g.double(), b.double(), 1e-12) * v
    assert _rel(y, ref) < 2e-6
    dyj = _rand(B * S, N, seed=35)
    dsel = dyj.view(B, S, N)[:, T:, :].reshape(B * R, N)
    ref.backward(dsel.double())
    dx = torch.empty(B * R, N, device=DEV)
    ops.layernorm_bwd(dyj, x, g, mean, rstd, dx=dx, dy_map=ops.rowmap(R, T, S), row_valid=valid, valid_pre=True, valid_post=True)
    # rows zeroed before the LN have var = 0 -> rstd = 1e6; their dx is multiplied by pre = 0
    assert _rel(dx, xr.grad) < 1e-4


def test_layernorm_dropout_sites():
    rows, N = 512, 768
    x = _rand(rows, N, seed=40)
    g = 1 + 0.1 * _rand(N, seed=41)
    b = 0.1 * _rand(N, seed=42)
    out_site = DropSite(0.1, 7, 0)
    in_site = DropSite(0.1, 7, 1 << 30)
    y0 = torch.empty_like(x); y1 = torch.empty_like(x)
    mean = torch.empty(rows, device=DEV); rstd = torch.empty(rows, device=DEV)
    ops.layernorm_fwd(x, g, b, 1e-5, y=y0, mean=mean, rstd=rstd)
    ops.layernorm_fwd(x, g, b, 1e-5, y=y1, dropout=out_site)
    keep = (y1 != 0) | (y0 == 0)
    assert abs(float(keep.float().mean()) - 0.9) < 0.01
    assert _rel(y1[keep], y0[keep] / KEEP) < 1e-6
    dy = _rand(rows, N, seed=43)
    dx_ref = torch.empty_like(x)
    ops.layernorm_bwd((dy * keep / KEEP).contiguous(), x, g, mean, rstd, dx=dx_ref)
    dx = torch.empty_like(x); dxl = torch.empty_like(x)
    ops.layernorm_bwd(dy, x, g, mean, rstd, dx=dx, dx_lowp=dxl, dropout_out=out_site, dropout_in=in_site)
    assert _rel(dx, dx_ref) < 1e-6
    kin = dxl != 0
    assert abs(float(kin.float().mean()) - 0.9) < 0.01
    assert _rel(dxl[kin], dx[kin] / KEEP) < 1e-6


# ------------------------------------------------------------------------------------------------
# attention
# ------------------------------------------------------------------------------------------------
def _attn_ref(qkv, key_bias, B, S, H, scale, round_bf16):
    D = H * 64
    q, k, v = [t.view(B, S, H, 64).permute(0, 2, 1, 3) for t in qkv.double().split(D, dim=1)]
    s = (q @ k.transpose(-1, -2)) * scale + key_bias.double()[:, None, None, :]
    p = torch.softmax(s, dim=-1)
    if round_bf16:
        p = p.float().bfloat16().double()
    return (p @ v).permute(0, 2, 1, 3).reshape(B * S, D)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("S,minf", [(76, False), (140, True)])
def test_attention_fwd_bwd(dtype, S, minf):
    B, H = 3, 12
    D = H * 64
    qkv = _rand(B * S, 3 * D, dtype=dtype, scale=1.0, seed=50)
    nvalid = torch.tensor([S, S - 9, S // 2], device=DEV)
    valid = torch.arange(S, device=DEV)[None, :] < nvalid[:, None]
    key_bias = torch.where(valid, 0.0, float("-inf") if minf else -10000.0).float().contiguous()
    ctx = torch.empty(B * S, D, device=DEV, dtype=dtype)
    lse = torch.empty(B, H, S, device=DEV)
    ops.attention_fwd(qkv, key_bias, ctx, lse, B, S, H, 0.125)
    qr = qkv.double().requires_grad_(True)
    D_ = D
    q, k, v = [t.view(B, S, H, 64).permute(0, 2, 1, 3) for t in qr.split(D_, dim=1)]
    s = (q @ k.transpose(-1, -2)) * 0.125 + key_bias.double()[:, None, None, :]
    p = torch.softmax(s, dim=-1)
    ref = (p @ v).permute(0, 2, 1, 3).reshape(B * S, D)
    tol = 2e-5 if dtype == torch.float32 else 1.5e-2
    assert _rel(ctx, ref) < tol
    assert _rel(lse, torch.logsumexp(s, dim=-1)) < 1e-5
    dctx = _rand(B * S, D, dtype=dtype, seed=51)
    ref.backward(dctx.double())
    dqkv = torch.empty_like(qkv)
    ops.attention_bwd(qkv, key_bias, ctx, dctx, lse, dqkv, B, S, H, 0.125)
    assert _rel(dqkv, qr.grad) < (1e-4 if dtype == torch.float32 else 2e-2)


def test_attention_dropout_stats():
    B, S, H = 4, 76, 12
    D = H * 64
    qkv = _rand(B * S, 3 * D, seed=52)
    kb = torch.zeros(B, S, device=DEV)
    c0 = torch.empty(B * S, D, device=DEV); c1 = torch.empty(B * S, D, device=DEV)
    lse = torch.empty(B, H, S, device=DEV)
    ops.attention_fwd(qkv, kb, c0, lse, B, S, H, 0.125)
    ops.attention_fwd(qkv, kb, c1, lse, B, S, H, 0.125, DropSite(0.1, 99, 0))
    # dropout keeps the expectation: mean over many elements matches within noise
    assert abs(float(c1.mean() - c0.mean())) < 5e-3
    assert float((c1 - c0).abs().mean()) > 1e-3


# ------------------------------------------------------------------------------------------------
# embeddings, loss, optimiser, utilities
# ------------------------------------------------------------------------------------------------
def test_embed_text_fwd_bwd():
    B, T, N, V = 5, 40, 768, 1000
    g = torch.Generator().manual_seed(60)
    lens = torch.randint(6, T + 1, (B,), generator=g)
    ids = torch.randint(3, V, (B, T), generator=g)
    ids[torch.arange(T)[None, :] >= lens[:, None]] = 1
    ids[:, 0] = 0
    ids = ids.to(DEV)
    seg = torch.randint(0, 2, (B, T), generator=g).to(DEV)
    word = _rand(V, N, seed=61); pos = _rand(514, N, seed=62); typ = _rand(2, N, seed=63)
    S = T + 36
    out = torch.zeros(B * S, N, device=DEV)
    ops.embed_text_fwd(ids, seg, word, pos, typ, out, pad_id=1, out_map=ops.rowmap(T, 0, S))
    m = ids.ne(1).int()
    pid = (torch.cumsum(m, 1) * m).long() + 1
    ref = word[ids] + pos[pid] + typ[seg]
    got = out.view(B, S, N)[:, :T]
    assert torch.allclose(got, ref, rtol=0, atol=1e-6)
    dout = _rand(B * S, N, seed=64)
    dw = torch.zeros_like(word); dp = torch.zeros_like(pos); dt = torch.zeros_like(typ)
    ops.embed_text_bwd(ids, seg, dout, dw, dp, dt, pad_id=1, dout_map=ops.rowmap(T, 0, S))
    d = dout.view(B, S, N)[:, :T].reshape(B * T, N).double()
    rw = torch.zeros_like(word, dtype=torch.float64).index_add_(0, ids.reshape(-1), d)
    rw[1] = 0   # padding_idx
    rp = torch.zeros_like(pos, dtype=torch.float64).index_add_(0, pid.reshape(-1), d)
    rt = torch.zeros_like(typ, dtype=torch.float64).index_add_(0, seg.reshape(-1), d)
    assert _rel(dw, rw) < 1e-5 and _rel(dp, rp) < 1e-5 and _rel(dt, rt) < 1e-5
    assert float(dw[1].abs().max()) == 0.0


def _loss_ref(z, target, dist, lam):
    z = z.double().requires_grad_(True)
    p = torch.softmax(z, -1)
    pk, ik = torch.topk(p, 10)
    sem = (pk * dist.double()[torch.arange(z.shape[0]).unsqueeze(1), ik]).sum(-1).mean()
    y = torch.argmax(target.long(), 1)
    loss = torch.nn.functional.cross_entropy(z, y) * target.shape[1] + lam * sem * target.shape[1]
    loss.backward()
    score = target[torch.arange(z.shape[0]), z.argmax(1)].sum() / z.shape[0]
    return loss.detach(), score, z.grad


@pytest.mark.parametrize("B", [1, 16, 256])
def test_gqa_prior_loss(B):
    A = 1842
    z = _rand(B, A, scale=2.0, seed=70)
    g = torch.Generator().manual_seed(71)
    lab = torch.randint(0, A, (B,), generator=g)
    target = torch.zeros(B, A)
    target[torch.arange(B), lab] = 1.0
    if B > 1:
        target[1] = 0.0            # all-zero row -> class 0 (argmax of zeros)
        target[0, 5] = 0.7         # soft score < 1 truncates to 0
    z[0, lab[0]] += 10             # make one prediction correct
    target = target.to(DEV)
    dist = (torch.rand(B, A, generator=g) * 1.2).to(DEV)
    dl = torch.empty(B, A, device=DEV)
    rb = torch.empty(B, 4, device=DEV); out = torch.zeros(2, device=DEV)
    pred = torch.empty(B, dtype=torch.int64, device=DEV)
    ops.gqa_prior_loss(z, target, dist, dlogits=dl, row_buf=rb, out=out, pred=pred, lam=10.0, gscale=0.25)
    loss, score, grad = _loss_ref(z, target, dist, 10.0)
    assert abs(float(out[0]) - float(loss)) / float(loss) < 1e-5
    assert abs(float(out[1]) - float(score)) < 1e-6
    assert _rel(dl, 0.25 * grad) < 1e-4
    assert torch.equal(pred, z.argmax(1))
    # validation form: CE only, summed score
    ops.gqa_prior_loss(z, target, None, row_buf=rb, out=out, score_mean=False)
    ce = torch.nn.functional.cross_entropy(z.double(), torch.argmax(target.long(), 1)) * A
    assert abs(float(out[0]) - float(ce)) / float(ce) < 1e-5


def test_adamw_and_clip_and_shadow():
    import ctypes
    from clg_vqa_b200 import _lib as L
    torch.manual_seed(0)
    shapes = [(768, 768), (768,), (1842,), (3, 7)]
    wds = [1e-2, 0.0, 1e-2, 0.0]
    lrs = [4e-5, 4e-5, 1e-4, 4e-5]
    starts, n = [], 0
    for s in shapes:
        starts.append(n)
        n += (math.prod(s) + L.CHUNK - 1) // L.CHUNK * L.CHUNK
    n += L.CHUNK      # one trailing chunk that belongs to no tensor
    p = torch.zeros(n, device=DEV); g = torch.zeros(n, device=DEV)
    m = torch.zeros(n, device=DEV); v = torch.zeros(n, device=DEV)
    chunk_tensor = torch.full((n // L.CHUNK,), -1, dtype=torch.int32)
    metas = (L.TensorMeta * len(shapes))()
    shadow = torch.zeros(n, device=DEV, dtype=torch.bfloat16)
    maskf = (torch.rand(768 * 768, device=DEV) > 0.41).float()
    bits = torch.zeros(n // 32, device=DEV, dtype=torch.int32)
    ops.pack_mask(maskf, bits[:768 * 768 // 32])
    refs = []
    for i, s in enumerate(shapes):
        k = math.prod(s)
        p[starts[i]:starts[i] + k] = torch.randn(k, device=DEV) * 0.05
        g[starts[i]:starts[i] + k] = torch.randn(k, device=DEV) * 3.0
        chunk_tensor[starts[i] // L.CHUNK:(starts[i] + k + L.CHUNK - 1) // L.CHUNK] = i
        metas[i] = L.TensorMeta(lrs[i], wds[i], starts[i] if i in (0, 3) else -1, 0 if i == 0 else -1, starts[i], k)
    p[starts[0]:starts[0] + 768 * 768] *= maskf          # pre-zeroed masked weights (train_task_sft.py:L432-453)
    g[starts[0]:starts[0] + 768 * 768] *= maskf
    for i, s in enumerate(shapes):
        k = math.prod(s)
        refs.append([p[starts[i]:starts[i] + k].double().clone(), torch.zeros(k, device=DEV, dtype=torch.float64),
                     torch.zeros(k, device=DEV, dtype=torch.float64)])
    chunk_tensor = chunk_tensor.to(DEV)
    tens = torch.frombuffer(bytearray(bytes(metas)), dtype=torch.uint8).to(DEV)
    partials = torch.zeros(1024, device=DEV); ss = torch.zeros(1, device=DEV)
    for step in range(1, 4):
        ss.zero_()
        ops.sumsq(g, partials, ss)
        total = float(g.double().pow(2).sum().sqrt())
        assert abs(math.sqrt(float(ss)) - total) / total < 1e-5
        ops.adamw_step(p, g, m, v, chunk_tensor, tens, sumsq=ss, max_norm=1.0, gscale=1.0, eps=1e-6,
                       bias_c1=1 - 0.9 ** step, bias_c2=1 - 0.999 ** step, shadow=shadow, mask_bits=bits)
        coef = min(1.0, 1.0 / (total + 1e-6))
        for i, s in enumerate(shapes):
            k = math.prod(s)
            pr, mr, vr = refs[i]
            gg = g[starts[i]:starts[i] + k].double() * coef
            mr.mul_(0.9).add_(gg, alpha=0.1)
            vr.mul_(0.999).addcmul_(gg, gg, value=0.001)
            ssz = lrs[i] * math.sqrt(1 - 0.999 ** step) / (1 - 0.9 ** step)
            pr.addcdiv_(mr, vr.sqrt() + 1e-6, value=-ssz)
            if wds[i] > 0:
                pr.add_(pr, alpha=-lrs[i] * wds[i])
            assert _rel(p[starts[i]:starts[i] + k], pr) < 1e-5
    w = p[starts[0]:starts[0] + 768 * 768]
    assert torch.equal(w[maskf == 0], torch.zeros_like(w[maskf == 0]))       # masked weights stay exactly zero
    assert torch.equal(shadow[starts[0]:starts[0] + 768 * 768], (w * maskf).bfloat16())
    assert torch.equal(shadow[starts[3]:starts[3] + 21], p[starts[3]:starts[3] + 21].bfloat16())
    assert float(p[n - L.CHUNK:].abs().max()) == 0.0


def test_utilities():
    x = _rand(1000, 2048, seed=80)
    xb = torch.empty(1000, 2048, device=DEV, dtype=torch.bfloat16)
    ops.cast_f32_to_bf16(x, xb)
    assert torch.equal(xb, x.bfloat16())
    out = torch.zeros(2304, device=DEV)
    xs = _rand(3000, 2304, dtype=torch.bfloat16, seed=81)
    ops.colsum(xs, out)
    assert _rel(out, xs.double().sum(0)) < 1e-5
    tm = (torch.rand(7, 40, device=DEV) > 0.3).long(); im = (torch.rand(7, 36, device=DEV) > 0.2).long()
    kb = torch.empty(7, 76, device=DEV); rv = torch.empty(7, 76, device=DEV, dtype=torch.uint8)
    ops.build_key_bias(tm, im, kb, rv)
    assert torch.equal(kb, (1.0 - torch.cat([tm, im], 1).float()) * -10000.0)
    ops.build_key_bias(tm, im, kb, rv, m3p_prefix=True)
    n = tm.sum(1) + im.sum(1)
    valid = torch.arange(76, device=DEV)[None] < n[:, None]
    assert torch.equal(rv.bool(), valid)
    assert torch.equal(kb == 0, valid) and bool(torch.isinf(kb[~valid]).all())
    w = _rand(768 * 768, seed=82)
    mask = (torch.rand(768 * 768, device=DEV) > 0.5).float()
    bits = torch.zeros(768 * 768 // 32, device=DEV, dtype=torch.int32)
    ops.pack_mask(mask, bits)
    w2 = w.clone()
    ops.apply_mask(w2, bits)
    assert torch.equal(w2, w * mask)


@pytest.mark.parametrize("S,B", [(76, 5), (140, 5), (128, 5), (33, 5), (76, 67), (96, 40), (16, 30), (140, 67), (129, 9)])
def test_attention_tc_matches_ffma(S, B):
    """tcgen05 attention (bf16; backend 0 = persistent pipelined kernels where the shape allows, 2 = one CTA per head)
    against the FFMA kernel on the same inputs, incl. identical dropout masks.  B = 67 / 40 / 30 give the persistent
    CTAs several items each (ring wrap-around, both softmax groups, odd item counts)."""
    H = 12
    D = H * 64
    qkv = _rand(B * S, 3 * D, dtype=torch.bfloat16, seed=90)
    nvalid = torch.tensor(([S, S - 3, max(1, S // 2), S, 7] * ((B + 4) // 5))[:B], device=DEV)
    valid = torch.arange(S, device=DEV)[None, :] < nvalid[:, None]
    kb = torch.where(valid, 0.0, -10000.0).float().contiguous()
    dctx = _rand(B * S, D, dtype=torch.bfloat16, seed=91)
    outs = {}
    for backend in (0, 1, 2):
        ops.set_attention_backend(backend)
        for site in (ops.NO_DROP, DropSite(0.1, 5, 1 << 41)):
            ctx = torch.zeros(B * S, D, device=DEV, dtype=torch.bfloat16)
            lse = torch.zeros(B, H, S, device=DEV)
            dqkv = torch.zeros_like(qkv)
            ops.attention_fwd(qkv, kb, ctx, lse, B, S, H, 0.125, site)
            ops.attention_bwd(qkv, kb, ctx, dctx, lse, dqkv, B, S, H, 0.125, site)
            outs[(backend, site.p)] = (ctx.float(), lse, dqkv.float())
    ops.set_attention_backend(0)
    for pdrop in (0.0, 0.1):
        c1, l1, d1 = outs[(1, pdrop)]
        for backend in (0, 2):
            c0, l0, d0 = outs[(backend, pdrop)]
            assert _rel(l0, l1) < 1e-5
            assert _rel(c0, c1) < 1.5e-2
            assert _rel(d0, d1) < 3e-2
        # the two tcgen05 schedules agree up to bf16 rounding flips
        assert _rel(outs[(0, pdrop)][0], outs[(2, pdrop)][0]) < 1e-2       # same operands; exp formulated in base 2
        assert _rel(outs[(0, pdrop)][2], outs[(2, pdrop)][2]) < 8e-3     # bf16 rounding flips of dS (D_i summed in a different order)
    assert torch.isfinite(outs[(0, 0.1)][2]).all()


@pytest.mark.parametrize("S,B", [(76, 37), (48, 9)])
def test_attention_packed_rows_match_padded(S, B):
    """Packed variable-length layout (clg_pack_index + cu_rows): dropping the masked positions changes nothing for the
    kept ones - forward context and all three gradients equal the padded computation row for row."""
    H, T = 12, 40
    R = S - T
    D = H * 64
    g = torch.Generator(device="cpu").manual_seed(7)
    tlen = torch.randint(3, T + 1, (B,), generator=g)
    rlen = torch.randint(1, R + 1, (B,), generator=g) if R > 0 else torch.zeros(B, dtype=torch.long)
    tm = (torch.arange(T)[None] < tlen[:, None]).long().to(DEV)
    im = (torch.arange(R)[None] < rlen[:, None]).long().to(DEV)
    tm[1, 0] = 0                                             # a masked first token is still kept as a query row
    kb = torch.empty(B, S, device=DEV)
    ops.build_key_bias(tm, im, kb)
    qkv = _rand(B * S, 3 * D, dtype=torch.bfloat16, seed=95)
    dctx = _rand(B * S, D, dtype=torch.bfloat16, seed=96)
    site = DropSite(0.1, 5, 1 << 41)
    # padded reference
    ctx = torch.zeros(B * S, D, device=DEV, dtype=torch.bfloat16); lse = torch.zeros(B, H, S, device=DEV)
    dqkv = torch.zeros_like(qkv)
    ops.attention_fwd(qkv, kb, ctx, lse, B, S, H, 0.125, site)
    ops.attention_bwd(qkv, kb, ctx, dctx, lse, dqkv, B, S, H, 0.125, site)
    # packed
    cu = torch.zeros(B + 1, device=DEV, dtype=torch.int32); src = torch.zeros(B * S, device=DEV, dtype=torch.int32)
    ops.pack_index(tm, im, cu, src)
    keep = torch.cat([tm, im], 1).bool(); keep[:, 0] = True
    Mp = int(cu[-1])
    assert Mp == int(keep.sum())
    assert torch.equal(src[:Mp].long(), torch.nonzero(keep.view(-1)).flatten())
    assert torch.equal(cu[1:].long(), keep.sum(1).cumsum(0))
    qp = torch.full_like(qkv, float("nan")); dcp = torch.full_like(dctx, float("nan"))
    kbp = torch.full((B * S, 1), float("nan"), device=DEV)
    ops.rows_gather(qkv, src, qp, n=Mp); ops.rows_gather(dctx, src, dcp, n=Mp); ops.rows_gather(kb.view(-1, 1), src, kbp, n=Mp)
    assert torch.equal(qp[:Mp], qkv[src[:Mp].long()])
    ctxp = torch.zeros_like(ctx); lsep = torch.zeros_like(lse); dqp = torch.zeros_like(qkv)
    ops.attention_fwd(qp, kbp, ctxp, lsep, B, S, H, 0.125, site, cu_rows=cu, total_rows=Mp)
    ops.attention_bwd(qp, kbp, ctxp, dcp, lsep, dqp, B, S, H, 0.125, site, cu_rows=cu, total_rows=Mp)
    sel = src[:Mp].long()
    # the dropout stream is indexed by (sample, head, packed row, packed column): compare without dropout instead
    ctx0 = torch.zeros_like(ctx); lse0 = torch.zeros_like(lse); dq0 = torch.zeros_like(qkv)
    ops.attention_fwd(qkv, kb, ctx0, lse0, B, S, H, 0.125)
    ops.attention_bwd(qkv, kb, ctx0, dctx, lse0, dq0, B, S, H, 0.125)
    ctxp0 = torch.zeros_like(ctx); lsep0 = torch.zeros_like(lse); dqp0 = torch.zeros_like(qkv)
    ops.attention_fwd(qp, kbp, ctxp0, lsep0, B, S, H, 0.125, cu_rows=cu, total_rows=Mp)
    ops.attention_bwd(qp, kbp, ctxp0, dcp, lsep0, dqp0, B, S, H, 0.125, cu_rows=cu, total_rows=Mp)
    assert _rel(ctxp0[:Mp], ctx0[sel]) < 1e-2
    assert bool(torch.isfinite(dqp0[:Mp].float()).all())
    # gradients of kept rows: the padded run also back-propagates through the (unused) outputs of masked query rows, so
    # feed it a dctx that is zero there
    dctx_m = torch.zeros_like(dctx); dctx_m[sel] = dctx[sel]
    ops.attention_bwd(qkv, kb, ctx0, dctx_m, lse0, dq0, B, S, H, 0.125)
    assert _rel(dqp0[:Mp], dq0[sel]) < 2e-2
    assert bool(torch.isfinite(ctxp[:Mp].float()).all()) and bool(torch.isfinite(dqp[:Mp].float()).all())
    # round trip of the row scatter
    back = torch.zeros_like(qkv)
    ops.rows_scatter(qp, src, back, n=Mp)
    assert torch.equal(back[sel], qkv[sel]) and float(back[~keep.view(-1)].abs().max()) == 0.0


@pytest.mark.parametrize("S,B,T", [(140, 37, 40), (160, 13, 40), (133, 50, 33)])
def test_attention_wide_packed_forward_matches_padded(S, B, T):
    """S in (128, 160] (X101 features / M3P): the pipelined forward reads the packed layout - samples with <= 128 kept rows
    have one query tile, longer ones two; context rows and log-sum-exp equal the padded computation row for row."""
    H = 12
    R = S - T
    D = H * 64
    g = torch.Generator(device="cpu").manual_seed(17)
    tlen = torch.randint(3, T + 1, (B,), generator=g)
    rlen = torch.randint(R - 30, R + 1, (B,), generator=g)
    tlen[0], rlen[0] = T, R                                  # a full-length sample (two query tiles)
    tlen[1], rlen[1] = 3, R - 30                             # a short one (one tile)
    tm = (torch.arange(T)[None] < tlen[:, None]).long().to(DEV)
    im = (torch.arange(R)[None] < rlen[:, None]).long().to(DEV)
    kb = torch.empty(B, S, device=DEV)
    ops.build_key_bias(tm, im, kb)
    qkv = _rand(B * S, 3 * D, dtype=torch.bfloat16, seed=195)
    ctx0 = torch.zeros(B * S, D, device=DEV, dtype=torch.bfloat16); lse0 = torch.zeros(B, H, S, device=DEV)
    ops.attention_fwd(qkv, kb, ctx0, lse0, B, S, H, 0.125)
    cu = torch.zeros(B + 1, device=DEV, dtype=torch.int32); src = torch.zeros(B * S, device=DEV, dtype=torch.int32)
    ops.pack_index(tm, im, cu, src)
    Mp = int(cu[-1])
    lens = (cu[1:] - cu[:-1]).cpu()
    assert int(lens.max()) > 128 and int(lens.min()) <= 128
    qp = torch.full_like(qkv, float("nan")); kbp = torch.full((B * S, 1), float("nan"), device=DEV)
    ops.rows_gather(qkv, src, qp, n=Mp); ops.rows_gather(kb.view(-1, 1), src, kbp, n=Mp)
    qp[Mp:] = 0                                              # rows past the packed end are read by the last K / V boxes
    ctxp = torch.full((B * S, D), float("nan"), device=DEV, dtype=torch.bfloat16); lsep = torch.zeros(B, H, S, device=DEV)
    ops.attention_fwd(qp, kbp, ctxp, lsep, B, S, H, 0.125, cu_rows=cu, total_rows=Mp)
    sel = src[:Mp].long()
    assert bool(torch.isfinite(ctxp[:Mp].float()).all())
    assert bool(torch.isnan(ctxp[Mp:].float()).all())        # nothing is written past the packed rows
    assert _rel(ctxp[:Mp], ctx0[sel]) < 1e-2
    # log-sum-exp of the kept query rows ([B, H, S] indexed by the position inside the packed sample)
    keep = torch.cat([tm, im], 1).bool(); keep[:, 0] = True
    for b in (0, 1, B - 1):
        pos = torch.nonzero(keep[b]).flatten()
        assert _rel(lsep[b, :, :pos.numel()], lse0[b][:, pos]) < 1e-5
    # with dropout: finite, and a different mask than without
    ctxd = torch.zeros_like(ctxp)
    ops.attention_fwd(qp, kbp, ctxd, lsep, B, S, H, 0.125, DropSite(0.1, 5, 1 << 41), cu_rows=cu, total_rows=Mp)
    assert bool(torch.isfinite(ctxd[:Mp].float()).all()) and float((ctxd[:Mp].float() - ctxp[:Mp].float()).abs().max()) > 0
