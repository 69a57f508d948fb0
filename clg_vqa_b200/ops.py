"""Thin torch-tensor wrappers over the C ABI (include/clgvqa.h).  Device memory, streams and
dtype bookkeeping only; every arithmetic step happens in libclgvqa.so.  All functions enqueue on
torch's current CUDA stream and raise `ClgError` on any failure (no fallback)."""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch

from . import _lib as L
from ._lib import (ACT_DGELU, ACT_DRELU, ACT_DTANH, ACT_GELU, ACT_GELU_GRAD, ACT_MUL_AUX, ACT_NONE, ACT_RELU, ACT_TANH,  # noqa: F401
                   CLG_BF16, CLG_F32, GEMM_AUTO, GEMM_SIMT, GEMM_TC, ClgError)

_launches = 0   # number of C-ABI kernel entry calls (bench.py reports it as gpu_launches)

# CLG_NVTX=1: NVTX ranges around the phases of the step (forward, every encoder layer, head, backward, gradient exchange,
# clip, AdamW) for Nsight Systems / `ncu --nvtx --nvtx-include`; off by default (a range costs ~1 us of host time)
NVTX = os.environ.get("CLG_NVTX", "0") == "1"


class nvtx_range:
    __slots__ = ("name",)

    def __init__(self, name: str):
        self.name = name

    def __enter__(self):
        if NVTX:
            torch.cuda.nvtx.range_push(self.name)

    def __exit__(self, *exc):
        if NVTX:
            torch.cuda.nvtx.range_pop()
        return False


def launch_count() -> int:
    return _launches


_gemm_timing = None   # when set to a list, tcgen05 GEMM launches append (start_event, end_event, flops)


def set_gemm_timing(rec):
    global _gemm_timing
    _gemm_timing = rec


def _count(n=1):
    global _launches
    _launches += n


def _stream() -> int:
    # raw handle of torch's current stream (torch.cuda.current_stream() builds a Stream object: 15 us per launch)
    return torch._C._cuda_getCurrentRawStream(torch.cuda.current_device())


# ----------------------------------------------------------------------------------------------------
# launch plans: the per-layer launch sequence of the step engine, recorded once and replayed from C
# ----------------------------------------------------------------------------------------------------
# A training step issues ~270 kernel launches; built one by one from Python each costs ~20 us of host time (ctypes struct
# marshalling, tensor-map encoding, launch), i.e. 6-7 ms per step - invisible at 256 samples per GPU (12 ms of GPU work), the
# bound at the reference's strong-scaling point of 32 per GPU.  The launches of an encoder layer are the same every step
# except for the packed row count and the dropout seed, so the engine records them once (begin_plan / end_plan: the
# wrappers below append their filled parameter struct instead of launching) and clg_run_plan replays the list natively,
# patching those two values.
OP_GEMM, OP_LN_FWD, OP_LN_BWD, OP_ATTN_FWD, OP_ATTN_BWD, OP_COLSUM = range(6)
_plan_rec = None


def begin_plan():
    global _plan_rec
    if _plan_rec is not None:
        raise ClgError("launch plans do not nest")
    _plan_rec = []


def end_plan():
    """-> plan object for run_plan (entries table + parameter blob, host memory owned by the returned object)"""
    global _plan_rec
    rec, _plan_rec = _plan_rec, None
    n = len(rec)
    entries = (L.PlanEntry * max(n, 1))()
    blob = bytearray()
    for i, (op, raw, rows_off, seed_off) in enumerate(rec):
        while len(blob) % 16:
            blob.append(0)
        ro = list(rows_off) + [-1] * (2 - len(rows_off))
        so = list(seed_off) + [-1] * (2 - len(seed_off))
        entries[i] = L.PlanEntry(op, len(blob), (C.c_int32 * 2)(*ro), (C.c_int32 * 2)(*so))
        blob += raw
    buf = (C.c_uint8 * max(len(blob), 16)).from_buffer_copy(bytes(blob) + bytes(max(0, 16 - len(blob))))
    return {"entries": entries, "n": n, "blob": buf}


def run_plan(plan, rows: int, seed: int):
    L.check(L.load().clg_run_plan(plan["entries"], plan["n"], plan["blob"], int(rows), C.c_uint64(seed & 0xFFFFFFFFFFFFFFFF), _stream()),
            "clg_run_plan")
    _count(plan["n"])


def _issue(op, name, p, rows_fields=(), seed_fields=()):
    """launch now, or - while a plan is being recorded - keep the filled struct for clg_run_plan"""
    rec = _plan_rec
    if rec is not None:
        cls = type(p)
        rows_off = [getattr(cls, f).offset for f in rows_fields]
        seed_off = [getattr(cls, f).offset + L.Dropout.seed.offset for f in seed_fields]
        rec.append((op, bytes(p), rows_off, seed_off))
        return
    L.check(getattr(L.load(), name)(C.byref(p), _stream()), name)
    _count()


def _dt(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return CLG_F32
    if t.dtype == torch.bfloat16:
        return CLG_BF16
    raise ClgError("unsupported dtype %s" % t.dtype)


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise ClgError("clg_vqa_b200 kernels need CUDA tensors (got a %s tensor); there is no CPU path" % t.device)


class DropSite:
    """(p, seed, offset) of one dropout site; p = 0 disables."""
    __slots__ = ("p", "seed", "offset")

    def __init__(self, p: float = 0.0, seed: int = 0, offset: int = 0):
        self.p, self.seed, self.offset = float(p), int(seed), int(offset)

    def c(self) -> L.Dropout:
        return L.Dropout(self.p, 0, self.seed & 0xFFFFFFFFFFFFFFFF, self.offset & 0xFFFFFFFFFFFFFFFF)


NO_DROP = DropSite()


def rowmap(group_rows=0, row_offset=0, group_stride=0) -> L.RowMap:
    return L.RowMap(int(group_rows), int(row_offset), int(group_stride))


def _ld(t: torch.Tensor) -> int:
    """leading dimension (row pitch in elements) of a 2-D row-major view"""
    if t.dim() != 2 or t.stride(1) != 1:
        raise ClgError("expected a 2-D tensor with unit inner stride, got shape %s stride %s" % (tuple(t.shape), t.stride()))
    return t.stride(0) if t.shape[0] > 1 else max(t.stride(0), t.shape[1])


def gemm(a, b, c, *, a_kmajor=True, b_kmajor=True, M=None, N=None, K=None, alpha=1.0, act=ACT_NONE, bias=None,
         residual=None, aux_in=None, aux_out=None, mask_bits=None, accumulate=False, dropout: DropSite = NO_DROP,
         drop_before_act=False, backend=GEMM_AUTO, split_k=0, colsum=None, group2=None):
    """C[M,N] = epi(alpha * A(m,k) B(n,k)).  a: [M,K] if a_kmajor else [K,M]; b: [N,K] if b_kmajor else [K,N]."""
    _need_cuda(a, b, c)
    M_given, K_given = M is not None, K is not None          # explicit = the packed row count (patched when a plan is replayed)
    if M is None:
        M = a.shape[0] if a_kmajor else a.shape[1]
    if K is None:
        K = a.shape[1] if a_kmajor else a.shape[0]
    if N is None:
        N = b.shape[0] if b_kmajor else b.shape[1]
    p = L.GemmParams()
    p.A, p.B, p.C = a.data_ptr(), b.data_ptr(), c.data_ptr()
    p.a_dtype, p.b_dtype, p.c_dtype = _dt(a), _dt(b), _dt(c)
    p.M, p.N, p.K = int(M), int(N), int(K)
    p.lda, p.ldb, p.ldc = _ld(a), _ld(b), _ld(c)
    p.a_kmajor, p.b_kmajor = int(a_kmajor), int(b_kmajor)
    p.alpha, p.act = float(alpha), int(act)
    p.bias = _ptr(bias)
    if residual is not None:
        p.residual, p.ld_res = residual.data_ptr(), _ld(residual)
    if aux_in is not None:
        p.aux_in, p.ld_aux_in, p.aux_in_dtype = aux_in.data_ptr(), _ld(aux_in), _dt(aux_in)
    if aux_out is not None:
        p.aux_out, p.ld_aux_out, p.aux_out_dtype = aux_out.data_ptr(), _ld(aux_out), _dt(aux_out)
    p.mask_bits = _ptr(mask_bits)
    p.accumulate = int(bool(accumulate))
    p.drop_before_act = int(bool(drop_before_act))
    p.dropout = dropout.c()
    p.backend, p.split_k = int(backend), int(split_k)
    p.colsum = _ptr(colsum)
    if group2 is not None:          # (a2, b2, c2, mask_bits2): a second GEMM of the same N, K and options stacked under this one
        a2, b2, c2, mb2 = group2
        _need_cuda(a2, b2, c2)
        p.A2, p.B2, p.C2 = a2.data_ptr(), b2.data_ptr(), c2.data_ptr()
        p.lda2, p.ldb2, p.ldc2 = _ld(a2), _ld(b2), _ld(c2)
        p.mask_bits2 = _ptr(mb2)
        p.M2 = int(a2.shape[0] if a_kmajor else a2.shape[1])
    if _plan_rec is not None:
        _issue(OP_GEMM, "clg_gemm", p, [f for f, v in (("M", M_given), ("K", K_given)) if v], ["dropout"] if dropout.p > 0 else [])
        return c
    rec = _gemm_timing
    if rec is not None:
        # same predicate as clg_gemm's auto dispatch: these launches run gemm_tc_kernel
        is_tc = (backend != GEMM_SIMT and p.a_dtype == CLG_BF16 and p.b_dtype == CLG_BF16 and
                 (backend == GEMM_TC or (M >= 128 and N >= 64 and K >= 64)) and p.lda % 8 == 0 and p.ldb % 8 == 0)
        if is_tc:
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            L.check(L.load().clg_gemm(C.byref(p), _stream()), "clg_gemm")
            e1.record()
            rec.append((e0, e1, 2.0 * (M + int(p.M2)) * N * K))
            _count()
            return c
    L.check(L.load().clg_gemm(C.byref(p), _stream()), "clg_gemm")
    _count()
    return c


def layernorm_fwd(x, gamma, beta, eps, *, y=None, y_lowp=None, mean=None, rstd=None, rows=None, out_map=None,
                  add_mat=None, add_map=None, add_vec=None, row_valid=None, valid_pre=False, valid_post=False,
                  dropout: DropSite = NO_DROP):
    _need_cuda(x, gamma, beta)
    p = L.LayerNormFwdParams()
    p.x, p.ldx = x.data_ptr(), _ld(x)
    p.gamma, p.beta = gamma.data_ptr(), beta.data_ptr()
    if y is not None:
        p.y, p.ldy = y.data_ptr(), _ld(y)
    if y_lowp is not None:
        p.y_lowp, p.ldy_lowp, p.lowp_dtype = y_lowp.data_ptr(), _ld(y_lowp), _dt(y_lowp)
    p.mean, p.rstd = _ptr(mean), _ptr(rstd)
    p.rows = int(x.shape[0] if rows is None else rows)
    p.N = int(x.shape[1])
    p.eps = float(eps)
    if out_map is not None:
        p.out_map = out_map
    if add_mat is not None:
        p.add_mat, p.ld_add = add_mat.data_ptr(), _ld(add_mat)
        if add_map is not None:
            p.add_map = add_map
    p.add_vec = _ptr(add_vec)
    p.row_valid = _ptr(row_valid)
    p.valid_pre, p.valid_post = int(valid_pre), int(valid_post)
    p.dropout = dropout.c()
    _issue(OP_LN_FWD, "clg_layernorm_fwd", p, ["rows"] if rows is not None else [], ["dropout"] if dropout.p > 0 else [])


def layernorm_bwd(dy, x, gamma, mean, rstd, *, dx=None, dx_lowp=None, dgamma=None, dbeta=None, dbias=None,
                  rows=None, dy_map=None, row_valid=None, valid_pre=False, valid_post=False,
                  dropout_out: DropSite = NO_DROP, dropout_in: DropSite = NO_DROP):
    _need_cuda(dy, x, gamma)
    p = L.LayerNormBwdParams()
    p.dy, p.lddy = dy.data_ptr(), _ld(dy)
    if dy_map is not None:
        p.dy_map = dy_map
    p.x, p.ldx = x.data_ptr(), _ld(x)
    p.gamma, p.mean, p.rstd = gamma.data_ptr(), mean.data_ptr(), rstd.data_ptr()
    if dx is not None:
        p.dx, p.lddx = dx.data_ptr(), _ld(dx)
    if dx_lowp is not None:
        p.dx_lowp, p.lddx_lowp, p.lowp_dtype = dx_lowp.data_ptr(), _ld(dx_lowp), _dt(dx_lowp)
    p.dgamma, p.dbeta, p.dbias = _ptr(dgamma), _ptr(dbeta), _ptr(dbias)
    p.rows = int(x.shape[0] if rows is None else rows)
    p.N = int(x.shape[1])
    p.row_valid = _ptr(row_valid)
    p.valid_pre, p.valid_post = int(valid_pre), int(valid_post)
    p.dropout_out, p.dropout_in = dropout_out.c(), dropout_in.c()
    _issue(OP_LN_BWD, "clg_layernorm_bwd", p, ["rows"] if rows is not None else [],
           [f for f, d in (("dropout_out", dropout_out), ("dropout_in", dropout_in)) if d.p > 0])


def set_gemm_sm_limit(n: int):
    """SMs the persistent GEMM grids occupy (0 = all); see clg_set_gemm_sm_limit"""
    L.check(L.load().clg_set_gemm_sm_limit(int(n)), "clg_set_gemm_sm_limit")


def set_attention_backend(backend: int):
    """0 = auto (tcgen05 for bf16), 1 = force the FFMA kernel (debug / cross-check)"""
    L.check(L.load().clg_set_attention_backend(int(backend)), "clg_set_attention_backend")


def attention_fwd(qkv, key_bias, ctx, lse, B, S, H, scale, dropout: DropSite = NO_DROP, cu_rows=None, total_rows=0):
    _need_cuda(qkv, key_bias, ctx)
    p = L.AttentionFwdParams()
    p.qkv, p.ld_qkv = qkv.data_ptr(), _ld(qkv)
    p.key_bias = key_bias.data_ptr()
    p.ctx, p.ld_ctx = ctx.data_ptr(), _ld(ctx)
    p.lse = _ptr(lse)
    p.dtype = _dt(qkv)
    p.B, p.S, p.H = int(B), int(S), int(H)
    p.scale = float(scale)
    p.dropout = dropout.c()
    p.cu_rows, p.total_rows = _ptr(cu_rows), int(total_rows)
    _issue(OP_ATTN_FWD, "clg_attention_fwd", p, ["total_rows"] if cu_rows is not None else [], ["dropout"] if dropout.p > 0 else [])


def attention_bwd(qkv, key_bias, ctx, dctx, lse, dqkv, B, S, H, scale, dropout: DropSite = NO_DROP, cu_rows=None,
                  total_rows=0):
    _need_cuda(qkv, key_bias, ctx, dctx, lse, dqkv)
    p = L.AttentionBwdParams()
    p.qkv, p.ld_qkv = qkv.data_ptr(), _ld(qkv)
    p.key_bias = key_bias.data_ptr()
    p.ctx, p.ld_ctx = ctx.data_ptr(), _ld(ctx)
    p.dctx, p.ld_dctx = dctx.data_ptr(), _ld(dctx)
    p.lse = lse.data_ptr()
    p.dqkv, p.ld_dqkv = dqkv.data_ptr(), _ld(dqkv)
    p.dtype = _dt(qkv)
    p.B, p.S, p.H = int(B), int(S), int(H)
    p.scale = float(scale)
    p.dropout = dropout.c()
    p.cu_rows, p.total_rows = _ptr(cu_rows), int(total_rows)
    _issue(OP_ATTN_BWD, "clg_attention_bwd", p, ["total_rows"] if cu_rows is not None else [], ["dropout"] if dropout.p > 0 else [])


def embed_text_fwd(ids, seg, word, pos, type_tab, out, *, pad_id, out_map=None, pos_mode=0, pos_base=0):
    _need_cuda(ids, word, out)
    p = L.EmbedTextParams()
    p.ids, p.seg = ids.data_ptr(), _ptr(seg)
    p.word, p.pos, p.type = word.data_ptr(), _ptr(pos), _ptr(type_tab)
    p.out, p.ld_out = out.data_ptr(), _ld(out)
    if out_map is not None:
        p.out_map = out_map
    p.B, p.T, p.N, p.pad_id = int(ids.shape[0]), int(ids.shape[1]), int(word.shape[1]), int(pad_id)
    p.pos_mode, p.pos_base, p.vocab = int(pos_mode), int(pos_base), int(word.shape[0])
    L.check(L.load().clg_embed_text_fwd(C.byref(p), _stream()), "clg_embed_text_fwd")
    _count()


def embed_text_bwd(ids, seg, dout, dword, dpos, dtype_tab, *, pad_id, dout_map=None, pos_mode=0, pos_base=0, N=None,
                   type_row0_has_total=False):
    _need_cuda(ids, dout)
    p = L.EmbedTextBwdParams()
    p.ids, p.seg = ids.data_ptr(), _ptr(seg)
    p.dout, p.ld_dout = dout.data_ptr(), _ld(dout)
    if dout_map is not None:
        p.dout_map = dout_map
    p.dword, p.dpos, p.dtype_tab = _ptr(dword), _ptr(dpos), _ptr(dtype_tab)
    p.B, p.T, p.pad_id = int(ids.shape[0]), int(ids.shape[1]), int(pad_id)
    p.N = int(dout.shape[1] if N is None else N)
    p.pos_mode, p.pos_base = int(pos_mode), int(pos_base)
    p.vocab = int(dword.shape[0]) if dword is not None else 0
    p.type_row0_has_total = int(bool(type_row0_has_total))
    L.check(L.load().clg_embed_text_bwd(C.byref(p), _stream()), "clg_embed_text_bwd")
    _count()


def segment_sum_rows(rows, order, sorted_ids, table, pad_id, ids_out=None):
    """runs of equal sorted_ids summed in list order (see clg_segment_sum_rows): into table (ids_out None) or, as a local
    pre-reduction, into the run's first row with the other members' ids_out set to pad_id"""
    _need_cuda(rows, order, sorted_ids)
    L.check(L.load().clg_segment_sum_rows(rows.data_ptr(), _ld(rows), order.data_ptr(), sorted_ids.data_ptr(),
                                          int(sorted_ids.numel()), _ptr(table), int(rows.shape[1]), int(pad_id),
                                          0 if ids_out is None else 1, _ptr(ids_out), _stream()), "clg_segment_sum_rows")
    _count()


def gqa_prior_loss(logits, target, dist, *, dlogits=None, row_buf, out, pred=None, topk=10, lam=10.0,
                   gscale=1.0, score_mean=True):
    _need_cuda(logits, target, row_buf, out)
    p = L.GqaLossParams()
    p.logits, p.ld_logits = logits.data_ptr(), _ld(logits)
    p.target, p.ld_target = target.data_ptr(), _ld(target)
    if dist is not None:
        p.dist, p.ld_dist = dist.data_ptr(), _ld(dist)
    if dlogits is not None:
        p.dlogits, p.ld_dlogits = dlogits.data_ptr(), _ld(dlogits)
    p.row_buf, p.out, p.pred = row_buf.data_ptr(), out.data_ptr(), _ptr(pred)
    p.B, p.A, p.topk = int(logits.shape[0]), int(logits.shape[1]), int(topk)
    p.lambda_, p.gscale, p.score_mean = float(lam), float(gscale), int(bool(score_mean))
    L.check(L.load().clg_gqa_prior_loss(C.byref(p), _stream()), "clg_gqa_prior_loss")
    _count(2)


def sumsq(g_flat, partials, out, hot=None, sparse=None):
    """sparse = (first block, end block, touched map): blocks of that range are read only where marked (clg_sumsq_sparse)"""
    _need_cuda(g_flat, partials, out)
    if sparse is None:
        L.check(L.load().clg_sumsq(g_flat.data_ptr(), g_flat.numel(), partials.data_ptr(), partials.numel(),
                                   out.data_ptr(), _ptr(hot), _stream()), "clg_sumsq")
    else:
        lo, hi, touched = sparse
        L.check(L.load().clg_sumsq_sparse(g_flat.data_ptr(), g_flat.numel(), partials.data_ptr(), partials.numel(),
                                          out.data_ptr(), _ptr(hot), int(lo), int(hi), touched.data_ptr(), _stream()),
                "clg_sumsq_sparse")
    _count(2)


def mark_rows(ids, row_elems, first_elem, num_rows, skip_id, touched):
    """touched[block] = 1 for the 256-element arena blocks under the gradient rows of `ids` (int64)"""
    _need_cuda(ids, touched)
    ids = ids.reshape(-1)
    if ids.dtype != torch.int64 or not ids.is_contiguous():
        ids = ids.to(torch.int64).contiguous()
    L.check(L.load().clg_mark_rows(ids.data_ptr(), ids.numel(), int(row_elems), int(first_elem), int(num_rows), int(skip_id),
                                   touched.data_ptr(), _stream()), "clg_mark_rows")
    _count(1)


def adamw_step(p_flat, g_flat, m_flat, v_flat, chunk_tensor, tensors_dev, *, sumsq=None, max_norm=1.0, gscale=1.0,
               beta1=0.9, beta2=0.999, eps=1e-6, bias_c1=1.0, bias_c2=1.0, shadow=None, mask_bits=None, hot=None, zero_grad=False):
    _need_cuda(p_flat, g_flat, m_flat, v_flat, chunk_tensor, tensors_dev)
    p = L.AdamWParams()
    p.p, p.g, p.m, p.v = p_flat.data_ptr(), g_flat.data_ptr(), m_flat.data_ptr(), v_flat.data_ptr()
    p.n = p_flat.numel()
    p.chunk_tensor, p.tensors = chunk_tensor.data_ptr(), tensors_dev.data_ptr()
    p.sumsq = _ptr(sumsq)
    p.max_norm, p.gscale = float(max_norm), float(gscale)
    p.beta1, p.beta2, p.eps = float(beta1), float(beta2), float(eps)
    p.bias_c1, p.bias_c2 = float(bias_c1), float(bias_c2)
    p.shadow, p.mask_bits = _ptr(shadow), _ptr(mask_bits)
    p.hot = _ptr(hot)
    p.zero_grad = int(bool(zero_grad))
    L.check(L.load().clg_adamw_step(C.byref(p), _stream()), "clg_adamw_step")
    _count()


def cast_shadow(p_flat, chunk_tensor, tensors_dev, shadow, mask_bits=None):
    _need_cuda(p_flat, chunk_tensor, tensors_dev, shadow)
    L.check(L.load().clg_cast_shadow(p_flat.data_ptr(), p_flat.numel(), chunk_tensor.data_ptr(), tensors_dev.data_ptr(),
                                     shadow.data_ptr(), _ptr(mask_bits), _stream()), "clg_cast_shadow")
    _count()


def cast_f32_to_bf16(src, dst):
    _need_cuda(src, dst)
    L.check(L.load().clg_cast_f32_to_bf16(src.data_ptr(), dst.data_ptr(), src.numel(), _stream()), "clg_cast_f32_to_bf16")
    _count()


def cast_f32_to_bf16_2d(src, dst):
    """dst[:, :cols] = bf16(src); dst's padding columns (row pitch > cols) are zeroed"""
    _need_cuda(src, dst)
    rows, cols = src.shape
    L.check(L.load().clg_cast_f32_to_bf16_2d(src.data_ptr(), _ld(src), dst.data_ptr(), _ld(dst), int(rows), int(cols), _stream()),
            "clg_cast_f32_to_bf16_2d")
    _count()


def pack_index(txt_mask, img_mask, cu_rows, src_rows):
    """kept-row tables of the packed joint sequence (see clg_pack_index); cu_rows [B+1] int32, src_rows [B*(T+R)] int32"""
    _need_cuda(txt_mask, img_mask, cu_rows, src_rows)
    B, T = txt_mask.shape
    L.check(L.load().clg_pack_index(txt_mask.data_ptr(), img_mask.data_ptr(), int(B), int(T), int(img_mask.shape[1]),
                                    cu_rows.data_ptr(), src_rows.data_ptr(), _stream()), "clg_pack_index")
    _count()


def rows_gather(src, idx, dst, n=None):
    """dst[i] = src[idx[i]] for i < n (2-D row-major views, same dtype / row length)"""
    _need_cuda(src, idx, dst)
    L.check(L.load().clg_rows_gather(src.data_ptr(), _ld(src), idx.data_ptr(), int(idx.numel() if n is None else n),
                                     dst.data_ptr(), _ld(dst), int(src.shape[1]), _dt(src), _stream()), "clg_rows_gather")
    _count()


def rows_scatter(src, idx, dst, n=None):
    """dst[idx[i]] = src[i] for i < n"""
    _need_cuda(src, idx, dst)
    L.check(L.load().clg_rows_scatter(src.data_ptr(), _ld(src), idx.data_ptr(), int(idx.numel() if n is None else n),
                                      dst.data_ptr(), _ld(dst), int(src.shape[1]), _dt(src), _stream()), "clg_rows_scatter")
    _count()


def pack_mask(mask, bits, n=None):
    _need_cuda(mask, bits)
    L.check(L.load().clg_pack_mask(mask.data_ptr(), bits.data_ptr(), mask.numel() if n is None else n, _stream()), "clg_pack_mask")
    _count()


def apply_mask(w, bits, n=None):
    _need_cuda(w, bits)
    L.check(L.load().clg_apply_mask(w.data_ptr(), bits.data_ptr(), w.numel() if n is None else n, _stream()), "clg_apply_mask")
    _count()


def colsum(x, out, M=None, N=None):
    _need_cuda(x, out)
    if _plan_rec is not None:
        p = L.ColsumParams(x.data_ptr(), out.data_ptr(), _ld(x), _dt(x), int(x.shape[0] if M is None else M),
                           int(x.shape[1] if N is None else N), 0)
        _issue(OP_COLSUM, "clg_colsum", p, ["M"] if M is not None else [], [])
        return
    L.check(L.load().clg_colsum(x.data_ptr(), _dt(x), _ld(x), int(x.shape[0] if M is None else M),
                                int(x.shape[1] if N is None else N), out.data_ptr(), _stream()), "clg_colsum")
    _count()


def build_key_bias(txt_mask, img_mask, key_bias, row_valid=None, m3p_prefix=False):
    _need_cuda(txt_mask, img_mask, key_bias)
    B, T = txt_mask.shape
    R = img_mask.shape[1]
    L.check(L.load().clg_build_key_bias(txt_mask.data_ptr(), img_mask.data_ptr(), B, T, R, int(m3p_prefix),
                                        key_bias.data_ptr(), _ptr(row_valid), _stream()), "clg_build_key_bias")
    _count()


def dropout_inplace(x, site: DropSite):
    _need_cuda(x)
    if site.p <= 0:
        return
    d = site.c()
    L.check(L.load().clg_dropout_inplace(x.data_ptr(), x.numel(), C.byref(d), _stream()), "clg_dropout_inplace")
    _count()


def act_bwd(dy, aux, out, act):
    _need_cuda(dy, aux, out)
    L.check(L.load().clg_act_bwd(dy.data_ptr(), aux.data_ptr(), out.data_ptr(), dy.numel(), int(act), _stream()), "clg_act_bwd")
    _count()


def device_check(device: int = 0):
    L.check(L.load().clg_device_check(device), "clg_device_check")
