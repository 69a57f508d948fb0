// GEMM epilogue shared by the tcgen05 and the SIMT kernels.
//   v = alpha*acc ; +bias ; [dropout] ; act ; [dropout] ; +residual ; *mask ; store / accumulate
#pragma once
#include "common.cuh"

namespace clg {

struct EpiArgs {
  void* C; int64_t ldc; int c_dtype;
  int M, N;
  float alpha;
  int act;
  const float* bias;
  const float* residual; int64_t ld_res;
  const void* aux_in; int64_t ld_aux_in; int aux_in_dtype;
  void* aux_out; int64_t ld_aux_out; int aux_out_dtype;
  const uint32_t* mask_bits;
  int accumulate;     // 0: store, 1: C += v (plain RMW), 2: C += v (atomic; split-K)
  int drop_before_act;
  Drop drop;
  float* colsum;      // optional [N]: += column sums of the epilogue result
  // second row group (see clgGemmParams): its output, mask and row count
  void* C2; int64_t ldc2; const uint32_t* mask_bits2; int M2;
};

inline EpiArgs make_epi(const clgGemmParams& p) {
  EpiArgs e;
  e.C = p.C; e.ldc = p.ldc; e.c_dtype = p.c_dtype;
  e.M = p.M; e.N = p.N;
  e.alpha = p.alpha;
  e.act = p.act;
  e.bias = p.bias;
  e.residual = p.residual; e.ld_res = p.ld_res;
  e.aux_in = p.aux_in; e.ld_aux_in = p.ld_aux_in; e.aux_in_dtype = p.aux_in_dtype;
  e.aux_out = p.aux_out; e.ld_aux_out = p.ld_aux_out; e.aux_out_dtype = p.aux_out_dtype;
  e.mask_bits = p.mask_bits;
  e.accumulate = p.accumulate ? 1 : 0;
  e.drop_before_act = p.drop_before_act;
  e.drop = make_drop(p.dropout);
  e.colsum = p.colsum;
  e.C2 = p.C2; e.ldc2 = p.ldc2; e.mask_bits2 = p.mask_bits2; e.M2 = p.M2;
  return e;
}

// Column sums of a [32 lanes (rows)] x [32 columns] register tile: a reduce-scatter over the warp (16 + 8 + 4 + 2 + 1
// shuffles); lane l returns the sum of column l.  Destroys v.
__device__ __forceinline__ float warp_colsum32(float* v, int lane) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const bool upper = (lane & off) != 0;
#pragma unroll
    for (int k = 0; k < off; ++k) {
      const float keep = upper ? v[k + off] : v[k];
      const float send = upper ? v[k] : v[k + off];
      v[k] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  return v[0];
}

__device__ __forceinline__ float epi_scalar(const EpiArgs& e, int row, int col, float acc) {
  float v = acc * e.alpha;
  if (e.bias) v += __ldg(e.bias + col);
  const uint64_t idx = (uint64_t)row * (uint64_t)e.N + (uint64_t)col;
  if (e.drop.on && e.drop_before_act) v = drop_apply(e.drop, idx, v);
  switch (e.act) {
    case CLG_ACT_GELU:
      if (e.aux_out) store_from_f32(e.aux_out, e.aux_out_dtype, (int64_t)row * e.ld_aux_out + col, v);
      v = gelu_erf(v);
      break;
    case CLG_ACT_GELU_GRAD:
      if (e.aux_out) store_from_f32(e.aux_out, e.aux_out_dtype, (int64_t)row * e.ld_aux_out + col, dgelu_erf(v));
      v = gelu_erf(v);
      break;
    case CLG_ACT_MUL_AUX: v *= load_as_f32(e.aux_in, e.aux_in_dtype, (int64_t)row * e.ld_aux_in + col); break;
    case CLG_ACT_RELU: v = fmaxf(v, 0.f); break;
    case CLG_ACT_TANH: v = tanhf(v); break;
    case CLG_ACT_DGELU: v *= dgelu_erf(load_as_f32(e.aux_in, e.aux_in_dtype, (int64_t)row * e.ld_aux_in + col)); break;
    case CLG_ACT_DRELU: v = load_as_f32(e.aux_in, e.aux_in_dtype, (int64_t)row * e.ld_aux_in + col) > 0.f ? v : 0.f; break;
    case CLG_ACT_DTANH: {
      const float y = load_as_f32(e.aux_in, e.aux_in_dtype, (int64_t)row * e.ld_aux_in + col);
      v *= (1.f - y * y);
    } break;
    default: break;
  }
  if (e.drop.on && !e.drop_before_act) v = drop_apply(e.drop, idx, v);
  if (e.residual) v += __ldg(e.residual + (int64_t)row * e.ld_res + col);
  if (e.mask_bits) v *= mask_bit(e.mask_bits, (int64_t)idx) ? 1.f : 0.f;
  return v;
}

// Process W consecutive columns [col0, col0+W) of one row; `acc` holds the fp32 accumulators.
// Vectorised fast path when the whole chunk is in range, 16B-aligned and the epilogue is the
// plain linear / bias / GeLU / residual form; otherwise falls back to the scalar routine.
template <int W>
__device__ __forceinline__ void epi_row_chunk(const EpiArgs& e, int row, int col0, const float* acc) {
  static_assert(W % 8 == 0, "chunk width must be a multiple of 8");
  if (row >= e.M) return;
  const int ncols = min(W, e.N - col0);
  if (ncols <= 0) return;
  float v[W];
  const bool full = (ncols == W);
  const bool simple = full &&
                      (e.act == CLG_ACT_NONE || e.act == CLG_ACT_GELU || e.act == CLG_ACT_DGELU) &&
                      ((col0 & 7) == 0) && ((e.ldc & 7) == 0);
  if (simple) {
#pragma unroll
    for (int j = 0; j < W; ++j) v[j] = acc[j] * e.alpha;
    if (e.bias) {
      const float4* b4 = reinterpret_cast<const float4*>(e.bias + col0);
#pragma unroll
      for (int j = 0; j < W / 4; ++j) {
        const float4 b = __ldg(b4 + j);
        v[4 * j + 0] += b.x; v[4 * j + 1] += b.y; v[4 * j + 2] += b.z; v[4 * j + 3] += b.w;
      }
    }
    const uint64_t eidx = (uint64_t)row * (uint64_t)e.N + (uint64_t)col0;
    if (e.drop.on && e.drop_before_act) {
#pragma unroll
      for (int j = 0; j < W; ++j) v[j] = drop_keep(e.drop, eidx + j) ? v[j] * e.drop.scale : 0.f;
    }
    if (e.act == CLG_ACT_GELU) {
      if (e.aux_out) {
        if (e.aux_out_dtype == CLG_BF16 && (e.ld_aux_out & 7) == 0) {
          uint4* o = reinterpret_cast<uint4*>((bf16*)e.aux_out + (int64_t)row * e.ld_aux_out + col0);
#pragma unroll
          for (int j = 0; j < W / 8; ++j) {
            __nv_bfloat162 h0 = __floats2bfloat162_rn(v[8 * j + 0], v[8 * j + 1]);
            __nv_bfloat162 h1 = __floats2bfloat162_rn(v[8 * j + 2], v[8 * j + 3]);
            __nv_bfloat162 h2 = __floats2bfloat162_rn(v[8 * j + 4], v[8 * j + 5]);
            __nv_bfloat162 h3 = __floats2bfloat162_rn(v[8 * j + 6], v[8 * j + 7]);
            uint4 u;
            u.x = *reinterpret_cast<uint32_t*>(&h0); u.y = *reinterpret_cast<uint32_t*>(&h1);
            u.z = *reinterpret_cast<uint32_t*>(&h2); u.w = *reinterpret_cast<uint32_t*>(&h3);
            o[j] = u;
          }
        } else {
#pragma unroll
          for (int j = 0; j < W; ++j)
            store_from_f32(e.aux_out, e.aux_out_dtype, (int64_t)row * e.ld_aux_out + col0 + j, v[j]);
        }
      }
#pragma unroll
      for (int j = 0; j < W; ++j) v[j] = gelu_erf(v[j]);
    } else if (e.act == CLG_ACT_DGELU) {
      if (e.aux_in_dtype == CLG_BF16 && (e.ld_aux_in & 7) == 0) {
        const uint4* a = reinterpret_cast<const uint4*>((const bf16*)e.aux_in + (int64_t)row * e.ld_aux_in + col0);
#pragma unroll
        for (int j = 0; j < W / 8; ++j) {
          const uint4 u = __ldg(a + j);
          const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            const float2 f = __bfloat1622float2(h[t]);
            v[8 * j + 2 * t + 0] *= dgelu_erf(f.x);
            v[8 * j + 2 * t + 1] *= dgelu_erf(f.y);
          }
        }
      } else {
#pragma unroll
        for (int j = 0; j < W; ++j)
          v[j] *= dgelu_erf(load_as_f32(e.aux_in, e.aux_in_dtype, (int64_t)row * e.ld_aux_in + col0 + j));
      }
    }
    if (e.drop.on && !e.drop_before_act) {
#pragma unroll
      for (int j = 0; j < W; ++j) v[j] = drop_keep(e.drop, eidx + j) ? v[j] * e.drop.scale : 0.f;
    }
    if (e.residual) {
      if ((e.ld_res & 3) == 0) {
        const float4* r4 = reinterpret_cast<const float4*>(e.residual + (int64_t)row * e.ld_res + col0);
#pragma unroll
        for (int j = 0; j < W / 4; ++j) {
          const float4 r = __ldg(r4 + j);
          v[4 * j + 0] += r.x; v[4 * j + 1] += r.y; v[4 * j + 2] += r.z; v[4 * j + 3] += r.w;
        }
      } else {
#pragma unroll
        for (int j = 0; j < W; ++j) v[j] += __ldg(e.residual + (int64_t)row * e.ld_res + col0 + j);
      }
    }
    if (e.mask_bits) {
      if ((eidx & 31) == 0 && (W % 32) == 0) {
#pragma unroll
        for (int w = 0; w < W / 32; ++w) {
          const uint32_t word = __ldg(e.mask_bits + (eidx >> 5) + w);
#pragma unroll
          for (int j = 0; j < 32; ++j) v[32 * w + j] *= ((word >> j) & 1u) ? 1.f : 0.f;
        }
      } else {
#pragma unroll
        for (int j = 0; j < W; ++j) v[j] *= mask_bit(e.mask_bits, (int64_t)(eidx + j)) ? 1.f : 0.f;
      }
    }
    if (e.colsum) {
#pragma unroll
      for (int j = 0; j < W; ++j) atomicAdd(e.colsum + col0 + j, v[j]);
    }
    if (e.c_dtype == CLG_BF16) {
      if (e.accumulate) {
        bf16* c = (bf16*)e.C + (int64_t)row * e.ldc + col0;
#pragma unroll
        for (int j = 0; j < W; ++j) c[j] = __float2bfloat16_rn(__bfloat162float(c[j]) + v[j]);
      } else {
        uint4* o = reinterpret_cast<uint4*>((bf16*)e.C + (int64_t)row * e.ldc + col0);
#pragma unroll
        for (int j = 0; j < W / 8; ++j) {
          __nv_bfloat162 h0 = __floats2bfloat162_rn(v[8 * j + 0], v[8 * j + 1]);
          __nv_bfloat162 h1 = __floats2bfloat162_rn(v[8 * j + 2], v[8 * j + 3]);
          __nv_bfloat162 h2 = __floats2bfloat162_rn(v[8 * j + 4], v[8 * j + 5]);
          __nv_bfloat162 h3 = __floats2bfloat162_rn(v[8 * j + 6], v[8 * j + 7]);
          uint4 u;
          u.x = *reinterpret_cast<uint32_t*>(&h0); u.y = *reinterpret_cast<uint32_t*>(&h1);
          u.z = *reinterpret_cast<uint32_t*>(&h2); u.w = *reinterpret_cast<uint32_t*>(&h3);
          o[j] = u;
        }
      }
    } else {
      float* c = (float*)e.C + (int64_t)row * e.ldc + col0;
      if (e.accumulate == 2) {
#pragma unroll
        for (int j = 0; j < W; ++j) atomicAdd(c + j, v[j]);
      } else if (e.accumulate == 1) {
        float4* c4 = reinterpret_cast<float4*>(c);
#pragma unroll
        for (int j = 0; j < W / 4; ++j) {
          float4 o = c4[j];
          o.x += v[4 * j + 0]; o.y += v[4 * j + 1]; o.z += v[4 * j + 2]; o.w += v[4 * j + 3];
          c4[j] = o;
        }
      } else {
        float4* c4 = reinterpret_cast<float4*>(c);
#pragma unroll
        for (int j = 0; j < W / 4; ++j)
          c4[j] = make_float4(v[4 * j + 0], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
      }
    }
    return;
  }
  // general (scalar) path
#pragma unroll
  for (int j = 0; j < W; ++j) {
    if (j >= ncols) break;
    const int col = col0 + j;
    const float r = epi_scalar(e, row, col, acc[j]);
    if (e.colsum) atomicAdd(e.colsum + col, r);
    const int64_t ci = (int64_t)row * e.ldc + col;
    if (e.c_dtype == CLG_BF16) {
      bf16* c = (bf16*)e.C;
      c[ci] = __float2bfloat16_rn(e.accumulate ? __bfloat162float(c[ci]) + r : r);
    } else {
      float* c = (float*)e.C;
      if (e.accumulate == 2) atomicAdd(c + ci, r);
      else if (e.accumulate == 1) c[ci] += r;
      else c[ci] = r;
    }
  }
}

}  // namespace clg
