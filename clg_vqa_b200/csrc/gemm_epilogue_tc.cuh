// Warp-collective epilogue of the tcgen05 GEMM.  A warp owns 32 accumulator rows (its TMEM lane
// quarter; lane = row) and processes 32 columns at a time.  Every global tensor the epilogue touches
// (residual, saved pre-activation, output, pre-activation output, accumulate target) goes through a
// per-warp [32][33] fp32 staging tile in shared memory so that global loads / stores are issued
// row-coalesced (a warp instruction covers 4 x 128 B rows for fp32, 8 x 64 B rows for bf16) instead of
// 32 different rows per instruction.
#pragma once
#include "gemm_epilogue.cuh"

namespace clg {

static constexpr int STG_LD = 33;                   // floats per staging row (conflict-free row / column access)
static constexpr int STG_FLOATS = 32 * STG_LD;

// global [32 rows x 32 cols] (row r0.., col c0..) -> staging tile -> v[32] of this lane's row
template <typename T>
__device__ __forceinline__ void stage_load(float* stg, const T* g, int64_t ld, int r0, int c0, int M, int lane, float* v) {
  constexpr int EPV = 16 / sizeof(T);               // elements per 16 B
  constexpr int LPR = 32 / EPV;                     // lanes per row
  constexpr int RPI = 32 / LPR;                     // rows per instruction
  const int lr = lane / LPR, lc = (lane % LPR) * EPV;
#pragma unroll
  for (int it = 0; it < 32 / RPI; ++it) {
    const int r = it * RPI + lr;
    float f[EPV];
    if (r0 + r < M) {
      const uint4 u = __ldg(reinterpret_cast<const uint4*>(g + (int64_t)(r0 + r) * ld + c0 + lc));
      if (sizeof(T) == 4) {
        const float* pf = reinterpret_cast<const float*>(&u);
#pragma unroll
        for (int j = 0; j < EPV; ++j) f[j] = pf[j];
      } else {
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
        for (int j = 0; j < EPV / 2; ++j) { const float2 t = __bfloat1622float2(h[j]); f[2 * j] = t.x; f[2 * j + 1] = t.y; }
      }
    } else {
#pragma unroll
      for (int j = 0; j < EPV; ++j) f[j] = 0.f;
    }
#pragma unroll
    for (int j = 0; j < EPV; ++j) stg[r * STG_LD + lc + j] = f[j];
  }
  __syncwarp();
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = stg[lane * STG_LD + j];
  __syncwarp();
}

// v[32] of this lane's row -> staging tile -> global, optionally as atomic adds (split-K) or += (accumulate)
template <typename T>
__device__ __forceinline__ void stage_store(float* stg, T* g, int64_t ld, int r0, int c0, int M, int lane, const float* v,
                                            int accumulate) {
  constexpr int EPV = 16 / sizeof(T);
  constexpr int LPR = 32 / EPV;
  constexpr int RPI = 32 / LPR;
#pragma unroll
  for (int j = 0; j < 32; ++j) stg[lane * STG_LD + j] = v[j];
  __syncwarp();
  const int lr = lane / LPR, lc = (lane % LPR) * EPV;
#pragma unroll
  for (int it = 0; it < 32 / RPI; ++it) {
    const int r = it * RPI + lr;
    if (r0 + r >= M) continue;
    float f[EPV];
#pragma unroll
    for (int j = 0; j < EPV; ++j) f[j] = stg[r * STG_LD + lc + j];
    T* p = g + (int64_t)(r0 + r) * ld + c0 + lc;
    if (sizeof(T) == 4) {
      float* pf = reinterpret_cast<float*>(p);
      if (accumulate == 2) {
        atomicAdd(reinterpret_cast<float4*>(pf), make_float4(f[0], f[1], f[2], f[3]));
      } else {
        float4 o = make_float4(f[0], f[1], f[2], f[3]);
        if (accumulate == 1) { const float4 c = *reinterpret_cast<const float4*>(pf); o.x += c.x; o.y += c.y; o.z += c.z; o.w += c.w; }
        *reinterpret_cast<float4*>(pf) = o;
      }
    } else {
      bf16* pb = reinterpret_cast<bf16*>(p);
      if (accumulate) {
        const uint4 c = *reinterpret_cast<const uint4*>(pb);
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&c);
#pragma unroll
        for (int j = 0; j < EPV / 2; ++j) { const float2 t = __bfloat1622float2(h[j]); f[2 * j] += t.x; f[2 * j + 1] += t.y; }
      }
      uint4 u;
      uint32_t* pu = reinterpret_cast<uint32_t*>(&u);
#pragma unroll
      for (int j = 0; j < EPV / 2; ++j) {
        __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * j], f[2 * j + 1]);
        pu[j] = *reinterpret_cast<uint32_t*>(&h);
      }
      *reinterpret_cast<uint4*>(pb) = u;
    }
  }
  __syncwarp();
}

// bf16 tensors are staged as packed bf16x2 words ([32][17] uint32): half the shared-memory traffic, and the
// cooperative phase moves 16 B per lane without any conversion.
static constexpr int STG_LDW = 17;

__device__ __forceinline__ void stage_store_bf16(float* stg_f, bf16* g, int64_t ld, int r0, int c0, int M, int lane,
                                                 const float* v, int accumulate) {
  uint32_t* stg = reinterpret_cast<uint32_t*>(stg_f);
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
    stg[lane * STG_LDW + j] = *reinterpret_cast<uint32_t*>(&h);
  }
  __syncwarp();
  const int lr = lane >> 2, lw = (lane & 3) * 4;
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const int r = it * 8 + lr;
    if (r0 + r < M) {
      uint4 u;
      u.x = stg[r * STG_LDW + lw + 0]; u.y = stg[r * STG_LDW + lw + 1];
      u.z = stg[r * STG_LDW + lw + 2]; u.w = stg[r * STG_LDW + lw + 3];
      bf16* p = g + (int64_t)(r0 + r) * ld + c0 + lw * 2;
      if (accumulate) {
        const uint4 c = *reinterpret_cast<const uint4*>(p);
        const __nv_bfloat162* hc = reinterpret_cast<const __nv_bfloat162*>(&c);
        const __nv_bfloat162* hu = reinterpret_cast<const __nv_bfloat162*>(&u);
        __nv_bfloat162 o0 = __floats2bfloat162_rn(__low2float(hc[0]) + __low2float(hu[0]), __high2float(hc[0]) + __high2float(hu[0]));
        __nv_bfloat162 o1 = __floats2bfloat162_rn(__low2float(hc[1]) + __low2float(hu[1]), __high2float(hc[1]) + __high2float(hu[1]));
        __nv_bfloat162 o2 = __floats2bfloat162_rn(__low2float(hc[2]) + __low2float(hu[2]), __high2float(hc[2]) + __high2float(hu[2]));
        __nv_bfloat162 o3 = __floats2bfloat162_rn(__low2float(hc[3]) + __low2float(hu[3]), __high2float(hc[3]) + __high2float(hu[3]));
        u.x = *reinterpret_cast<uint32_t*>(&o0); u.y = *reinterpret_cast<uint32_t*>(&o1);
        u.z = *reinterpret_cast<uint32_t*>(&o2); u.w = *reinterpret_cast<uint32_t*>(&o3);
      }
      *reinterpret_cast<uint4*>(p) = u;
    }
  }
  __syncwarp();
}

__device__ __forceinline__ void stage_load_bf16(float* stg_f, const bf16* g, int64_t ld, int r0, int c0, int M, int lane, float* v) {
  uint32_t* stg = reinterpret_cast<uint32_t*>(stg_f);
  const int lr = lane >> 2, lw = (lane & 3) * 4;
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const int r = it * 8 + lr;
    uint4 u = make_uint4(0, 0, 0, 0);
    if (r0 + r < M) u = __ldg(reinterpret_cast<const uint4*>(g + (int64_t)(r0 + r) * ld + c0 + lw * 2));
    stg[r * STG_LDW + lw + 0] = u.x; stg[r * STG_LDW + lw + 1] = u.y;
    stg[r * STG_LDW + lw + 2] = u.z; stg[r * STG_LDW + lw + 3] = u.w;
  }
  __syncwarp();
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const uint32_t w = stg[lane * STG_LDW + j];
    v[2 * j] = __uint_as_float(w << 16);
    v[2 * j + 1] = __uint_as_float(w & 0xffff0000u);
  }
  __syncwarp();
}

__device__ __forceinline__ void stage_load_any(float* stg, const void* g, int dtype, int64_t ld, int r0, int c0, int M, int lane,
                                               float* v) {
  if (dtype == CLG_BF16) stage_load_bf16(stg, (const bf16*)g, ld, r0, c0, M, lane, v);
  else stage_load<float>(stg, (const float*)g, ld, r0, c0, M, lane, v);
}
__device__ __forceinline__ void stage_store_any(float* stg, void* g, int dtype, int64_t ld, int r0, int c0, int M, int lane,
                                                const float* v, int accumulate) {
  if (dtype == CLG_BF16) stage_store_bf16(stg, (bf16*)g, ld, r0, c0, M, lane, v, accumulate);
  else stage_store<float>(stg, (float*)g, ld, r0, c0, M, lane, v, accumulate);
}

// fp32 output whose rows are not 16 B aligned (the 1842-wide GQA logits): one column per lane, a 128 B row per instruction
__device__ __forceinline__ void stage_store_lanes(float* stg, float* g, int64_t ld, int r0, int c0, int M, int lane,
                                                  const float* v, int accumulate) {
#pragma unroll
  for (int j = 0; j < 32; ++j) stg[lane * STG_LD + j] = v[j];
  __syncwarp();
#pragma unroll 8
  for (int r = 0; r < 32; ++r) {
    if (r0 + r >= M) break;
    const float f = stg[r * STG_LD + lane];
    float* p = g + (int64_t)(r0 + r) * ld + c0 + lane;
    if (accumulate == 2) atomicAdd(p, f);
    else *p = accumulate == 1 ? *p + f : f;
  }
  __syncwarp();
}

// 1: every tensor of the epilogue can use 16 B row-coalesced accesses for full 32-column chunks; 2: all but an fp32
// output with unaligned rows, which is stored one column per lane; 0: per-thread path
__host__ inline int epi_stage_mode(const EpiArgs& e) {
  auto ok = [](const void* p, int64_t ld, int dtype) {
    if (!p) return true;
    const int64_t m = dtype == CLG_BF16 ? 8 : 4;
    return (reinterpret_cast<uintptr_t>(p) & 15) == 0 && (ld % m) == 0;
  };
  if (!(ok(e.residual, e.ld_res, CLG_F32) && ok(e.aux_in, e.ld_aux_in, e.aux_in_dtype) &&
        ok(e.aux_out, e.ld_aux_out, e.aux_out_dtype) && (!e.bias || (reinterpret_cast<uintptr_t>(e.bias) & 15) == 0)))
    return 0;
  if (ok(e.C, e.ldc, e.c_dtype)) return 1;
  return e.c_dtype == CLG_F32 ? 2 : 0;
}
__host__ inline bool epi_can_stage(const EpiArgs& e) { return epi_stage_mode(e) == 1; }

// warp-collective: rows r0 .. r0+31 (lane = row), columns c0 .. c0+31
__device__ __forceinline__ void tc_epi_chunk(const EpiArgs& e, float* stg, int r0, int c0, int lane, int staged, float* acc) {
  if (!staged || c0 + 32 > e.N) {          // ragged N tail or unaligned tensors: per-thread path
    epi_row_chunk<32>(e, r0 + lane, c0, acc);
    return;
  }
  const int row = r0 + lane;
  float v[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = acc[j] * e.alpha;
  if (e.bias) {
    const float4* b4 = reinterpret_cast<const float4*>(e.bias + c0);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 b = __ldg(b4 + j);
      v[4 * j + 0] += b.x; v[4 * j + 1] += b.y; v[4 * j + 2] += b.z; v[4 * j + 3] += b.w;
    }
  }
  const uint32_t eidx = (uint32_t)((uint64_t)row * (uint64_t)e.N + (uint64_t)c0);    // even: c0 % 32 == 0, N even
  auto dropout = [&]() {
#pragma unroll
    for (int j = 0; j < 32; j += 2) {
      const uint32_t h = lowbias32(((eidx + j) >> 1) ^ e.drop.key);
      v[j] = ((h & 0xffffu) >= e.drop.thresh) ? v[j] * e.drop.scale : 0.f;
      v[j + 1] = ((h >> 16) >= e.drop.thresh) ? v[j + 1] * e.drop.scale : 0.f;
    }
  };
  if (e.drop.on && e.drop_before_act) dropout();
  float aux[32];
  switch (e.act) {
    case CLG_ACT_GELU:
      if (e.aux_out) stage_store_any(stg, e.aux_out, e.aux_out_dtype, e.ld_aux_out, r0, c0, e.M, lane, v, 0);
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = gelu_erf(v[j]);
      break;
    case CLG_ACT_GELU_GRAD:
      if (e.c_dtype == CLG_BF16) {        // bf16 mode: one-exponential fast pair (error far below a bf16 ulp)
#pragma unroll
        for (int j = 0; j < 32; ++j) gelu_pair_fast(v[j], v[j], aux[j]);
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) { aux[j] = dgelu_erf(v[j]); v[j] = gelu_erf(v[j]); }
      }
      if (e.aux_out) stage_store_any(stg, e.aux_out, e.aux_out_dtype, e.ld_aux_out, r0, c0, e.M, lane, aux, 0);
      break;
    case CLG_ACT_MUL_AUX:
      stage_load_any(stg, e.aux_in, e.aux_in_dtype, e.ld_aux_in, r0, c0, e.M, lane, aux);
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] *= aux[j];
      break;
    case CLG_ACT_RELU:
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
      break;
    case CLG_ACT_TANH:
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = tanhf(v[j]);
      break;
    case CLG_ACT_DGELU:
      stage_load_any(stg, e.aux_in, e.aux_in_dtype, e.ld_aux_in, r0, c0, e.M, lane, aux);
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] *= dgelu_erf(aux[j]);
      break;
    case CLG_ACT_DRELU:
      stage_load_any(stg, e.aux_in, e.aux_in_dtype, e.ld_aux_in, r0, c0, e.M, lane, aux);
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = aux[j] > 0.f ? v[j] : 0.f;
      break;
    case CLG_ACT_DTANH:
      stage_load_any(stg, e.aux_in, e.aux_in_dtype, e.ld_aux_in, r0, c0, e.M, lane, aux);
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] *= (1.f - aux[j] * aux[j]);
      break;
    default: break;
  }
  if (e.drop.on && !e.drop_before_act) dropout();
  if (e.residual) {
    stage_load<float>(stg, e.residual, e.ld_res, r0, c0, e.M, lane, aux);
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] += aux[j];
  }
  if (e.mask_bits && row < e.M) {
    const uint64_t bidx = (uint64_t)row * (uint64_t)e.N + (uint64_t)c0;
    if ((bidx & 31) == 0) {
      const uint32_t word = __ldg(e.mask_bits + (bidx >> 5));
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] *= ((word >> j) & 1u) ? 1.f : 0.f;
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] *= mask_bit(e.mask_bits, (int64_t)(bidx + j)) ? 1.f : 0.f;
    }
  }
  if (staged == 2) stage_store_lanes(stg, (float*)e.C, e.ldc, r0, c0, e.M, lane, v, e.accumulate);
  else stage_store_any(stg, e.C, e.c_dtype, e.ldc, r0, c0, e.M, lane, v, e.accumulate);
  if (e.colsum) {
    if (row >= e.M) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = 0.f;
    }
    const float cs = warp_colsum32(v, lane);
    atomicAdd(e.colsum + c0 + lane, cs);
  }
}

}  // namespace clg
