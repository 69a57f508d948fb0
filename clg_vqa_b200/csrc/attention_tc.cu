// tcgen05 / TMEM / TMA fused attention for the bf16 mode (S <= 160, d_head = 64).
//
// One CTA (128 threads = 4 warps = the 128 TMEM lanes) per (query tile of 128 rows, head, sample).
// Forward:  TMA loads Q[128x64], K[NPx64], V[NPx64] (128B-swizzled boxes straight out of the fused
//           qkv activation, NP = keys rounded up to 16);  S = Q K^T with tcgen05.mma into TMEM;  every
//           thread owns one query row: 3 passes over its TMEM row (max, sum, normalised P) in fp32;
//           P (bf16, the oracle's rounding point) goes to swizzled smem;  O = P V with tcgen05.mma
//           (V as an MN-major B operand - no transpose);  O -> bf16 ctx.
// Backward: recompute S, dP = dO V^T (both in TMEM), P_drop and dS (bf16) to swizzled smem, then
//           dV = P_drop^T dO, dK = dS^T Q (MN-major A operands reading the SAME buffers transposed via
//           the descriptor) and dQ = dS K.
// Rows / keys beyond S inside a tile come from the neighbouring sample (or TMA zero fill) and are
// neutralised with -inf key bias / zeroed P rows, never stored.
#include <mutex>
#include <string.h>
#include "tc_common.cuh"

namespace clg {

static constexpr int ADH = 64;
static constexpr int QT = 128;           // query rows per CTA (= TMEM lanes)
static constexpr int A_THREADS = 128;

// smem tile helpers: [rows][64] bf16, 128 B per row, 16 B chunks XOR (row % 8)  (SWIZZLE_128B)
__device__ __forceinline__ uint32_t sw_off(int row, int col /*0..63*/) {
  return (uint32_t)(row * 128 + ((((col >> 3) ^ (row & 7)) << 4)) + ((col & 7) << 1));
}

struct AttnTcArgs {
  int B, S, H, NP;          // NP = keys padded to 16
  int n_qt;                 // query tiles per (b, h)
  float scale;
  const float* key_bias;    // [B, S]
  float* lse;               // [B, H, S]
  bf16* ctx; int64_t ld_ctx;
  // backward only
  const bf16* ctx_in; const bf16* dctx; int64_t ld_dctx;
  bf16* dqkv; int64_t ld_dqkv;
  Drop drop;
};

// ---- shared softmax-row helper: reads NP fp32 scores of this thread's TMEM row chunk by chunk
template <typename F>
__device__ __forceinline__ void for_each_score(uint32_t taddr, int NP, F&& f) {
  int c = 0;
  for (; c + 32 <= NP; c += 32) {
    float v[32];
    tmem_ld32(taddr + (uint32_t)c, v);
#pragma unroll
    for (int j = 0; j < 32; ++j) f(c + j, v[j]);
  }
  if (c < NP) {   // NP is a multiple of 16
    float v[16];
    tmem_ld16(taddr + (uint32_t)c, v);
#pragma unroll
    for (int j = 0; j < 16; ++j) f(c + j, v[j]);
  }
}

// ------------------------------------------------------------------------------------------------
// forward
// Shared memory: [barriers | Q (QR rows) | K (NP) | V (NP) | P (KB blocks of QR rows)] + 16 KiB read slack,
// QR = NP when the whole sequence fits one query tile (rows NP..127 of the M = 128 MMAs then alias the
// following tiles and only produce accumulator rows that are never used), else 128.
// TMEM: S in columns [0, NP); O reuses columns [0, 64) once P has been written to shared memory.
// NCH > 0: the NP = 16*NCH exponentials of a row stay in registers between the sum pass and the P pass.
// ------------------------------------------------------------------------------------------------
template <int NCH>
__global__ void __launch_bounds__(A_THREADS)
attn_tc_fwd_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_kv, const AttnTcArgs a,
                   const int QR, const uint32_t tmem_cols) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  const int NP = a.NP;
  const int KB = (NP + 63) / 64;                    // 64-wide key blocks of the P tile
  const uint32_t TQ = (uint32_t)QR * 128u, TN = (uint32_t)NP * 128u;
  const uint32_t sBar = base;
  const uint32_t sQ = base + 1024;
  const uint32_t sK = sQ + TQ;
  const uint32_t sV = sK + TN;
  const uint32_t sP = sV + TN;                      // KB x [QR][64]
  uint8_t* gP = gen + (sP - base);
  float* kb = reinterpret_cast<float*>(gen + 64);
  volatile uint32_t* tslot = reinterpret_cast<volatile uint32_t*>(gen + 32);
  const uint32_t bar_tma = sBar, bar_mma = sBar + 8;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int qt = blockIdx.x % a.n_qt;
  const int h = (blockIdx.x / a.n_qt) % a.H;
  const int b = blockIdx.x / (a.n_qt * a.H);
  const int S = a.S;

  if (tid == 0) {
    tma_prefetch_desc(&tm_q); tma_prefetch_desc(&tm_kv);
    mbar_init(bar_tma, 1); mbar_init(bar_mma, 1);
    fence_barrier_init(); fence_proxy_async();
  }
  if (warp == 0) tmem_alloc(sBar + 32, tmem_cols);
  for (int j = tid; j < NP; j += A_THREADS) kb[j] = (j < S) ? a.key_bias[(int64_t)b * S + j] : -INFINITY;
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = *tslot;
  const uint32_t tS = tmem, tO = tmem;

  if (tid == 0) {
    mbar_expect_tx(bar_tma, TQ + 2 * TN);
    tma_load_2d(sQ, &tm_q, bar_tma, h * ADH, b * S + qt * QT);
    tma_load_2d(sK, &tm_kv, bar_tma, a.H * ADH + h * ADH, b * S);
    tma_load_2d(sV, &tm_kv, bar_tma, 2 * a.H * ADH + h * ADH, b * S);
    mbar_wait(bar_tma, 0);
    tcgen05_fence_after();
    const uint32_t idesc = make_idesc_mn(128, NP, false, false);
#pragma unroll
    for (int k = 0; k < ADH / 16; ++k)
      umma_bf16(tS, umma_desc(sQ + k * 32, 16, 1024), umma_desc(sK + k * 32, 16, 1024), idesc, k > 0);
    umma_commit(bar_mma);
  }
  mbar_wait(bar_mma, 0);
  __syncwarp();
  tcgen05_fence_after();

  const int row = warp * 32 + lane;                 // row inside the query tile == TMEM lane
  const int qi = qt * QT + row;                     // query index inside the sample
  const uint32_t trow = tS + ((uint32_t)(warp * 32) << 16);
  const float scale = a.scale;
  const bool valid = qi < S;
  const bool in_tile = row < QR;
  const bool warp_live = (qt * QT + warp * 32) < S; // warp-uniform: any valid row in this warp
  const uint64_t drow = (((uint64_t)b * a.H + h) * S + (valid ? qi : 0)) * (uint64_t)S;
  float m = -INFINITY, sum = 0.f;
  uint32_t pk[4];
  auto emit = [&](int j, float p) {                  // 8 consecutive probabilities -> one swizzled 16 B chunk
    if (a.drop.on && j < S) p = drop_apply(a.drop, drow + j, p);
    const int e = j & 7;
    const uint32_t hbits = (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(p));
    if (e & 1) pk[e >> 1] |= hbits << 16; else pk[e >> 1] = hbits;
    if (e == 7 && in_tile)
      *reinterpret_cast<uint4*>(gP + (j >> 6) * TQ + sw_off(row, (j & 63) - 7)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
  };
  if (warp_live) {
    if (NCH > 0) {
      float ex[NCH > 0 ? NCH * 16 : 1];
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        tmem_ld16(trow + c * 16, &ex[c * 16]);
#pragma unroll
        for (int j = 0; j < 16; ++j) { ex[c * 16 + j] = ex[c * 16 + j] * scale + kb[c * 16 + j]; m = fmaxf(m, ex[c * 16 + j]); }
      }
#pragma unroll
      for (int j = 0; j < NCH * 16; ++j) { ex[j] = __expf(ex[j] - m); sum += ex[j]; }
      const float inv = valid ? 1.0f / sum : 0.f;
#pragma unroll
      for (int j = 0; j < NCH * 16; ++j) emit(j, ex[j] * inv);
    } else {
      for_each_score(trow, NP, [&](int j, float s) { m = fmaxf(m, s * scale + kb[j]); });
      for_each_score(trow, NP, [&](int j, float s) { sum += __expf(s * scale + kb[j] - m); });
      const float inv = valid ? 1.0f / sum : 0.f;
      for_each_score(trow, NP, [&](int j, float s) { emit(j, __expf(s * scale + kb[j] - m) * inv); });
    }
    if (valid && a.lse) a.lse[((int64_t)b * a.H + h) * S + qi] = m + __logf(sum);
  } else if (in_tile) {
    for (int j = 0; j < NP; j += 8)
      *reinterpret_cast<uint4*>(gP + (j >> 6) * TQ + sw_off(row, j & 63)) = make_uint4(0, 0, 0, 0);
  }
  (void)KB;
  fence_proxy_async();
  tcgen05_fence_before();
  __syncthreads();
  if (tid == 0) {
    tcgen05_fence_after();
    const uint32_t idesc = make_idesc_mn(128, ADH, false, true);
    const int ksteps = NP / 16;
    for (int k = 0; k < ksteps; ++k) {
      const uint32_t pa = sP + (k >> 2) * TQ + (k & 3) * 32;
      umma_bf16(tO, umma_desc(pa, 16, 1024), umma_desc(sV + k * 2048, 8192, 1024), idesc, k > 0);
    }
    umma_commit(bar_mma);
  }
  mbar_wait(bar_mma, 1);
  __syncwarp();
  tcgen05_fence_after();
  if (warp_live) {
    float o[32];
    bf16* og = a.ctx + ((int64_t)b * S + qi) * a.ld_ctx + h * ADH;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      tmem_ld32(tO + ((uint32_t)(warp * 32) << 16) + half * 32, o);
      if (valid) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          __nv_bfloat162 h0 = __floats2bfloat162_rn(o[8 * j + 0], o[8 * j + 1]);
          __nv_bfloat162 h1 = __floats2bfloat162_rn(o[8 * j + 2], o[8 * j + 3]);
          __nv_bfloat162 h2 = __floats2bfloat162_rn(o[8 * j + 4], o[8 * j + 5]);
          __nv_bfloat162 h3 = __floats2bfloat162_rn(o[8 * j + 6], o[8 * j + 7]);
          uint4 u;
          u.x = *reinterpret_cast<uint32_t*>(&h0); u.y = *reinterpret_cast<uint32_t*>(&h1);
          u.z = *reinterpret_cast<uint32_t*>(&h2); u.w = *reinterpret_cast<uint32_t*>(&h3);
          *reinterpret_cast<uint4*>(og + half * 32 + 8 * j) = u;
        }
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) { __syncwarp(); tcgen05_fence_after(); tmem_dealloc(tmem, tmem_cols); }
}

// ------------------------------------------------------------------------------------------------
// backward (single query tile: S <= 128; S > 128 uses the FFMA kernel)
// Shared memory holds only NP = ceil16(S) rows per operand tile and TMEM only 256 columns (the dV / dK /
// dQ accumulators reuse the S / dP columns once those have been read), so two CTAs fit per SM.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(A_THREADS)
attn_tc_bwd_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_do, const AttnTcArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  const int NP = a.NP;                               // <= 128, multiple of 16
  const uint32_t TN = (uint32_t)NP * 128u;           // bytes of one [NP][64] bf16 tile (multiple of 2 KiB)
  const uint32_t sBar = base;                        // barriers, tmem slot, key bias (first KiB)
  const uint32_t sQ = base + 1024, sK = sQ + TN, sV = sK + TN, sdO = sV + TN;
  const uint32_t sP = sdO + TN;                      // 2 x [NP][64]: P_drop  (rows i, cols j)
  const uint32_t sdS = sP + 2 * TN;                  // 2 x [NP][64]: dS      (+ 16 KiB slack behind it: the M = 128
                                                     //  K-major reads of dS touch rows NP..127 of a block)
  uint8_t* gP = gen + (sP - base);
  uint8_t* gdS = gen + (sdS - base);
  float* kb = reinterpret_cast<float*>(gen + 64);
  volatile uint32_t* tslot = reinterpret_cast<volatile uint32_t*>(gen + 32);
  const uint32_t bar_tma = sBar, bar_mma = sBar + 8;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int h = blockIdx.x % a.H, b = blockIdx.x / a.H;
  const int S = a.S;

  if (tid == 0) {
    tma_prefetch_desc(&tm_q); tma_prefetch_desc(&tm_do);
    mbar_init(bar_tma, 1); mbar_init(bar_mma, 1);
    fence_barrier_init(); fence_proxy_async();
  }
  if (warp == 0) tmem_alloc(sBar + 32, 256);
  for (int j = tid; j < 128; j += A_THREADS) kb[j] = (j < S) ? a.key_bias[(int64_t)b * S + j] : -INFINITY;
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = *tslot;
  // TMEM columns: S [0,128) | dP [128,256)   then   dV [0,64) | dK [64,128) | dQ [128,192)
  const uint32_t tS = tmem, tdP = tmem + 128, tdV = tmem, tdK = tmem + 64, tdQ = tmem + 128;

  if (tid == 0) {
    mbar_expect_tx(bar_tma, 4 * TN);
    tma_load_2d(sQ, &tm_q, bar_tma, h * ADH, b * S);
    tma_load_2d(sK, &tm_q, bar_tma, a.H * ADH + h * ADH, b * S);
    tma_load_2d(sV, &tm_q, bar_tma, 2 * a.H * ADH + h * ADH, b * S);
    tma_load_2d(sdO, &tm_do, bar_tma, h * ADH, b * S);
    mbar_wait(bar_tma, 0);
    tcgen05_fence_after();
    // M = 128 rows are read from the NP-row Q / dO tiles: rows NP..127 alias the following tiles and only
    // produce accumulator rows that are never used
    const uint32_t idesc = make_idesc_mn(128, NP, false, false);
#pragma unroll
    for (int k = 0; k < ADH / 16; ++k) {
      umma_bf16(tS, umma_desc(sQ + k * 32, 16, 1024), umma_desc(sK + k * 32, 16, 1024), idesc, k > 0);
      umma_bf16(tdP, umma_desc(sdO + k * 32, 16, 1024), umma_desc(sV + k * 32, 16, 1024), idesc, k > 0);
    }
    umma_commit(bar_mma);
  }
  // D_i = dO_i . O_i while the tensor core works
  const int row = warp * 32 + lane;
  const bool valid = row < S;
  float Di = 0.f, lse_i = 0.f;
  if (valid) {
    const bf16* dor = a.dctx + ((int64_t)b * S + row) * a.ld_dctx + h * ADH;
    const bf16* orow = a.ctx_in + ((int64_t)b * S + row) * a.ld_ctx + h * ADH;
#pragma unroll
    for (int c = 0; c < ADH; c += 8) {
      const uint4 u = *reinterpret_cast<const uint4*>(dor + c);
      const uint4 w = *reinterpret_cast<const uint4*>(orow + c);
      const __nv_bfloat162* x = reinterpret_cast<const __nv_bfloat162*>(&u);
      const __nv_bfloat162* y = reinterpret_cast<const __nv_bfloat162*>(&w);
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const float2 fx = __bfloat1622float2(x[t]), fy = __bfloat1622float2(y[t]);
        Di += fx.x * fy.x + fx.y * fy.y;
      }
    }
    lse_i = a.lse[((int64_t)b * a.H + h) * S + row];
  }
  mbar_wait(bar_mma, 0);
  __syncwarp();
  tcgen05_fence_after();
  {
    const uint32_t lane_off = ((uint32_t)(warp * 32) << 16);
    const uint64_t drow = (((uint64_t)b * a.H + h) * S + (valid ? row : 0)) * (uint64_t)S;
    const float scale = a.scale;
    const bool in_tile = row < NP;                  // the P / dS buffers only have NP rows
    for (int c = 0; c < NP; c += 16) {
      float sv[16], dv[16];
      tmem_ld16_nowait(tS + lane_off + c, sv);
      tmem_ld16_nowait(tdP + lane_off + c, dv);
      tmem_wait_ld();
      uint32_t pp[8], dd[8];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int col = c + j;
        float p = 0.f, ds = 0.f;
        if (valid && col < S) {
          p = __expf(sv[j] * scale + kb[col] - lse_i);
          float dp = dv[j];
          float pd = p;
          if (a.drop.on) {
            const bool keep = drop_keep(a.drop, drow + col);
            pd = keep ? p * a.drop.scale : 0.f;
            dp = keep ? dp * a.drop.scale : 0.f;
          }
          ds = p * (dp - Di);
          p = pd;
        }
        const uint32_t pb = (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(p));
        const uint32_t db = (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(ds));
        if (j & 1) { pp[j >> 1] |= pb << 16; dd[j >> 1] |= db << 16; } else { pp[j >> 1] = pb; dd[j >> 1] = db; }
      }
      if (in_tile) {
        const int kblk = c >> 6, cc = c & 63;
        *reinterpret_cast<uint4*>(gP + kblk * TN + sw_off(row, cc)) = make_uint4(pp[0], pp[1], pp[2], pp[3]);
        *reinterpret_cast<uint4*>(gP + kblk * TN + sw_off(row, cc + 8)) = make_uint4(pp[4], pp[5], pp[6], pp[7]);
        *reinterpret_cast<uint4*>(gdS + kblk * TN + sw_off(row, cc)) = make_uint4(dd[0], dd[1], dd[2], dd[3]);
        *reinterpret_cast<uint4*>(gdS + kblk * TN + sw_off(row, cc + 8)) = make_uint4(dd[4], dd[5], dd[6], dd[7]);
      }
    }
    if (in_tile) {
      for (int c = NP; c < 128; c += 8) {           // zero the unused key columns (read as M rows below)
        *reinterpret_cast<uint4*>(gP + (c >> 6) * TN + sw_off(row, c & 63)) = make_uint4(0, 0, 0, 0);
        *reinterpret_cast<uint4*>(gdS + (c >> 6) * TN + sw_off(row, c & 63)) = make_uint4(0, 0, 0, 0);
      }
    }
  }
  fence_proxy_async();
  tcgen05_fence_before();
  __syncthreads();
  if (tid == 0) {
    tcgen05_fence_after();
    // dV[j,d] = sum_i P[i,j] dO[i,d]   A = P^T (MN-major over the [i][j] buffer), B = dO (MN-major), K = i (NP)
    // dK[j,d] = sum_i dS[i,j] Q[i,d]   A = dS^T (MN-major),                      B = Q  (MN-major)
    const uint32_t id_t = make_idesc_mn(128, ADH, true, true);
    for (int k = 0; k < NP / 16; ++k) {
      // MN-major A: 64-wide M blocks (j) are the two [NP][64] sub-tiles (LBO = TN); 8-row k groups 1024 B apart
      umma_bf16(tdV, umma_desc(sP + k * 2048, TN, 1024), umma_desc(sdO + k * 2048, 8192, 1024), id_t, k > 0);
      umma_bf16(tdK, umma_desc(sdS + k * 2048, TN, 1024), umma_desc(sQ + k * 2048, 8192, 1024), id_t, k > 0);
    }
    // dQ[i,d] = sum_j dS[i,j] K[j,d]   A = dS (K-major), B = K (MN-major), K = j (NP)
    const uint32_t id_q = make_idesc_mn(128, ADH, false, true);
    for (int k = 0; k < NP / 16; ++k)
      umma_bf16(tdQ, umma_desc(sdS + (k >> 2) * TN + (k & 3) * 32, 16, 1024), umma_desc(sK + k * 2048, 8192, 1024), id_q, k > 0);
    umma_commit(bar_mma);
  }
  mbar_wait(bar_mma, 1);
  __syncwarp();
  tcgen05_fence_after();
  {
    const uint32_t lane_off = ((uint32_t)(warp * 32) << 16);
    bf16* out = a.dqkv + ((int64_t)b * S + row) * a.ld_dqkv + h * ADH;
    const int Hd = a.H * ADH;
    const uint32_t srcs[3] = {tdQ, tdK, tdV};
    const float mul[3] = {a.scale, a.scale, 1.f};
#pragma unroll
    for (int which = 0; which < 3; ++which) {
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        float o[32];
        tmem_ld32(srcs[which] + lane_off + half * 32, o);
        if (valid) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float s = mul[which];
            __nv_bfloat162 h0 = __floats2bfloat162_rn(o[8 * j + 0] * s, o[8 * j + 1] * s);
            __nv_bfloat162 h1 = __floats2bfloat162_rn(o[8 * j + 2] * s, o[8 * j + 3] * s);
            __nv_bfloat162 h2 = __floats2bfloat162_rn(o[8 * j + 4] * s, o[8 * j + 5] * s);
            __nv_bfloat162 h3 = __floats2bfloat162_rn(o[8 * j + 6] * s, o[8 * j + 7] * s);
            uint4 u;
            u.x = *reinterpret_cast<uint32_t*>(&h0); u.y = *reinterpret_cast<uint32_t*>(&h1);
            u.z = *reinterpret_cast<uint32_t*>(&h2); u.w = *reinterpret_cast<uint32_t*>(&h3);
            *reinterpret_cast<uint4*>(out + which * Hd + half * 32 + 8 * j) = u;
          }
        }
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) { __syncwarp(); tcgen05_fence_after(); tmem_dealloc(tmem, 256); }
}


// ------------------------------------------------------------------------------------------------
// backward for 128 < S <= 160 (M3P / X101: 100 regions + 40 tokens).  The single-tile kernel above puts the
// queries on the 128 TMEM lanes; here the KEYS go on the lanes and the CTA walks two key tiles (0..127, 128..S-1):
//   S^T_t = K_t Q^T, dP^T_t = V_t dO^T            [128 keys x NQ queries] in TMEM columns [0,160) / [160,320)
//   P^T_t, dS^T_t (bf16) -> shared memory          one thread per key row, loop over the queries
//   dV_t = P^T_t dO, dK_t = dS^T_t Q               complete per key tile (columns [0,64) / [64,128), reusing S^T)
//   dQ  += dS_t K_t                                accumulated over the key tiles in columns [320,384) (queries 0..127)
//                                                  and [384,448) (queries 128..S-1), A = (dS^T_t)^T via an MN-major descriptor
// Shared memory: 4 KiB header (barriers, lse*log2e, D_i) | Q, dO [NQ][64] | K_t, V_t [128][64] | P^T, dS^T: 4 blocks of
// [128][64] each (block 3 is only read as the unused upper half of the second dQ row tile).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(A_THREADS, 1)
attn_tc_bwd_wide_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_kv,
                        const __grid_constant__ CUtensorMap tm_do, const AttnTcArgs a, const int NQ) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t TQ = (uint32_t)NQ * 128u, TB = 16384u;
  const uint32_t sQ = base + 4096, sdO = sQ + ((TQ + 1023u) & ~1023u), sK = sdO + ((TQ + 1023u) & ~1023u), sV = sK + TB;
  const uint32_t sPT = sV + TB, sdST = sPT + 4 * TB;
  uint8_t* gPT = gen + (sPT - base);
  uint8_t* gdST = gen + (sdST - base);
  float* lse2 = reinterpret_cast<float*>(gen + 1024);
  float* Dm = reinterpret_cast<float*>(gen + 2048);
  volatile uint32_t* tslot = reinterpret_cast<volatile uint32_t*>(gen + 64);
  const uint32_t bar_q = base, bar_kv = base + 8, bar_mma = base + 16;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int h = blockIdx.x % a.H, b = blockIdx.x / a.H;
  const int S = a.S, Hd = a.H * ADH;
  const int n_kt = (S + 127) / 128;
  constexpr float L2E = 1.4426950408889634f;

  if (tid == 0) {
    tma_prefetch_desc(&tm_q); tma_prefetch_desc(&tm_kv); tma_prefetch_desc(&tm_do);
    mbar_init(bar_q, 1); mbar_init(bar_kv, 1); mbar_init(bar_mma, 1);
    fence_barrier_init(); fence_proxy_async();
  }
  if (warp == 0) tmem_alloc(base + 64, 512);
  // D_i = dO_i . O_i and lse_i * log2(e) for every query row (two rows per thread when S > 128)
  for (int i = tid; i < 192; i += A_THREADS) {
    float di = 0.f, l2 = INFINITY;                 // rows >= S: p = 2^(-inf) = 0
    if (i < S) {
      const bf16* dor = a.dctx + ((int64_t)b * S + i) * a.ld_dctx + h * ADH;
      const bf16* orow = a.ctx_in + ((int64_t)b * S + i) * a.ld_ctx + h * ADH;
#pragma unroll
      for (int c = 0; c < ADH; c += 8) {
        const uint4 u = *reinterpret_cast<const uint4*>(dor + c);
        const uint4 w = *reinterpret_cast<const uint4*>(orow + c);
        const __nv_bfloat162* x = reinterpret_cast<const __nv_bfloat162*>(&u);
        const __nv_bfloat162* y = reinterpret_cast<const __nv_bfloat162*>(&w);
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const float2 fx = __bfloat1622float2(x[t]), fy = __bfloat1622float2(y[t]);
          di += fx.x * fy.x + fx.y * fy.y;
        }
      }
      l2 = a.lse[((int64_t)b * a.H + h) * S + i] * L2E;
    }
    Dm[i] = di; lse2[i] = l2;
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = *tslot;
  const uint32_t tST = tmem, tdPT = tmem + 160, tdQA = tmem + 320, tdQB = tmem + 384, tdV = tmem, tdK = tmem + 64;
  const uint32_t lane_off = ((uint32_t)(warp * 32) << 16);
  const int j = warp * 32 + lane;                  // key row inside the tile == TMEM lane
  const float scale2 = a.scale * L2E;
  const uint64_t dhead = ((uint64_t)b * a.H + h) * (uint64_t)S;

  if (tid == 0) {
    mbar_expect_tx(bar_q, 2 * TQ);
    tma_load_2d(sQ, &tm_q, bar_q, h * ADH, b * S);
    tma_load_2d(sdO, &tm_do, bar_q, h * ADH, b * S);
  }
  for (int kt = 0; kt < n_kt; ++kt) {
    if (tid == 0) {
      mbar_expect_tx(bar_kv, 2 * TB);
      tma_load_2d(sK, &tm_kv, bar_kv, Hd + h * ADH, b * S + kt * 128);
      tma_load_2d(sV, &tm_kv, bar_kv, 2 * Hd + h * ADH, b * S + kt * 128);
      if (kt == 0) mbar_wait(bar_q, 0);
      mbar_wait(bar_kv, kt & 1);
      tcgen05_fence_after();
      const uint32_t idesc = make_idesc_mn(128, NQ, false, false);
#pragma unroll
      for (int k = 0; k < ADH / 16; ++k) {
        umma_bf16(tST, umma_desc(sK + k * 32, 16, 1024), umma_desc(sQ + k * 32, 16, 1024), idesc, k > 0);
        umma_bf16(tdPT, umma_desc(sV + k * 32, 16, 1024), umma_desc(sdO + k * 32, 16, 1024), idesc, k > 0);
      }
      umma_commit(bar_mma);
    }
    const int jg = kt * 128 + j;
    const bool validj = jg < S;
    const float kb2 = validj ? a.key_bias[(int64_t)b * S + jg] * L2E : -INFINITY;
    mbar_wait(bar_mma, 0);                         // two commits per key tile: phases 0, 1
    __syncwarp();
    tcgen05_fence_after();
    for (int c = 0; c < NQ; c += 16) {
      float sv[16], dv[16];
      tmem_ld16_nowait(tST + lane_off + c, sv);
      tmem_ld16_nowait(tdPT + lane_off + c, dv);
      tmem_wait_ld_dep16(sv);
      tmem_wait_ld_dep16(dv);
      uint32_t pp[8], dd[8];
#pragma unroll
      for (int t = 0; t < 16; ++t) {
        const int i = c + t;
        float p = exp2f(fmaf(sv[t], scale2, kb2) - lse2[i]);      // 0 for masked keys and for query rows >= S
        if (!validj) p = 0.f;
        float dp = dv[t], pd = p;
        if (a.drop.on) {
          const bool keep = drop_keep(a.drop, (dhead + (uint64_t)(i < S ? i : 0)) * (uint64_t)S + (uint64_t)(validj ? jg : 0));
          pd = keep ? p * a.drop.scale : 0.f;
          dp = keep ? dp * a.drop.scale : 0.f;
        }
        const float ds = p * (dp - Dm[i]);
        const uint32_t pb = (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(pd));
        const uint32_t db = (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(p > 0.f ? ds : 0.f));
        if (t & 1) { pp[t >> 1] |= pb << 16; dd[t >> 1] |= db << 16; } else { pp[t >> 1] = pb; dd[t >> 1] = db; }
      }
      const int blk = c >> 6, cc = c & 63;
      *reinterpret_cast<uint4*>(gPT + blk * TB + sw_off(j, cc)) = make_uint4(pp[0], pp[1], pp[2], pp[3]);
      *reinterpret_cast<uint4*>(gPT + blk * TB + sw_off(j, cc + 8)) = make_uint4(pp[4], pp[5], pp[6], pp[7]);
      *reinterpret_cast<uint4*>(gdST + blk * TB + sw_off(j, cc)) = make_uint4(dd[0], dd[1], dd[2], dd[3]);
      *reinterpret_cast<uint4*>(gdST + blk * TB + sw_off(j, cc + 8)) = make_uint4(dd[4], dd[5], dd[6], dd[7]);
    }
    fence_proxy_async();
    tcgen05_fence_before();
    __syncthreads();
    if (tid == 0) {
      tcgen05_fence_after();
      // dV_t[j,d] = sum_i P^T[j,i] dO[i,d]   A = P^T (K-major, k = i), B = dO (MN-major)
      // dK_t[j,d] = sum_i dS^T[j,i] Q[i,d]
      const uint32_t id_kv = make_idesc_mn(128, ADH, false, true);
      for (int k = 0; k < NQ / 16; ++k) {
        umma_bf16(tdV, umma_desc(sPT + (k >> 2) * TB + (k & 3) * 32, 16, 1024), umma_desc(sdO + k * 2048, 8192, 1024), id_kv, k > 0);
        umma_bf16(tdK, umma_desc(sdST + (k >> 2) * TB + (k & 3) * 32, 16, 1024), umma_desc(sQ + k * 2048, 8192, 1024), id_kv, k > 0);
      }
      // dQ[i,d] += sum_j dS^T[j,i] K_t[j,d]   A = (dS^T)^T (MN-major over the [j][i] blocks, M = i), B = K_t (MN-major), k = j
      const uint32_t id_q = make_idesc_mn(128, ADH, true, true);
      for (int k = 0; k < 8; ++k) {
        umma_bf16(tdQA, umma_desc(sdST + k * 2048, TB, 1024), umma_desc(sK + k * 2048, 8192, 1024), id_q, (kt > 0 || k > 0) ? 1u : 0u);
        if (NQ > 128)
          umma_bf16(tdQB, umma_desc(sdST + 2 * TB + k * 2048, TB, 1024), umma_desc(sK + k * 2048, 8192, 1024), id_q, (kt > 0 || k > 0) ? 1u : 0u);
      }
      umma_commit(bar_mma);
    }
    mbar_wait(bar_mma, 1);
    __syncwarp();
    tcgen05_fence_after();
    {
      const uint32_t srcs[2] = {tdK, tdV};
      const float mul[2] = {a.scale, 1.f};
#pragma unroll
      for (int which = 0; which < 2; ++which) {
        float o[64];
        tmem_ld32_nowait(srcs[which] + lane_off, o);
        tmem_ld32_nowait(srcs[which] + lane_off + 32, o + 32);
        tmem_wait_ld_dep32(o);
        tmem_wait_ld_dep32(o + 32);
        if (validj) {
          uint4* og = reinterpret_cast<uint4*>(a.dqkv + ((int64_t)b * S + jg) * a.ld_dqkv + (which + 1) * Hd + h * ADH);
          const float sc = mul[which];
#pragma unroll
          for (int t = 0; t < 8; ++t) {
            __nv_bfloat162 h0 = __floats2bfloat162_rn(o[8 * t + 0] * sc, o[8 * t + 1] * sc);
            __nv_bfloat162 h1 = __floats2bfloat162_rn(o[8 * t + 2] * sc, o[8 * t + 3] * sc);
            __nv_bfloat162 h2 = __floats2bfloat162_rn(o[8 * t + 4] * sc, o[8 * t + 5] * sc);
            __nv_bfloat162 h3 = __floats2bfloat162_rn(o[8 * t + 6] * sc, o[8 * t + 7] * sc);
            uint4 u;
            u.x = *reinterpret_cast<uint32_t*>(&h0); u.y = *reinterpret_cast<uint32_t*>(&h1);
            u.z = *reinterpret_cast<uint32_t*>(&h2); u.w = *reinterpret_cast<uint32_t*>(&h3);
            og[t] = u;
          }
        }
      }
    }
    tcgen05_fence_before();
    __syncthreads();                               // the next key tile overwrites K_t / V_t, P^T / dS^T and S^T / dP^T
    tcgen05_fence_after();
  }
  // dQ: query rows 0..127 (tile A) and 128..S-1 (tile B)
  for (int qt = 0; qt < (NQ > 128 ? 2 : 1); ++qt) {
    const int i = qt * 128 + j;
    if (qt * 128 + warp * 32 < S) {
      float o[64];
      const uint32_t src = (qt ? tdQB : tdQA) + lane_off;
      tmem_ld32_nowait(src, o);
      tmem_ld32_nowait(src + 32, o + 32);
      tmem_wait_ld_dep32(o);
      tmem_wait_ld_dep32(o + 32);
      if (i < S) {
        uint4* og = reinterpret_cast<uint4*>(a.dqkv + ((int64_t)b * S + i) * a.ld_dqkv + h * ADH);
        const float sc = a.scale;
#pragma unroll
        for (int t = 0; t < 8; ++t) {
          __nv_bfloat162 h0 = __floats2bfloat162_rn(o[8 * t + 0] * sc, o[8 * t + 1] * sc);
          __nv_bfloat162 h1 = __floats2bfloat162_rn(o[8 * t + 2] * sc, o[8 * t + 3] * sc);
          __nv_bfloat162 h2 = __floats2bfloat162_rn(o[8 * t + 4] * sc, o[8 * t + 5] * sc);
          __nv_bfloat162 h3 = __floats2bfloat162_rn(o[8 * t + 6] * sc, o[8 * t + 7] * sc);
          uint4 u;
          u.x = *reinterpret_cast<uint32_t*>(&h0); u.y = *reinterpret_cast<uint32_t*>(&h1);
          u.z = *reinterpret_cast<uint32_t*>(&h2); u.w = *reinterpret_cast<uint32_t*>(&h3);
          og[t] = u;
        }
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) { __syncwarp(); tcgen05_fence_after(); tmem_dealloc(tmem, 512); }
}

static int fwd_smem_bytes(int QR, int NP) {
  const int KB = (NP + 63) / 64;
  return 1024 /*align*/ + 1024 /*barriers, key bias*/ + QR * 128 + 2 * NP * 128 + KB * QR * 128 + 16384 /*read slack*/;
}
static int bwd_smem_bytes(int NP) { return 1024 /*align*/ + 1024 /*barriers, key bias*/ + 8 * NP * 128 + 16384 /*read slack*/; }

template <int NCH>
static int launch_attn_fwd(const CUtensorMap& tq, const CUtensorMap& tkv, const AttnTcArgs& a, int QR, uint32_t tmem_cols, int grid,
                           cudaStream_t stream) {
  static std::once_flag once;
  static cudaError_t err = cudaSuccess;
  std::call_once(once, [] {
    err = cudaFuncSetAttribute(attn_tc_fwd_kernel<NCH>, cudaFuncAttributeMaxDynamicSharedMemorySize, fwd_smem_bytes(128, 160));
  });
  CLG_CUDA_OK(err, "cudaFuncSetAttribute(attn_tc_fwd)");
  attn_tc_fwd_kernel<NCH><<<grid, A_THREADS, fwd_smem_bytes(QR, a.NP), stream>>>(tq, tkv, a, QR, tmem_cols);
  return check_launch("attn_tc_fwd_kernel");
}

int attention_tc_fwd(const clgAttentionFwdParams& p, cudaStream_t stream) {
  if (p.dtype != CLG_BF16 || p.S > 160 || p.S < 1) return CLG_ERR_UNSUPPORTED;
  const int NP = (p.S + 15) / 16 * 16;
  const int n_qt = (p.S + QT - 1) / QT;
  const int QR = (n_qt == 1) ? NP : QT;
  CUtensorMap tq, tkv;
  const int64_t rows = (int64_t)p.B * p.S, cols = 3LL * p.H * ADH;
  int st = make_tmap_bf16(&tq, p.qkv, cols, rows, p.ld_qkv, QR);
  if (st != CLG_OK) return st;
  st = make_tmap_bf16(&tkv, p.qkv, cols, rows, p.ld_qkv, NP);
  if (st != CLG_OK) return st;
  AttnTcArgs a;
  memset(&a, 0, sizeof(a));
  a.B = p.B; a.S = p.S; a.H = p.H; a.NP = NP; a.n_qt = n_qt;
  a.scale = p.scale; a.key_bias = p.key_bias; a.lse = p.lse;
  a.ctx = (bf16*)p.ctx; a.ld_ctx = p.ld_ctx;
  a.drop = make_drop(p.dropout);
  const uint32_t tmem_cols = NP <= 64 ? 64 : (NP <= 128 ? 128 : 256);
  const int grid = p.B * p.H * n_qt;
  switch (NP / 16) {
    case 1: return launch_attn_fwd<1>(tq, tkv, a, QR, tmem_cols, grid, stream);
    case 2: return launch_attn_fwd<2>(tq, tkv, a, QR, tmem_cols, grid, stream);
    case 3: return launch_attn_fwd<3>(tq, tkv, a, QR, tmem_cols, grid, stream);
    case 4: return launch_attn_fwd<4>(tq, tkv, a, QR, tmem_cols, grid, stream);
    case 5: return launch_attn_fwd<5>(tq, tkv, a, QR, tmem_cols, grid, stream);
    case 6: return launch_attn_fwd<6>(tq, tkv, a, QR, tmem_cols, grid, stream);
    default: return launch_attn_fwd<0>(tq, tkv, a, QR, tmem_cols, grid, stream);
  }
}

static int attention_tc_bwd_wide(const clgAttentionBwdParams& p, cudaStream_t stream) {
  const int NQ = (p.S + 15) / 16 * 16;
  const int TQ = ((NQ * 128) + 1023) & ~1023;
  const int smem = 1024 + 4096 + 2 * TQ + 2 * 16384 + 8 * 16384;
  static std::once_flag once;
  static cudaError_t err = cudaSuccess;
  std::call_once(once, [] { err = cudaFuncSetAttribute(attn_tc_bwd_wide_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448); });
  CLG_CUDA_OK(err, "cudaFuncSetAttribute(attn_tc_bwd_wide)");
  if (smem > 232448) return CLG_ERR_UNSUPPORTED;
  CUtensorMap tq, tkv, tdo;
  const int64_t rows = (int64_t)p.B * p.S;
  int st = make_tmap_bf16(&tq, p.qkv, 3LL * p.H * ADH, rows, p.ld_qkv, NQ);
  if (st != CLG_OK) return st;
  st = make_tmap_bf16(&tkv, p.qkv, 3LL * p.H * ADH, rows, p.ld_qkv, 128);
  if (st != CLG_OK) return st;
  st = make_tmap_bf16(&tdo, p.dctx, (int64_t)p.H * ADH, rows, p.ld_dctx, NQ);
  if (st != CLG_OK) return st;
  AttnTcArgs a;
  memset(&a, 0, sizeof(a));
  a.B = p.B; a.S = p.S; a.H = p.H; a.NP = NQ; a.n_qt = 1;
  a.scale = p.scale; a.key_bias = p.key_bias; a.lse = const_cast<float*>(p.lse);
  a.ctx_in = (const bf16*)p.ctx; a.ld_ctx = p.ld_ctx;
  a.dctx = (const bf16*)p.dctx; a.ld_dctx = p.ld_dctx;
  a.dqkv = (bf16*)p.dqkv; a.ld_dqkv = p.ld_dqkv;
  a.drop = make_drop(p.dropout);
  attn_tc_bwd_wide_kernel<<<p.B * p.H, A_THREADS, smem, stream>>>(tq, tkv, tdo, a, NQ);
  return check_launch("attn_tc_bwd_wide_kernel");
}

int attention_tc_bwd(const clgAttentionBwdParams& p, cudaStream_t stream) {
  if (p.dtype != CLG_BF16 || p.S < 1) return CLG_ERR_UNSUPPORTED;
  if (p.S > 128 && p.S <= 160) return attention_tc_bwd_wide(p, stream);
  if (p.S > 128) return CLG_ERR_UNSUPPORTED;
  static std::once_flag once;
  static cudaError_t err = cudaSuccess;
  std::call_once(once, [] { err = cudaFuncSetAttribute(attn_tc_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bwd_smem_bytes(128)); });
  CLG_CUDA_OK(err, "cudaFuncSetAttribute(attn_tc_bwd)");
  const int NP = (p.S + 15) / 16 * 16;
  const int BWD_SMEM = bwd_smem_bytes(NP);
  CUtensorMap tq, tdo;
  const int64_t rows = (int64_t)p.B * p.S;
  int st = make_tmap_bf16(&tq, p.qkv, 3LL * p.H * ADH, rows, p.ld_qkv, NP);
  if (st != CLG_OK) return st;
  st = make_tmap_bf16(&tdo, p.dctx, (int64_t)p.H * ADH, rows, p.ld_dctx, NP);
  if (st != CLG_OK) return st;
  AttnTcArgs a;
  memset(&a, 0, sizeof(a));
  a.B = p.B; a.S = p.S; a.H = p.H; a.NP = NP; a.n_qt = 1;
  a.scale = p.scale; a.key_bias = p.key_bias; a.lse = const_cast<float*>(p.lse);
  a.ctx_in = (const bf16*)p.ctx; a.ld_ctx = p.ld_ctx;
  a.dctx = (const bf16*)p.dctx; a.ld_dctx = p.ld_dctx;
  a.dqkv = (bf16*)p.dqkv; a.ld_dqkv = p.ld_dqkv;
  a.drop = make_drop(p.dropout);
  attn_tc_bwd_kernel<<<p.B * p.H, A_THREADS, BWD_SMEM, stream>>>(tq, tdo, a);
  return check_launch("attn_tc_bwd_kernel");
}

}  // namespace clg
