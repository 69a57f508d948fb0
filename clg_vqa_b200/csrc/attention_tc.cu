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

int attention_tc_bwd(const clgAttentionBwdParams& p, cudaStream_t stream) {
  if (p.dtype != CLG_BF16 || p.S > 128 || p.S < 1) return CLG_ERR_UNSUPPORTED;
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
