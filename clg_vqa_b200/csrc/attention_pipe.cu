// Persistent, warp-specialised tcgen05 attention for the short joint sequence (bf16 mode, d_head = 64,
// forward S <= 128, backward S <= 96).  Same math and rounding points as attention_tc.cu (which stays the path
// for longer sequences); what changes is the schedule: the one-CTA-per-(sample, head) kernels were a serial
// latency chain (TMEM alloc -> TMA -> MMA -> softmax -> MMA -> store) with two resident CTAs per SM.  Here one
// CTA per SM walks its (sample, head) items with
//   warp 0      TMA producer: Q/K/V (+ dO, O in the backward) boxes of item i+1.. into a ring of stages
//   warp 1      MMA issuer:   S = Q K^T (and dP = dO V^T) of item i+1 is issued before the P V (dV/dK/dQ) MMAs
//                             of item i, so the tensor core never waits for the softmax of the same item
//   warps 2-5 / 6-9   two softmax groups (128 threads = the 128 TMEM lanes each) working on alternate items,
//                             each with its own TMEM accumulators, P / dS tiles and output staging
// Results leave through TMA stores whose box is exactly S rows tall, so nothing of the neighbouring sample is
// touched.  Scores never reach HBM.
#include <mutex>
#include <stdlib.h>
#include <string.h>
#include "tc_common.cuh"

namespace clg {

static constexpr int PD = 64;             // head dim
static constexpr int P_THREADS = 320;

struct AttnPipeArgs {
  int B, S, H, NP;                        // NP = keys / queries padded to 16
  int items;                              // B * H
  int nst;                                // smem stages
  float scale;
  const float* key_bias;                  // [B, S]
  float* lse;                             // [B, H, S]
  Drop drop;
  // packed variable-length layout (optional): sample b owns rows [cu[b], cu[b+1]); outputs are then written row by row
  // (a TMA box of fixed height would spill into the next sample)
  const int32_t* cu;
  bf16* out; int64_t ld_out;              // ctx (forward) / dqkv (backward), used by the varlen stores
  long long* trace;                       // debug (tools/attn_trace.py): clock64 stamps of CTA 0's first items, else NULL
};

// trace[(role * 16 + item) * 8 + event] = clock64() - roles: 0 TMA producer, 1 MMA issuer, 2 + g softmax group g
#define ATR(role, item, ev) do { if constexpr (TRACE) { if (a.trace && blockIdx.x == 0 && (item) < 16) a.trace[((role) * 16 + (item)) * 8 + (ev)] = clock64(); } } while (0)
static long long* g_attn_trace = nullptr;

__device__ __forceinline__ uint32_t psw_off(int row, int col /*0..63*/) {       // SWIZZLE_128B, 128 B rows
  return (uint32_t)(row * 128 + ((((col >> 3) ^ (row & 7)) << 4)) + ((col & 7) << 1));
}
__device__ __forceinline__ void named_bar(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}
__device__ __forceinline__ uint32_t pk2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

// dropout of 8 consecutive elements whose first index is even: one hash per pair (same bits as drop_keep)
__device__ __forceinline__ void drop8_even(const Drop& d, uint64_t base, float* p) {
  const uint32_t i0 = (uint32_t)base >> 1;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const uint32_t h = lowbias32((i0 + k) ^ d.key);
    p[2 * k] = ((h & 0xffffu) >= d.thresh) ? p[2 * k] * d.scale : 0.f;
    p[2 * k + 1] = ((h >> 16) >= d.thresh) ? p[2 * k + 1] * d.scale : 0.f;
  }
}
__device__ __forceinline__ float ex2f(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
static constexpr float LOG2E = 1.4426950408889634f;
static constexpr float LN2 = 0.6931471805599453f;

// barrier slots (8 B each) inside the first KiB of shared memory
enum { PB_FULL = 0, PB_EMPTY = 4, PB_SRDY = 8, PB_PRDY = 11, PB_ORDY = 14, PB_TFREE = 17, PB_SFREE = 20, PB_COUNT = 23 };

// lane 0's probe result for the whole (converged) warp
__device__ __forceinline__ bool warp_probe(uint32_t bar, uint32_t parity) {
  return __shfl_sync(0xffffffffu, mbar_test(bar, parity) ? 1 : 0, 0) != 0;
}

// ------------------------------------------------------------------------------------------------
// forward
// shared memory: [3 KiB: barriers, tmem slot, key bias per group] [nst x (Q | K | V)] [P per group]
//                [O staging per group] [16 KiB read slack: the M = 128 MMAs read rows NP..127 of the P tiles]
// TMEM: group g owns TG columns: S in [0, NP), O in [TO, TO + 64) - separate, so S of the group's next item is
// computed while the current item's P V product and output are still in flight (NG = 3: TG = 160, TO = 96)
// ------------------------------------------------------------------------------------------------
template <int NCH, int NG>      // NP = 16 * NCH keys; NG softmax groups (3 when 3 x (S + O) columns fit in TMEM)
__global__ void __launch_bounds__(64 + 128 * NG, 1)
attn_pipe_fwd_kernel(const __grid_constant__ CUtensorMap tm_qkv, const __grid_constant__ CUtensorMap tm_ctx, const AttnPipeArgs a) {
  constexpr int NP = NCH * 16;
  constexpr int KB = (NP + 63) / 64;
  constexpr uint32_t TN = (uint32_t)NP * 128u;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  constexpr int TG = NG == 3 ? 160 : 256, TO = NG == 3 ? 96 : 128;
  static_assert(NG == 2 || NP <= 96, "three groups need S + O <= 160 columns");
  const int NST = a.nst;
  const uint32_t sStage = base + 3072;
  const uint32_t sP0 = sStage + (uint32_t)NST * 3u * TN;
  const uint32_t sO0 = sP0 + (uint32_t)NG * KB * TN;
  auto bar = [&](int i) { return base + 8u * (uint32_t)i; };
  volatile uint32_t* tslot = reinterpret_cast<volatile uint32_t*>(gen + 512);
  float* kbuf = reinterpret_cast<float*>(gen + 1024);          // key bias (x log2 e) of the current item: [NG][128] fp32

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int S = a.S, H = a.H;
  pdl_trigger();

  if (tid == 0) {
    tma_prefetch_desc(&tm_qkv); tma_prefetch_desc(&tm_ctx);
    for (int s = 0; s < 4; ++s) { mbar_init(bar(PB_FULL + s), 1); mbar_init(bar(PB_EMPTY + s), 1); }
    for (int g = 0; g < NG; ++g) {
      mbar_init(bar(PB_SRDY + g), 1); mbar_init(bar(PB_PRDY + g), 4); mbar_init(bar(PB_ORDY + g), 1); mbar_init(bar(PB_TFREE + g), 4);
      mbar_init(bar(PB_SFREE + g), 4);
    }
    fence_barrier_init(); fence_proxy_async();
  }
  if (warp == 1) tmem_alloc(base + 512, 512);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = *tslot;
  pdl_wait();

  const int n_it = ((int)blockIdx.x < a.items) ? (a.items - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

  if (warp == 0) {
    // ---------------- TMA producer ----------------
    for (int it = 0; it < n_it; ++it) {
      const int w = (int)blockIdx.x + it * (int)gridDim.x;
      const int s = it % NST;
      mbar_wait(bar(PB_EMPTY + s), ((it / NST) & 1) ^ 1u);
      const int b = w / H, h = w % H;
      const int row0 = a.cu ? __ldg(a.cu + b) : b * S;
      if (elect_one()) {
        const uint32_t st = sStage + (uint32_t)s * 3u * TN;
        mbar_expect_tx(bar(PB_FULL + s), 3u * TN);
        tma_load_2d(st, &tm_qkv, bar(PB_FULL + s), h * PD, row0);
        tma_load_2d(st + TN, &tm_qkv, bar(PB_FULL + s), H * PD + h * PD, row0);
        tma_load_2d(st + 2 * TN, &tm_qkv, bar(PB_FULL + s), 2 * H * PD + h * PD, row0);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer: issues whichever of {S(next), PV(next)} has its operands ready, so the two
    // softmax groups never wait on each other ----------------
    auto issue_s = [&](int it) {
      const int s = it % NST, g = it % NG;
      tcgen05_fence_after();
      if (elect_one()) {
        const uint32_t st = sStage + (uint32_t)s * 3u * TN;
        constexpr uint32_t idesc = make_idesc_mn(128, NP, false, false);
        const uint32_t tS = tmem + (uint32_t)(g * TG);
#pragma unroll
        for (int k = 0; k < PD / 16; ++k)
          umma_bf16(tS, umma_desc(st + k * 32, 16, 1024), umma_desc(st + TN + k * 32, 16, 1024), idesc, k > 0);
        umma_commit(bar(PB_SRDY + g));
      }
      __syncwarp();
    };
    auto issue_pv = [&](int it) {
      const int s = it % NST, g = it % NG;
      tcgen05_fence_after();
      if (elect_one()) {
        const uint32_t sV = sStage + (uint32_t)s * 3u * TN + 2 * TN;
        const uint32_t sP = sP0 + (uint32_t)g * KB * TN;
        constexpr uint32_t idesc = make_idesc_mn(128, PD, false, true);
        const uint32_t tO = tmem + (uint32_t)(g * TG + TO);
#pragma unroll
        for (int k = 0; k < NP / 16; ++k)
          umma_bf16(tO, umma_desc(sP + (k >> 2) * TN + (k & 3) * 32, 16, 1024), umma_desc(sV + k * 2048, 8192, 1024), idesc, k > 0);
        umma_commit(bar(PB_ORDY + g));
        umma_commit(bar(PB_EMPTY + s));          // Q, K, V of this stage are consumed
      }
      __syncwarp();
    };
    int ns = 0, npv = 0;
    while (npv < n_it) {
      if (npv < ns) {
        const int g = npv % NG, u = npv / NG;
        if (warp_probe(bar(PB_PRDY + g), u & 1) && warp_probe(bar(PB_TFREE + g), (u & 1) ^ 1u)) { issue_pv(npv); ++npv; continue; }
      }
      if (ns < n_it) {
        const int s = ns % NST, g = ns % NG, u = ns / NG;
        if (warp_probe(bar(PB_FULL + s), (ns / NST) & 1) && warp_probe(bar(PB_SFREE + g), (u & 1) ^ 1u)) { issue_s(ns); ++ns; continue; }
      }
      __nanosleep(64);         // idle poll: leave the issue slots of this sub-partition to its softmax warps
    }
  } else {
    // ---------------- softmax groups ----------------
    const int g = (warp - 2) >> 2;
    const int wq = warp & 3;                       // TMEM lane quarter of this warp
    const int row = wq * 32 + lane;                // query row == TMEM lane
    const int tg = (warp - 2 - 4 * g) * 32 + lane; // thread index inside the group
    const bool in_tile = row < NP;
    float* kb = kbuf + g * 128;                    // key bias * log2(e)
    const uint32_t sP = sP0 + (uint32_t)g * KB * TN;
    uint8_t* gP = gen + (sP - base);
    const uint32_t sO = sO0 + (uint32_t)g * TN;
    uint8_t* gO = gen + (sO - base);
    const uint32_t trow = tmem + (uint32_t)(g * TG) + ((uint32_t)(wq * 32) << 16);
    const float scale2 = a.scale * LOG2E;          // softmax in base 2: p = 2^(s * scale * log2e + kb * log2e - m)
    const bool even_rows = (S & 1) == 0;           // every row's first dropout index is even: pairwise hashing
    // rows of an item: [row0, row0 + Sb) of the (possibly packed) token dimension
    auto item_rows = [&](int it, int& row0, int& Sb) {
      const int b = ((int)blockIdx.x + it * (int)gridDim.x) / H;
      if (a.cu) { row0 = __ldg(a.cu + b); Sb = __ldg(a.cu + b + 1) - row0; } else { row0 = b * S; Sb = S; }
    };
    auto load_kb = [&](int it) -> float {
      if (it >= n_it) return -INFINITY;
      int r0, sb;
      item_rows(it, r0, sb);
      return tg < sb ? __ldg(a.key_bias + (int64_t)r0 + tg) * LOG2E : -INFINITY;
    };
    float kb_next = load_kb(g);
    for (int it = g; it < n_it; it += NG) {
      const int u = it / NG;
      const int w = (int)blockIdx.x + it * (int)gridDim.x;
      const int b = w / H, h = w % H;
      int row0, Sb;
      item_rows(it, row0, Sb);
      const bool valid = row < Sb;
      const bool warp_live = wq * 32 < Sb;
      const float kbv = kb_next;
      kb_next = load_kb(it + NG);                  // in flight during this item
      if (tg == 0) bulk_wait_read<0>();            // the previous TMA store has read the O staging tile
      named_bar(1 + g, 128);                       // every thread of the group is done with the previous item's key bias
      if (tg < NP) kb[tg] = kbv;
      named_bar(1 + g, 128);
      mbar_wait(bar(PB_SRDY + g), u & 1);
      __syncwarp();
      tcgen05_fence_after();
      float m = -INFINITY, sum = 0.f;
      if (warp_live) {
        const uint64_t drow = (((uint64_t)b * H + h) * S + (valid ? row : 0)) * (uint64_t)S;
        float ex[NP];
#pragma unroll
        for (int c = 0; c < NCH; ++c) tmem_ld16_nowait(trow + c * 16, &ex[c * 16]);
#pragma unroll
        for (int c = 0; c < NCH; ++c) tmem_wait_ld_dep16(&ex[c * 16]);
        // the scores are in registers: the S columns may be overwritten by the group's next item
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_relaxed(bar(PB_SFREE + g));
        float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};       // four independent chains
#pragma unroll
        for (int j = 0; j < NP; ++j) { ex[j] = fmaf(ex[j], scale2, kb[j]); m4[j & 3] = fmaxf(m4[j & 3], ex[j]); }
        m = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
        float s4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int j = 0; j < NP; ++j) { ex[j] = ex2f(ex[j] - m); s4[j & 3] += ex[j]; }
        sum = (s4[0] + s4[1]) + (s4[2] + s4[3]);
        const float inv = valid ? 1.0f / sum : 0.f;
        if (in_tile) {
#pragma unroll
          for (int j8 = 0; j8 < NP; j8 += 8) {
            float p[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) p[e] = ex[j8 + e] * inv;
            if (a.drop.on) {
              if (even_rows) {
                drop8_even(a.drop, drow + j8, p);                         // columns >= S hold exact zeros already
              } else {
#pragma unroll
                for (int e = 0; e < 8; ++e) p[e] = drop_apply(a.drop, drow + j8 + e, p[e]);
              }
            }
            *reinterpret_cast<uint4*>(gP + (j8 >> 6) * TN + psw_off(row, j8 & 63)) =
                make_uint4(pk2(p[0], p[1]), pk2(p[2], p[3]), pk2(p[4], p[5]), pk2(p[6], p[7]));
          }
        }
      } else {
        if (lane == 0) mbar_arrive_relaxed(bar(PB_SFREE + g));
        if (in_tile)
          for (int j = 0; j < NP; j += 8)
            *reinterpret_cast<uint4*>(gP + (j >> 6) * TN + psw_off(row, j & 63)) = make_uint4(0, 0, 0, 0);
      }
      fence_proxy_async();
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(PB_PRDY + g));
      if (warp_live && valid && a.lse) a.lse[((int64_t)b * H + h) * S + row] = (m + __log2f(sum)) * LN2;
      mbar_wait(bar(PB_ORDY + g), u & 1);
      __syncwarp();
      tcgen05_fence_after();
      if (warp_live) {
        float o[64];
        tmem_ld32_nowait(trow + TO, o);
        tmem_ld32_nowait(trow + TO + 32, o + 32);
        tmem_wait_ld_dep32(o);
        tmem_wait_ld_dep32(o + 32);
        if (valid) {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            *reinterpret_cast<uint4*>(gO + psw_off(row, 8 * j)) =
                make_uint4(pk2(o[8 * j + 0], o[8 * j + 1]), pk2(o[8 * j + 2], o[8 * j + 3]), pk2(o[8 * j + 4], o[8 * j + 5]),
                           pk2(o[8 * j + 6], o[8 * j + 7]));
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_relaxed(bar(PB_TFREE + g));
      if (!a.cu) {
        fence_proxy_async();
        named_bar(1 + g, 128);
        if (tg == 0) { tma_store_2d(&tm_ctx, sO, h * PD, row0); bulk_commit(); }
      } else {
        // packed layout: the sample owns Sb <= S rows, a fixed-height TMA box would spill into the next sample;
        // the group copies the staged tile out instead, 8 lanes per 128 B row (fully coalesced)
        named_bar(1 + g, 128);
        for (int c = tg; c < Sb * 8; c += 128) {
          const int r = c >> 3, j = c & 7;
          *reinterpret_cast<uint4*>(a.out + (int64_t)(row0 + r) * a.ld_out + h * PD + 8 * j) =
              *reinterpret_cast<const uint4*>(gO + psw_off(r, 8 * j));
        }
      }
    }
    if (tg == 0) bulk_wait_all();
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) { __syncwarp(); tcgen05_fence_after(); tmem_dealloc(tmem, 512); }
}

// ------------------------------------------------------------------------------------------------
// forward, wide: 128 < S <= 160 (X101 features: 40 text + 100 regions; M3P)
// The same persistent pipeline with the unit of work = (sample, head, query tile of 128 rows): K / V boxes hold up to 160
// keys, the scores of a query row (up to 160 fp32) stay in TMEM and are walked twice (online max / sum, then the
// normalised, rounded P), so the register file holds one 32-column chunk at a time.  In the packed layout a sample with
// Sb <= 128 kept rows has one query tile, and the S / PV MMAs only cover the item's own key count.
// shared memory: [3 KiB header] [3 stages x (Q 16 KiB | K 20 KiB | 12 KiB | V 20 KiB)]: the unit's P tile (3 blocks of
//                16 KiB) is written over its own Q | K (dead once S = Q K^T has completed) and the O staging tile over the
//                first P block (free once the PV MMAs have completed), so three units fit and the loads run two ahead
// TMEM: group g owns columns [256 g, 256 g + 256): S in [0, 160), O in [192, 256).  The softmax walks the row three
//       times - max; e = 2^(s - max) written back over S while the sum is formed; p = e / sum rounded to bf16 - so every
//       score costs one exponential and the register file holds one 32-column chunk at a time
// ------------------------------------------------------------------------------------------------
static constexpr int WNP = 160;                          // key capacity
static constexpr uint32_t WTQ = 128u * 128u;             // Q tile / one P block: 128 rows x 128 B
static constexpr uint32_t WTK = (uint32_t)WNP * 128u;    // K / V tile
static constexpr int WKB = 3;                            // P column blocks of 64 keys
static constexpr uint32_t WSTAGE = 69632u;               // Q 16 KiB | K 20 KiB | pad 12 KiB | V 20 KiB  (P = the first 48 KiB)
static constexpr uint32_t WOFF_V = 49152u;

struct WideCursor {            // walks the (item, query tile) units of this CTA in issue order; every role keeps its own
  int it, t, c;                // item index, tile index inside the item, units issued so far
  int b, h, row0, Sb, nt;
  int p_row0, p_end;           // row range of item it + 1, fetched while item `it` is processed (no exposed L2 round trip)
  __device__ __forceinline__ void fetch(const AttnPipeArgs& a, int n_it, int item) {
    if (item >= n_it) return;
    const int bb = ((int)blockIdx.x + item * (int)gridDim.x) / a.H;
    if (a.cu) { p_row0 = __ldg(a.cu + bb); p_end = __ldg(a.cu + bb + 1); } else { p_row0 = bb * a.S; p_end = p_row0 + a.S; }
  }
  __device__ __forceinline__ void adopt(const AttnPipeArgs& a, int n_it) {      // item `it` <- the prefetched range
    if (it >= n_it) return;
    const int w = (int)blockIdx.x + it * (int)gridDim.x;
    b = w / a.H; h = w - b * a.H;
    row0 = p_row0; Sb = p_end - p_row0;
    nt = (Sb + 127) >> 7;
    if (nt < 1) nt = 1;
    fetch(a, n_it, it + 1);
  }
  __device__ __forceinline__ void init(const AttnPipeArgs& a, int n_it) { it = 0; t = 0; c = 0; p_row0 = p_end = 0; fetch(a, n_it, 0); adopt(a, n_it); }
  __device__ __forceinline__ void next(const AttnPipeArgs& a, int n_it) {
    ++c;
    if (++t >= nt) { t = 0; ++it; adopt(a, n_it); }
  }
};

enum { PB_KB = 24 };           // + stage: the producer has written the stage's key-bias row (generic stores)

template <bool TRACE>
__global__ void __launch_bounds__(320, 1)
attn_pipe_fwd_wide_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_kv, const AttnPipeArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  constexpr int NST = 3;
  const uint32_t sStage = base + 3072;
  auto bar = [&](int i) { return base + 8u * (uint32_t)i; };
  volatile uint32_t* tslot = reinterpret_cast<volatile uint32_t*>(gen + 512);
  float* kbuf = reinterpret_cast<float*>(gen + 1024);          // [3][160] key bias x log2(e) of the stage's sample

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int S = a.S, H = a.H;
  pdl_trigger();
  if (tid == 0) {
    tma_prefetch_desc(&tm_q); tma_prefetch_desc(&tm_kv);
    for (int s = 0; s < 4; ++s) { mbar_init(bar(PB_FULL + s), 1); mbar_init(bar(PB_EMPTY + s), 4); mbar_init(bar(PB_KB + s), 1); }
    for (int g = 0; g < 2; ++g) {
      mbar_init(bar(PB_SRDY + g), 1); mbar_init(bar(PB_PRDY + g), 4); mbar_init(bar(PB_ORDY + g), 1); mbar_init(bar(PB_TFREE + g), 4);
      mbar_init(bar(PB_SFREE + g), 4);
    }
    fence_barrier_init(); fence_proxy_async();
  }
  if (warp == 1) tmem_alloc(base + 512, 512);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = *tslot;
  pdl_wait();
  const int n_it = ((int)blockIdx.x < a.items) ? (a.items - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

  if (warp == 0) {
    // ---------------- TMA producer (+ the stage's key-bias row, so that no softmax thread waits for a global load) ------
    WideCursor u;
    for (u.init(a, n_it); u.it < n_it; u.next(a, n_it)) {
      const int s = u.c % NST;
      mbar_wait(bar(PB_EMPTY + s), ((u.c / NST) & 1) ^ 1u);     // the softmax group has copied the stage's O tile out
      if (elect_one()) {
        const uint32_t st = sStage + (uint32_t)s * WSTAGE;
        ATR(0, u.c, 0);
        mbar_expect_tx(bar(PB_FULL + s), WTQ + 2u * WTK);
        tma_load_2d(st, &tm_q, bar(PB_FULL + s), u.h * PD, u.row0 + u.t * 128);
        tma_load_2d(st + WTQ, &tm_kv, bar(PB_FULL + s), H * PD + u.h * PD, u.row0);
        tma_load_2d(st + WOFF_V, &tm_kv, bar(PB_FULL + s), 2 * H * PD + u.h * PD, u.row0);
      }
      float* kb = kbuf + s * WNP;
#pragma unroll
      for (int j = lane; j < WNP; j += 32) kb[j] = j < u.Sb ? __ldg(a.key_bias + (int64_t)u.row0 + j) * LOG2E : -INFINITY;
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(PB_KB + s));       // release: orders the warp's stores (after __syncwarp) before the arrive
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer + scheduler: a unit goes to whichever softmax group is free (its S columns have been
    // consumed), not to a fixed one - a short second query tile no longer waits behind the other group's full tile, and
    // every group's P V is issued the moment its P is ready.  The unit's descriptor travels through shared memory; the
    // group picks it up when its S-ready barrier completes. ----------------
    volatile int* desc = reinterpret_cast<volatile int*>(gen + 2944);        // [2][8]: c, b, h, t, row0, Sb
    WideCursor us;
    us.init(a, n_it);
    int cnt[2] = {0, 0};           // units handed to group g so far
    int pst[2] = {-1, -1};         // stage of group g's unit whose P V has not been issued yet
    int pnk[2] = {0, 0};
    bool fin[2] = {false, false};
    while (!(fin[0] && fin[1])) {
      bool progress = false;
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        if (pst[g] >= 0) {
          const int k = cnt[g] - 1;
          if (warp_probe(bar(PB_PRDY + g), k & 1) && warp_probe(bar(PB_TFREE + g), (k & 1) ^ 1u)) {
            tcgen05_fence_after();
            if (elect_one()) {
              const uint32_t sP = sStage + (uint32_t)pst[g] * WSTAGE;
              const uint32_t sV = sP + WOFF_V;
              constexpr uint32_t idesc = make_idesc_mn(128, PD, false, true);
              const uint32_t tO = tmem + (uint32_t)(g * 256 + 192);
              for (int kk = 0; kk < pnk[g]; ++kk)
                umma_bf16(tO, umma_desc(sP + (kk >> 2) * WTQ + (kk & 3) * 32, 16, 1024), umma_desc(sV + kk * 2048, 8192, 1024), idesc, kk > 0);
              umma_commit(bar(PB_ORDY + g));
              ATR(1, cnt[0] + cnt[1] - 1, 1);
            }
            __syncwarp();
            pst[g] = -1;
            progress = true;
          }
        }
      }
      // next unit (or, when there is none left, the end-of-work mark) for a group whose previous unit is out of the way
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        if (fin[g] || pst[g] >= 0) continue;
        if (!warp_probe(bar(PB_SFREE + g), (cnt[g] & 1) ^ 1u)) continue;       // S columns + descriptor slot of the group are free
        if (us.it >= n_it) {
          if (lane == 0) { desc[g * 8] = -1; __threadfence_block(); mbar_arrive(bar(PB_SRDY + g)); }
          __syncwarp();
          fin[g] = true; progress = true;
          continue;
        }
        const int s = us.c % NST;
        if (!warp_probe(bar(PB_FULL + s), (us.c / NST) & 1)) continue;
        const int ncs = min(WNP, (us.Sb + 31) & ~31);            // score columns the softmax reads (32-column chunks)
        if (lane == 0) {
          volatile int* d = desc + g * 8;
          d[0] = us.c; d[1] = us.b; d[2] = us.h; d[3] = us.t; d[4] = us.row0; d[5] = us.Sb;
          __threadfence_block();
        }
        __syncwarp();
        tcgen05_fence_after();
        if (elect_one()) {
          const uint32_t st = sStage + (uint32_t)s * WSTAGE;
          const uint32_t idesc = make_idesc_mn(128, ncs, false, false);
          const uint32_t tS = tmem + (uint32_t)(g * 256);
#pragma unroll
          for (int k = 0; k < PD / 16; ++k)
            umma_bf16(tS, umma_desc(st + k * 32, 16, 1024), umma_desc(st + WTQ + k * 32, 16, 1024), idesc, k > 0);
          umma_commit(bar(PB_SRDY + g));
          ATR(1, us.c, 0);
        }
        __syncwarp();
        pst[g] = s;
        pnk[g] = min(WNP, (us.Sb + 15) & ~15) >> 4;           // 16-key steps that hold a key of this sample
        ++cnt[g];
        us.next(a, n_it);
        progress = true;
      }
      if (!progress) __nanosleep(64);
    }
  } else {
    // ---------------- softmax groups ----------------
    const int g = (warp - 2) >> 2;
    const int wq = warp & 3;
    const int row = wq * 32 + lane;                // TMEM lane == query row inside the tile
    const int tg = (warp - 2 - 4 * g) * 32 + lane;
    const uint32_t trow = tmem + (uint32_t)(g * 256) + ((uint32_t)(wq * 32) << 16);
    const float scale2 = a.scale * LOG2E;
    const bool even_rows = (S & 1) == 0;
    volatile int* desc = reinterpret_cast<volatile int*>(gen + 2944) + g * 8;
    for (int k = 0;; ++k) {
      if (tg == 0) ATR(2 + g, k, 0);
      mbar_wait(bar(PB_SRDY + g), k & 1);           // S in TMEM: Q and K of the stage are dead, P may overwrite them
      const int uc = desc[0];
      if (uc < 0) break;                            // no unit left for this group
      const int b = desc[1], h = desc[2], ut = desc[3], row0 = desc[4], Sb = desc[5];
      const int s = uc % NST;
      const float* kb = kbuf + s * WNP;
      uint8_t* gP = gen + (sStage - base) + (uint32_t)s * WSTAGE;     // P blocks = the stage's Q | K region
      const int q = ut * 128 + row;                 // query index inside the sample
      const bool valid = q < Sb;
      const bool warp_live = ut * 128 + wq * 32 < Sb;
      const int ncol = min(WNP, (Sb + 31) & ~31);   // == the N of this unit's S MMA
      mbar_wait(bar(PB_KB + s), (uc / NST) & 1);    // the stage's key bias is in shared memory
      __syncwarp();
      if (tg == 0) ATR(2 + g, k, 1);
      tcgen05_fence_after();
      float m = -INFINITY, sum = 0.f;
      if (warp_live) {
        float va[32], vb[32];
        // pass 1: row maximum (base-2 domain); the next chunk's TMEM load is in flight while this one is reduced
        float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
        auto max_chunk = [&](const float* v, int c) {
#pragma unroll
          for (int j = 0; j < 32; ++j) m4[j & 3] = fmaxf(m4[j & 3], fmaf(v[j], scale2, kb[c + j]));
        };
        tmem_ld32_nowait(trow, va);
        for (int c = 0; c < ncol; c += 64) {
          tmem_wait_ld_dep32(va);
          if (c + 32 < ncol) tmem_ld32_nowait(trow + c + 32, vb);
          max_chunk(va, c);
          if (c + 32 < ncol) {
            tmem_wait_ld_dep32(vb);
            if (c + 64 < ncol) tmem_ld32_nowait(trow + c + 64, va);
            max_chunk(vb, c + 32);
          }
        }
        m = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));      // finite: key 0 is never masked
        if (TRACE && tg == 0) { asm volatile("" :: "f"(m)); ATR(2 + g, k, 2); }
        // pass 2: e = 2^(s - m), written back over the scores; row sum
        float s4[4] = {0.f, 0.f, 0.f, 0.f};
        auto exp_chunk = [&](float* v, int c) {
#pragma unroll
          for (int j = 0; j < 32; ++j) { v[j] = ex2f(fmaf(v[j], scale2, kb[c + j]) - m); s4[j & 3] += v[j]; }
          tmem_st32(trow + c, v);
        };
        tmem_ld32_nowait(trow, va);
        for (int c = 0; c < ncol; c += 64) {
          tmem_wait_ld_dep32(va);
          if (c + 32 < ncol) tmem_ld32_nowait(trow + c + 32, vb);
          exp_chunk(va, c);
          if (c + 32 < ncol) {
            tmem_wait_ld_dep32(vb);
            if (c + 64 < ncol) tmem_ld32_nowait(trow + c + 64, va);      // (tcgen05.st took its operands at issue)
            exp_chunk(vb, c + 32);
          }
        }
        tmem_wait_st();
        sum = (s4[0] + s4[1]) + (s4[2] + s4[3]);
        if (TRACE && tg == 0) { asm volatile("" :: "f"(sum)); ATR(2 + g, k, 3); }
        const float inv = valid ? 1.0f / sum : 0.f;
        const uint64_t drow = (((uint64_t)b * H + h) * S + (valid ? q : 0)) * (uint64_t)S;
        // pass 3: normalised probabilities, dropout, bf16 (the oracle's rounding point) -> swizzled P blocks
        auto emit_chunk = [&](float* v, int c) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] *= inv;
#pragma unroll
          for (int j8 = 0; j8 < 32; j8 += 8) {
            if (a.drop.on) {
              if (even_rows) {
                drop8_even(a.drop, drow + c + j8, &v[j8]);
              } else {
#pragma unroll
                for (int e = 0; e < 8; ++e) v[j8 + e] = drop_apply(a.drop, drow + c + j8 + e, v[j8 + e]);
              }
            }
            const int col = c + j8;
            *reinterpret_cast<uint4*>(gP + (col >> 6) * WTQ + psw_off(row, col & 63)) =
                make_uint4(pk2(v[j8 + 0], v[j8 + 1]), pk2(v[j8 + 2], v[j8 + 3]), pk2(v[j8 + 4], v[j8 + 5]), pk2(v[j8 + 6], v[j8 + 7]));
          }
        };
        tmem_ld32_nowait(trow, va);
        for (int c = 0; c < ncol; c += 64) {
          tmem_wait_ld_dep32(va);
          if (c + 32 < ncol) tmem_ld32_nowait(trow + c + 32, vb);
          emit_chunk(va, c);
          if (c + 32 < ncol) {
            tmem_wait_ld_dep32(vb);
            if (c + 64 < ncol) tmem_ld32_nowait(trow + c + 64, va);
            emit_chunk(vb, c + 32);
          }
        }
      }
      // the scores (and the unit's descriptor) have been consumed: the group can be handed its next unit
      tcgen05_fence_before();
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) { mbar_arrive_relaxed(bar(PB_SFREE + g)); mbar_arrive(bar(PB_PRDY + g)); }
      if (tg == 0) ATR(2 + g, k, 4);
      if (warp_live && valid && a.lse) a.lse[((int64_t)b * H + h) * S + q] = (m + __log2f(sum)) * LN2;
      mbar_wait(bar(PB_ORDY + g), k & 1);           // O in TMEM; the P blocks are free again
      __syncwarp();
      if (tg == 0) ATR(2 + g, k, 5);
      tcgen05_fence_after();
      if (warp_live) {
        float o[64];
        tmem_ld32_nowait(trow + 192, o);
        tmem_ld32_nowait(trow + 224, o + 32);
        tmem_wait_ld_dep32(o);
        tmem_wait_ld_dep32(o + 32);
        if (valid) {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            *reinterpret_cast<uint4*>(gP + psw_off(row, 8 * j)) =
                make_uint4(pk2(o[8 * j + 0], o[8 * j + 1]), pk2(o[8 * j + 2], o[8 * j + 3]), pk2(o[8 * j + 4], o[8 * j + 5]),
                           pk2(o[8 * j + 6], o[8 * j + 7]));
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_relaxed(bar(PB_TFREE + g));
      // coalesced copy-out of the staged tile: 8 lanes per 128 B row
      named_bar(1 + g, 128);
      const int nrows = min(128, Sb - ut * 128);
      for (int c = tg; c < nrows * 8; c += 128) {
        const int r = c >> 3, j = c & 7;
        *reinterpret_cast<uint4*>(a.out + (int64_t)(row0 + ut * 128 + r) * a.ld_out + h * PD + 8 * j) =
            *reinterpret_cast<const uint4*>(gP + psw_off(r, 8 * j));
      }
      // this warp's reads of the staging tile have completed before the stage is handed back to the TMA producer
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(PB_EMPTY + s));
      if (tg == 0) ATR(2 + g, k, 6);
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) { __syncwarp(); tcgen05_fence_after(); tmem_dealloc(tmem, 512); }
}

// ------------------------------------------------------------------------------------------------
// backward (S <= 96)
// shared memory: [2 KiB header] [nst x (Q | K | V | dO)] [g=0: P (2 blocks) | dS (2 blocks)] [g=1: ...] [16 KiB slack]
// D_i = sum_j P_ij dP_ij is taken from the probabilities and their gradients themselves (the softmax backward as the
// reference's autograd computes it, in fp32) instead of rowsum(dO o O) over the bf16-rounded context: no O tile to load,
// which leaves room for a third stage at S <= 80
// TMEM: group g owns columns [256 g, 256 g + 256): S [0,128) | dP [128,256), then dV [0,64) | dK [64,128) | dQ [128,192)
// the dQ / dK / dV tiles are staged (bf16, swizzled) in the group's P / dS blocks once the MMAs have consumed them
// ------------------------------------------------------------------------------------------------
struct AttnPipeBwdArgs {
  AttnPipeArgs c;
  int Hd;                                  // H * 64
};

// PDP = true: D_i from P and dP (4 tiles per stage);  PDP = false: D_i = dO_i . O_i from the staged tiles (5 tiles per stage)
template <int NCH, bool PDP>
__global__ void __launch_bounds__(P_THREADS, 1)
attn_pipe_bwd_kernel(const __grid_constant__ CUtensorMap tm_qkv, const __grid_constant__ CUtensorMap tm_do,
                     const __grid_constant__ CUtensorMap tm_o, const __grid_constant__ CUtensorMap tm_dqkv, const AttnPipeBwdArgs ab) {
  constexpr uint32_t NT = PDP ? 4u : 5u;                       // tiles per stage
  constexpr int NP = NCH * 16;
  constexpr uint32_t TN = (uint32_t)NP * 128u;
  const AttnPipeArgs& a = ab.c;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  const int NST = a.nst;
  const uint32_t sStage = base + 3072;
  const uint32_t sG0 = sStage + (uint32_t)NST * NT * TN;       // per group: P (2 TN) | dS (2 TN)
  auto bar = [&](int i) { return base + 8u * (uint32_t)i; };
  volatile uint32_t* tslot = reinterpret_cast<volatile uint32_t*>(gen + 512);
  float* kbuf = reinterpret_cast<float*>(gen + 1024);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int S = a.S, H = a.H;
  pdl_trigger();

  if (tid == 0) {
    tma_prefetch_desc(&tm_qkv); tma_prefetch_desc(&tm_do); tma_prefetch_desc(&tm_o); tma_prefetch_desc(&tm_dqkv);
    for (int s = 0; s < 4; ++s) { mbar_init(bar(PB_FULL + s), 1); mbar_init(bar(PB_EMPTY + s), 1); }
    for (int g = 0; g < 2; ++g) {
      mbar_init(bar(PB_SRDY + g), 1); mbar_init(bar(PB_PRDY + g), 4); mbar_init(bar(PB_ORDY + g), 1); mbar_init(bar(PB_TFREE + g), 4);
    }
    fence_barrier_init(); fence_proxy_async();
  }
  if (warp == 1) tmem_alloc(base + 512, 512);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = *tslot;
  pdl_wait();
  const int n_it = ((int)blockIdx.x < a.items) ? (a.items - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

  if (warp == 0) {
    for (int it = 0; it < n_it; ++it) {
      const int w = (int)blockIdx.x + it * (int)gridDim.x;
      const int s = it % NST;
      mbar_wait(bar(PB_EMPTY + s), ((it / NST) & 1) ^ 1u);
      const int b = w / H, h = w % H;
      const int row0 = a.cu ? __ldg(a.cu + b) : b * S;
      if (elect_one()) {
        const uint32_t st = sStage + (uint32_t)s * NT * TN;
        mbar_expect_tx(bar(PB_FULL + s), NT * TN);
        tma_load_2d(st, &tm_qkv, bar(PB_FULL + s), h * PD, row0);
        tma_load_2d(st + TN, &tm_qkv, bar(PB_FULL + s), H * PD + h * PD, row0);
        tma_load_2d(st + 2 * TN, &tm_qkv, bar(PB_FULL + s), 2 * H * PD + h * PD, row0);
        tma_load_2d(st + 3 * TN, &tm_do, bar(PB_FULL + s), h * PD, row0);
        if (!PDP) tma_load_2d(st + 4 * TN, &tm_o, bar(PB_FULL + s), h * PD, row0);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    auto issue_sdp = [&](int it) {
      const int s = it % NST, g = it & 1;
      tcgen05_fence_after();
      if (elect_one()) {
        const uint32_t sQ = sStage + (uint32_t)s * NT * TN, sK = sQ + TN, sV = sK + TN, sdO = sV + TN;
        constexpr uint32_t idesc = make_idesc_mn(128, NP, false, false);
        const uint32_t tS = tmem + (uint32_t)(g * 256), tdP = tS + 128;
#pragma unroll
        for (int k = 0; k < PD / 16; ++k) {
          umma_bf16(tS, umma_desc(sQ + k * 32, 16, 1024), umma_desc(sK + k * 32, 16, 1024), idesc, k > 0);
          umma_bf16(tdP, umma_desc(sdO + k * 32, 16, 1024), umma_desc(sV + k * 32, 16, 1024), idesc, k > 0);
        }
        umma_commit(bar(PB_SRDY + g));
      }
      __syncwarp();
    };
    auto issue_grad = [&](int it) {
      const int s = it % NST, g = it & 1;
      tcgen05_fence_after();
      if (elect_one()) {
        const uint32_t sQ = sStage + (uint32_t)s * NT * TN, sK = sQ + TN, sdO = sQ + 3 * TN;
        const uint32_t sP = sG0 + (uint32_t)g * 4u * TN, sdS = sP + 2 * TN;
        const uint32_t tdV = tmem + (uint32_t)(g * 256), tdK = tdV + 64, tdQ = tdV + 128;
        // dV[j,d] = sum_i P[i,j] dO[i,d]   A = P^T (MN-major over the [i][j] tile), B = dO (MN-major), K = i
        // dK[j,d] = sum_i dS[i,j] Q[i,d]   A = dS^T (MN-major),                     B = Q  (MN-major)
        constexpr uint32_t id_t = make_idesc_mn(128, PD, true, true);
#pragma unroll
        for (int k = 0; k < NP / 16; ++k) {
          umma_bf16(tdV, umma_desc(sP + k * 2048, TN, 1024), umma_desc(sdO + k * 2048, 8192, 1024), id_t, k > 0);
          umma_bf16(tdK, umma_desc(sdS + k * 2048, TN, 1024), umma_desc(sQ + k * 2048, 8192, 1024), id_t, k > 0);
        }
        // dQ[i,d] = sum_j dS[i,j] K[j,d]   A = dS (K-major), B = K (MN-major), K = j
        constexpr uint32_t id_q = make_idesc_mn(128, PD, false, true);
#pragma unroll
        for (int k = 0; k < NP / 16; ++k)
          umma_bf16(tdQ, umma_desc(sdS + (k >> 2) * TN + (k & 3) * 32, 16, 1024), umma_desc(sK + k * 2048, 8192, 1024), id_q, k > 0);
        umma_commit(bar(PB_ORDY + g));
        umma_commit(bar(PB_EMPTY + s));
      }
      __syncwarp();
    };
    // issue whichever of {S/dP(next), grad(next)} is ready: the two groups' chains stay independent
    int ns = 0, ngr = 0;
    while (ngr < n_it) {
      if (ngr < ns) {
        const int g = ngr & 1, u = ngr >> 1;
        if (warp_probe(bar(PB_PRDY + g), u & 1)) { issue_grad(ngr); ++ngr; continue; }
      }
      if (ns < n_it) {
        const int s = ns % NST, g = ns & 1, u = ns >> 1;
        if (warp_probe(bar(PB_FULL + s), (ns / NST) & 1) && warp_probe(bar(PB_TFREE + g), (u & 1) ^ 1u)) { issue_sdp(ns); ++ns; continue; }
      }
      __nanosleep(64);
    }
  } else {
    const int g = (warp - 2) >> 2;
    const int wq = warp & 3;
    const int row = wq * 32 + lane;
    const int tg = (warp - 2 - 4 * g) * 32 + lane;
    const bool in_tile = row < NP;
    float* kb = kbuf + g * 128;
    const uint32_t sP = sG0 + (uint32_t)g * 4u * TN, sdS = sP + 2 * TN;
    uint8_t* gP = gen + (sP - base);
    uint8_t* gdS = gen + (sdS - base);
    const uint32_t tbase = tmem + (uint32_t)(g * 256) + ((uint32_t)(wq * 32) << 16);
    const float scale = a.scale;
    const float scale2 = a.scale * LOG2E;
    const bool even_rows = (S & 1) == 0;
    auto item_rows = [&](int it, int& row0, int& Sb) {
      const int b = ((int)blockIdx.x + it * (int)gridDim.x) / H;
      if (a.cu) { row0 = __ldg(a.cu + b); Sb = __ldg(a.cu + b + 1) - row0; } else { row0 = b * S; Sb = S; }
    };
    // the row range is fetched two items ahead and the key bias / log-sum-exp one item ahead, so that neither of the two
    // dependent L2 round trips is waited for where it is issued
    int row0_c = 0, Sb_c = 0, row0_n = 0, Sb_n = 0;
    float kb_c = -INFINITY, lse_c = 0.f;
    auto load_vals = [&](int it, int r0, int sb) {
      kb_c = -INFINITY; lse_c = 0.f;
      if (it >= n_it) return;
      const int w = (int)blockIdx.x + it * (int)gridDim.x;
      if (tg < sb) kb_c = __ldg(a.key_bias + (int64_t)r0 + tg) * LOG2E;
      if (row < sb) lse_c = __ldg(a.lse + ((int64_t)w) * S + row);      // w = b * H + h
    };
    if (g < n_it) item_rows(g, row0_c, Sb_c);
    load_vals(g, row0_c, Sb_c);
    if (g + 2 < n_it) item_rows(g + 2, row0_n, Sb_n);
    for (int it = g; it < n_it; it += 2) {
      const int u = it >> 1;
      const int s = it % NST;
      const int w = (int)blockIdx.x + it * (int)gridDim.x;
      const int b = w / H, h = w % H;
      const int row0 = row0_c, Sb = Sb_c;
      const bool valid = row < Sb;
      const bool warp_live = wq * 32 < Sb;         // this warp's 32 rows hold a query of the item
      const int cl = min(NP, (Sb + 15) & ~15);     // key columns [cl, 128) hold no key of the item: exact zeros
      const float kbv = kb_c, lse_i = lse_c;
      load_vals(it + 2, row0_n, Sb_n);             // in flight during this item
      row0_c = row0_n; Sb_c = Sb_n;
      if (it + 4 < n_it) item_rows(it + 4, row0_n, Sb_n);
      if (tg == 0) bulk_wait_read<0>();            // the previous dQ/dK/dV stores have read the staging tiles
      named_bar(1 + g, 128);
      if (tg < 128) kb[tg] = kbv;
      named_bar(1 + g, 128);
      mbar_wait(bar(PB_SRDY + g), u & 1);          // S, dP in TMEM; implies the stage's TMA data has landed
      __syncwarp();
      tcgen05_fence_after();
      if (warp_live) {
       if constexpr (PDP) {
        const uint64_t drow = (((uint64_t)b * H + h) * S + (valid ? row : 0)) * (uint64_t)S;
        // base-2 softmax recompute: p = 2^(s * scale * log2e + kb * log2e - lse * log2e); exact zeros for key columns
        // >= S (kb = -inf) and for query rows >= S (lse2 = +inf), so no per-element predicates are needed
        const float lse2 = valid ? lse_i * LOG2E : INFINITY;
        // pass 1: the row's probabilities stay in registers; D_i = sum_j Pdrop_ij dP_ij (four independent chains)
        float pr[NP];
        float d4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) {
          const int c = ch * 16;
          if (c < cl) {
            float sv[16], dv[16];
            tmem_ld16_nowait(tbase + c, sv);
            tmem_ld16_nowait(tbase + 128 + c, dv);
            tmem_wait_ld_dep16(sv);
            tmem_wait_ld_dep16(dv);
#pragma unroll
            for (int j = 0; j < 16; ++j) pr[c + j] = ex2f(fmaf(sv[j], scale2, kb[c + j]) - lse2);
            if (a.drop.on) {
              if (even_rows) {
                const uint32_t i0 = (uint32_t)(drow + c) >> 1;
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                  const uint32_t hsh = lowbias32((i0 + k) ^ a.drop.key);
                  if ((hsh & 0xffffu) >= a.drop.thresh) d4[(2 * k) & 3] = fmaf(pr[c + 2 * k], dv[2 * k], d4[(2 * k) & 3]);
                  if ((hsh >> 16) >= a.drop.thresh) d4[(2 * k + 1) & 3] = fmaf(pr[c + 2 * k + 1], dv[2 * k + 1], d4[(2 * k + 1) & 3]);
                }
              } else {
#pragma unroll
                for (int j = 0; j < 16; ++j)
                  if (drop_keep(a.drop, drow + c + j)) d4[j & 3] = fmaf(pr[c + j], dv[j], d4[j & 3]);
              }
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j) d4[j & 3] = fmaf(pr[c + j], dv[j], d4[j & 3]);
            }
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) pr[c + j] = 0.f;
          }
        }
        const float Di = ((d4[0] + d4[1]) + (d4[2] + d4[3])) * (a.drop.on ? a.drop.scale : 1.f);
        // pass 2: dS = P (dPdrop - D), Pdrop -> bf16 tiles
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) {
          const int c = ch * 16;
          if (c < cl) {
            float dv[16], pp[16], dd[16];
            tmem_ld16_nowait(tbase + 128 + c, dv);
            tmem_wait_ld_dep16(dv);
#pragma unroll
            for (int j = 0; j < 16; ++j) pp[j] = pr[c + j];
            if (a.drop.on) {
              if (even_rows) {
                const uint32_t i0 = (uint32_t)(drow + c) >> 1;
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                  const uint32_t hsh = lowbias32((i0 + k) ^ a.drop.key);
                  const bool k0 = (hsh & 0xffffu) >= a.drop.thresh, k1 = (hsh >> 16) >= a.drop.thresh;
                  const float d0 = k0 ? dv[2 * k] * a.drop.scale : 0.f, d1 = k1 ? dv[2 * k + 1] * a.drop.scale : 0.f;
                  dd[2 * k] = pp[2 * k] * (d0 - Di); dd[2 * k + 1] = pp[2 * k + 1] * (d1 - Di);
                  pp[2 * k] = k0 ? pp[2 * k] * a.drop.scale : 0.f; pp[2 * k + 1] = k1 ? pp[2 * k + 1] * a.drop.scale : 0.f;
                }
              } else {
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                  const bool keep = drop_keep(a.drop, drow + c + j);
                  const float dp = keep ? dv[j] * a.drop.scale : 0.f;
                  dd[j] = pp[j] * (dp - Di);
                  pp[j] = keep ? pp[j] * a.drop.scale : 0.f;
                }
              }
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j) dd[j] = pp[j] * (dv[j] - Di);
            }
            if (in_tile) {
              const int kblk = c >> 6, cc = c & 63;
              *reinterpret_cast<uint4*>(gP + kblk * TN + psw_off(row, cc)) = make_uint4(pk2(pp[0], pp[1]), pk2(pp[2], pp[3]), pk2(pp[4], pp[5]), pk2(pp[6], pp[7]));
              *reinterpret_cast<uint4*>(gP + kblk * TN + psw_off(row, cc + 8)) = make_uint4(pk2(pp[8], pp[9]), pk2(pp[10], pp[11]), pk2(pp[12], pp[13]), pk2(pp[14], pp[15]));
              *reinterpret_cast<uint4*>(gdS + kblk * TN + psw_off(row, cc)) = make_uint4(pk2(dd[0], dd[1]), pk2(dd[2], dd[3]), pk2(dd[4], dd[5]), pk2(dd[6], dd[7]));
              *reinterpret_cast<uint4*>(gdS + kblk * TN + psw_off(row, cc + 8)) = make_uint4(pk2(dd[8], dd[9]), pk2(dd[10], dd[11]), pk2(dd[12], dd[13]), pk2(dd[14], dd[15]));
            }
          }
        }
       } else {
        // D_i = dO_i . O_i from the staged tiles
        float Di = 0.f;
        if (valid) {
          const uint8_t* gdO = gen + (sStage - base) + (uint32_t)s * NT * TN + 3 * TN;
          const uint8_t* gOo = gdO + TN;
#pragma unroll
          for (int c = 0; c < PD; c += 8) {
            const uint4 x = *reinterpret_cast<const uint4*>(gdO + psw_off(row, c));
            const uint4 y = *reinterpret_cast<const uint4*>(gOo + psw_off(row, c));
            const __nv_bfloat162* xv = reinterpret_cast<const __nv_bfloat162*>(&x);
            const __nv_bfloat162* yv = reinterpret_cast<const __nv_bfloat162*>(&y);
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              const float2 fx = __bfloat1622float2(xv[t]), fy = __bfloat1622float2(yv[t]);
              Di = fmaf(fx.x, fy.x, Di); Di = fmaf(fx.y, fy.y, Di);
            }
          }
        }
        const uint64_t drow = (((uint64_t)b * H + h) * S + (valid ? row : 0)) * (uint64_t)S;
        // base-2 softmax recompute: p = 2^(s * scale * log2e + kb * log2e - lse * log2e); exact zeros for key columns
        // >= S (kb = -inf) and for query rows >= S (lse2 = +inf), so no per-element predicates are needed
        const float lse2 = valid ? lse_i * LOG2E : INFINITY;
#pragma unroll 1
        for (int c = 0; c < cl; c += 16) {
          float sv[16], dv[16];
          tmem_ld16_nowait(tbase + c, sv);
          tmem_ld16_nowait(tbase + 128 + c, dv);
          tmem_wait_ld_dep16(sv);
          tmem_wait_ld_dep16(dv);
          float pp[16], dd[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) pp[j] = ex2f(fmaf(sv[j], scale2, kb[c + j]) - lse2);
          if (a.drop.on) {
            if (even_rows) {
              const uint32_t i0 = (uint32_t)(drow + c) >> 1;
#pragma unroll
              for (int k = 0; k < 8; ++k) {
                const uint32_t hsh = lowbias32((i0 + k) ^ a.drop.key);
                const bool k0 = (hsh & 0xffffu) >= a.drop.thresh, k1 = (hsh >> 16) >= a.drop.thresh;
                const float d0 = k0 ? dv[2 * k] * a.drop.scale : 0.f, d1 = k1 ? dv[2 * k + 1] * a.drop.scale : 0.f;
                dd[2 * k] = pp[2 * k] * (d0 - Di); dd[2 * k + 1] = pp[2 * k + 1] * (d1 - Di);
                pp[2 * k] = k0 ? pp[2 * k] * a.drop.scale : 0.f; pp[2 * k + 1] = k1 ? pp[2 * k + 1] * a.drop.scale : 0.f;
              }
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                const bool keep = drop_keep(a.drop, drow + c + j);
                const float dp = keep ? dv[j] * a.drop.scale : 0.f;
                dd[j] = pp[j] * (dp - Di);
                pp[j] = keep ? pp[j] * a.drop.scale : 0.f;
              }
            }
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) dd[j] = pp[j] * (dv[j] - Di);
          }
          if (in_tile) {
            const int kblk = c >> 6, cc = c & 63;
            *reinterpret_cast<uint4*>(gP + kblk * TN + psw_off(row, cc)) = make_uint4(pk2(pp[0], pp[1]), pk2(pp[2], pp[3]), pk2(pp[4], pp[5]), pk2(pp[6], pp[7]));
            *reinterpret_cast<uint4*>(gP + kblk * TN + psw_off(row, cc + 8)) = make_uint4(pk2(pp[8], pp[9]), pk2(pp[10], pp[11]), pk2(pp[12], pp[13]), pk2(pp[14], pp[15]));
            *reinterpret_cast<uint4*>(gdS + kblk * TN + psw_off(row, cc)) = make_uint4(pk2(dd[0], dd[1]), pk2(dd[2], dd[3]), pk2(dd[4], dd[5]), pk2(dd[6], dd[7]));
            *reinterpret_cast<uint4*>(gdS + kblk * TN + psw_off(row, cc + 8)) = make_uint4(pk2(dd[8], dd[9]), pk2(dd[10], dd[11]), pk2(dd[12], dd[13]), pk2(dd[14], dd[15]));
          }
        }
       }
        if (in_tile) {
          for (int c = cl; c < 128; c += 8) {          // unused key columns are read as M rows of the transposed MMAs
            *reinterpret_cast<uint4*>(gP + (c >> 6) * TN + psw_off(row, c & 63)) = make_uint4(0, 0, 0, 0);
            *reinterpret_cast<uint4*>(gdS + (c >> 6) * TN + psw_off(row, c & 63)) = make_uint4(0, 0, 0, 0);
          }
        }
      } else if (in_tile) {
        // rows past the item's last query (in the packed layout they belong to the next item): P = dS = 0
        for (int c = 0; c < 128; c += 8) {
          *reinterpret_cast<uint4*>(gP + (c >> 6) * TN + psw_off(row, c & 63)) = make_uint4(0, 0, 0, 0);
          *reinterpret_cast<uint4*>(gdS + (c >> 6) * TN + psw_off(row, c & 63)) = make_uint4(0, 0, 0, 0);
        }
      }
      fence_proxy_async();
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(PB_PRDY + g));
      mbar_wait(bar(PB_ORDY + g), u & 1);          // dV, dK, dQ in TMEM; P / dS tiles are free again
      __syncwarp();
      tcgen05_fence_after();
      if (wq * 32 < Sb) {
        // staging: dQ -> P block 0, dK -> P block 1, dV -> dS block 0 (S rows x 128 B each)
        uint8_t* stg[3] = {gP, gP + TN, gdS};
        const uint32_t src[3] = {tbase + 128, tbase + 64, tbase};
        const float mul[3] = {scale, scale, 1.f};
#pragma unroll
        for (int which = 0; which < 3; ++which) {
          float o[64];
          tmem_ld32_nowait(src[which], o);
          tmem_ld32_nowait(src[which] + 32, o + 32);
          tmem_wait_ld_dep32(o);
          tmem_wait_ld_dep32(o + 32);
          if (valid) {
            const float sc = mul[which];
#pragma unroll
            for (int j = 0; j < 8; ++j)
              *reinterpret_cast<uint4*>(stg[which] + psw_off(row, 8 * j)) =
                  make_uint4(pk2(o[8 * j + 0] * sc, o[8 * j + 1] * sc), pk2(o[8 * j + 2] * sc, o[8 * j + 3] * sc),
                             pk2(o[8 * j + 4] * sc, o[8 * j + 5] * sc), pk2(o[8 * j + 6] * sc, o[8 * j + 7] * sc));
          }
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_relaxed(bar(PB_TFREE + g));
      if (!a.cu) {
        fence_proxy_async();
        named_bar(1 + g, 128);
        if (tg == 0) {
          tma_store_2d(&tm_dqkv, sP, h * PD, row0);
          tma_store_2d(&tm_dqkv, sP + TN, ab.Hd + h * PD, row0);
          tma_store_2d(&tm_dqkv, sdS, 2 * ab.Hd + h * PD, row0);
          bulk_commit();
        }
      } else {
        // packed layout (see the forward): coalesced copy-out of the three staged tiles, 8 lanes per 128 B row
        named_bar(1 + g, 128);
        const uint8_t* tiles[3] = {gP, gP + TN, gdS};
#pragma unroll
        for (int which = 0; which < 3; ++which)
          for (int c = tg; c < Sb * 8; c += 128) {
            const int r = c >> 3, j = c & 7;
            *reinterpret_cast<uint4*>(a.out + (int64_t)(row0 + r) * a.ld_out + which * ab.Hd + h * PD + 8 * j) =
                *reinterpret_cast<const uint4*>(tiles[which] + psw_off(r, 8 * j));
          }
      }
    }
    if (tg == 0) bulk_wait_all();
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) { __syncwarp(); tcgen05_fence_after(); tmem_dealloc(tmem, 512); }
}

// ------------------------------------------------------------------------------------------------
// host
// ------------------------------------------------------------------------------------------------
static int make_tmap_rows(CUtensorMap* m, const void* ptr, int64_t cols, int64_t rows, int64_t ld, int box_rows) {
  return make_tmap_bf16(m, ptr, cols, rows, ld, box_rows);        // box = 64 columns x box_rows, 128B swizzle
}

static bool pipe_enabled() {
  static const bool v = [] { const char* s = getenv("CLG_ATTN_PIPE"); return !(s && s[0] == '0'); }();
  return v;
}

static int fwd_pipe_smem(int NP, int nst, int ng) {
  const int KB = (NP + 63) / 64, TN = NP * 128;
  return 1024 + 3072 + nst * 3 * TN + ng * KB * TN + ng * TN + 16384;
}
static int bwd_pipe_smem(int NP, int nst, bool pdp) {
  const int TN = NP * 128;
  return 1024 + 3072 + nst * (pdp ? 4 : 5) * TN + 8 * TN + 16384;
}
static constexpr int P_SMEM_MAX = 232448;

template <int NCH>
static int launch_pipe_fwd(const CUtensorMap& tq, const CUtensorMap& tc, const AttnPipeArgs& a, int smem, cudaStream_t stream) {
  constexpr int NG = NCH <= 6 ? 3 : 2;
  static std::once_flag once;
  static cudaError_t err = cudaSuccess;
  std::call_once(once, [] { err = cudaFuncSetAttribute(attn_pipe_fwd_kernel<NCH, NG>, cudaFuncAttributeMaxDynamicSharedMemorySize, P_SMEM_MAX); });
  CLG_CUDA_OK(err, "cudaFuncSetAttribute(attn_pipe_fwd)");
  const int grid = a.items < num_sms() ? a.items : num_sms();
  CLG_CUDA_OK(launch_pdl(attn_pipe_fwd_kernel<NCH, NG>, dim3(grid), dim3(64 + 128 * NG), (size_t)smem, stream, tq, tc, a), "launch(attn_pipe_fwd)");
  return check_launch("attn_pipe_fwd_kernel");
}

template <int NCH, bool PDP>
static int launch_pipe_bwd_v(const CUtensorMap& tq, const CUtensorMap& tdo, const CUtensorMap& to, const CUtensorMap& tdq,
                             const AttnPipeBwdArgs& a, int smem, cudaStream_t stream) {
  static std::once_flag once;
  static cudaError_t err = cudaSuccess;
  std::call_once(once, [] { err = cudaFuncSetAttribute(attn_pipe_bwd_kernel<NCH, PDP>, cudaFuncAttributeMaxDynamicSharedMemorySize, P_SMEM_MAX); });
  CLG_CUDA_OK(err, "cudaFuncSetAttribute(attn_pipe_bwd)");
  const int grid = a.c.items < num_sms() ? a.c.items : num_sms();
  CLG_CUDA_OK(launch_pdl(attn_pipe_bwd_kernel<NCH, PDP>, dim3(grid), dim3(P_THREADS), (size_t)smem, stream, tq, tdo, to, tdq, a), "launch(attn_pipe_bwd)");
  return check_launch("attn_pipe_bwd_kernel");
}
template <int NCH>
static int launch_pipe_bwd(bool pdp, const CUtensorMap& tq, const CUtensorMap& tdo, const CUtensorMap& to, const CUtensorMap& tdq,
                           const AttnPipeBwdArgs& a, int smem, cudaStream_t stream) {
  return pdp ? launch_pipe_bwd_v<NCH, true>(tq, tdo, to, tdq, a, smem, stream) : launch_pipe_bwd_v<NCH, false>(tq, tdo, to, tdq, a, smem, stream);
}

static int attention_pipe_fwd_wide(const clgAttentionFwdParams& p, cudaStream_t stream) {
  static const bool on = [] { const char* s = getenv("CLG_ATTN_WIDE"); return !(s && s[0] == '0'); }();
  if (!on) return CLG_ERR_UNSUPPORTED;
  if ((p.ld_ctx * 2) % 16 != 0 || (reinterpret_cast<uintptr_t>(p.ctx) & 15)) return CLG_ERR_UNSUPPORTED;
  if (p.cu_rows && p.total_rows <= 0) return CLG_ERR_ARG;
  const int nst = 3;
  const int smem = 1024 + 3072 + nst * (int)WSTAGE;
  if (smem > P_SMEM_MAX) return CLG_ERR_UNSUPPORTED;
  const int64_t rows = p.cu_rows ? (int64_t)p.total_rows : (int64_t)p.B * p.S;   // rows beyond are zero-filled by TMA
  CUtensorMap tq, tkv;
  int st = make_tmap_rows(&tq, p.qkv, 3LL * p.H * PD, rows, p.ld_qkv, 128);
  if (st != CLG_OK) return st;
  st = make_tmap_rows(&tkv, p.qkv, 3LL * p.H * PD, rows, p.ld_qkv, WNP);
  if (st != CLG_OK) return st;
  AttnPipeArgs a;
  memset(&a, 0, sizeof(a));
  a.B = p.B; a.S = p.S; a.H = p.H; a.NP = WNP; a.items = p.B * p.H; a.nst = nst;
  a.scale = p.scale; a.key_bias = p.key_bias; a.lse = p.lse;
  a.drop = make_drop(p.dropout);
  a.cu = p.cu_rows; a.out = (bf16*)p.ctx; a.ld_out = p.ld_ctx;
  static std::once_flag once;
  static cudaError_t err = cudaSuccess;
  std::call_once(once, [] {
    err = cudaFuncSetAttribute(attn_pipe_fwd_wide_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, P_SMEM_MAX);
    if (err == cudaSuccess) err = cudaFuncSetAttribute(attn_pipe_fwd_wide_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, P_SMEM_MAX);
  });
  CLG_CUDA_OK(err, "cudaFuncSetAttribute(attn_pipe_fwd_wide)");
  const int grid = a.items < num_sms() ? a.items : num_sms();
  a.trace = g_attn_trace;
  if (a.trace) CLG_CUDA_OK(launch_pdl(attn_pipe_fwd_wide_kernel<true>, dim3(grid), dim3(320), (size_t)smem, stream, tq, tkv, a), "launch(attn_pipe_fwd_wide)");
  else CLG_CUDA_OK(launch_pdl(attn_pipe_fwd_wide_kernel<false>, dim3(grid), dim3(320), (size_t)smem, stream, tq, tkv, a), "launch(attn_pipe_fwd_wide)");
  return check_launch("attn_pipe_fwd_wide_kernel");
}

// returns CLG_ERR_UNSUPPORTED when the shape is outside the pipelined kernels' range (caller falls back)
int attention_pipe_fwd(const clgAttentionFwdParams& p, cudaStream_t stream) {
  if (pipe_enabled() && p.dtype == CLG_BF16 && p.S > 128 && p.S <= WNP) return attention_pipe_fwd_wide(p, stream);
  if (!pipe_enabled() || p.dtype != CLG_BF16 || p.S > 128 || p.S < 1) return CLG_ERR_UNSUPPORTED;
  const int NP = (p.S + 15) / 16 * 16;
  const int ng = NP <= 96 ? 3 : 2;
  int nst = 3;
  while (nst > 2 && fwd_pipe_smem(NP, nst, ng) > P_SMEM_MAX) --nst;
  if (fwd_pipe_smem(NP, nst, ng) > P_SMEM_MAX) return CLG_ERR_UNSUPPORTED;
  if ((p.ld_ctx * 2) % 16 != 0 || (reinterpret_cast<uintptr_t>(p.ctx) & 15)) return CLG_ERR_UNSUPPORTED;
  CUtensorMap tq, tc;
  if (p.cu_rows && p.total_rows <= 0) return CLG_ERR_ARG;
  const int64_t rows = p.cu_rows ? (int64_t)p.total_rows : (int64_t)p.B * p.S;   // rows beyond are zero-filled by TMA
  int st = make_tmap_rows(&tq, p.qkv, 3LL * p.H * PD, rows, p.ld_qkv, NP);
  if (st != CLG_OK) return st;
  st = make_tmap_rows(&tc, p.ctx, (int64_t)p.H * PD, rows, p.ld_ctx, p.S);
  if (st != CLG_OK) return st;
  AttnPipeArgs a;
  memset(&a, 0, sizeof(a));
  a.B = p.B; a.S = p.S; a.H = p.H; a.NP = NP; a.items = p.B * p.H; a.nst = nst;
  a.scale = p.scale; a.key_bias = p.key_bias; a.lse = p.lse;
  a.drop = make_drop(p.dropout);
  a.cu = p.cu_rows; a.out = (bf16*)p.ctx; a.ld_out = p.ld_ctx;
  const int smem = fwd_pipe_smem(NP, nst, ng);
  switch (NP / 16) {
    case 1: return launch_pipe_fwd<1>(tq, tc, a, smem, stream);
    case 2: return launch_pipe_fwd<2>(tq, tc, a, smem, stream);
    case 3: return launch_pipe_fwd<3>(tq, tc, a, smem, stream);
    case 4: return launch_pipe_fwd<4>(tq, tc, a, smem, stream);
    case 5: return launch_pipe_fwd<5>(tq, tc, a, smem, stream);
    case 6: return launch_pipe_fwd<6>(tq, tc, a, smem, stream);
    case 7: return launch_pipe_fwd<7>(tq, tc, a, smem, stream);
    case 8: return launch_pipe_fwd<8>(tq, tc, a, smem, stream);
    default: return CLG_ERR_UNSUPPORTED;
  }
}

int attention_pipe_bwd(const clgAttentionBwdParams& p, cudaStream_t stream) {
  if (!pipe_enabled() || p.dtype != CLG_BF16 || p.S > 96 || p.S < 1) return CLG_ERR_UNSUPPORTED;
  const int NP = (p.S + 15) / 16 * 16;
  // CLG_ATTN_BWD_PDP=1: D_i from P and dP (the reference's own softmax-backward formula, no O tile, a third stage where it
  // fits) - measured slower (two passes over the scores); default: D_i = dO . O from the staged bf16 tiles
  static const bool pdp = [] { const char* s = getenv("CLG_ATTN_BWD_PDP"); return s && s[0] == '1'; }();
  int nst = 3;                                                        // >= 2 stages (see the forward)
  while (nst > 2 && bwd_pipe_smem(NP, nst, pdp) > P_SMEM_MAX) --nst;
  if (bwd_pipe_smem(NP, nst, pdp) > P_SMEM_MAX) return CLG_ERR_UNSUPPORTED;
  if ((p.ld_dqkv * 2) % 16 != 0 || (reinterpret_cast<uintptr_t>(p.dqkv) & 15)) return CLG_ERR_UNSUPPORTED;
  CUtensorMap tq, tdo, to, tdq;
  if (p.cu_rows && p.total_rows <= 0) return CLG_ERR_ARG;
  const int64_t rows = p.cu_rows ? (int64_t)p.total_rows : (int64_t)p.B * p.S;
  int st = make_tmap_rows(&tq, p.qkv, 3LL * p.H * PD, rows, p.ld_qkv, NP);
  if (st != CLG_OK) return st;
  st = make_tmap_rows(&tdo, p.dctx, (int64_t)p.H * PD, rows, p.ld_dctx, NP);
  if (st != CLG_OK) return st;
  st = make_tmap_rows(&to, p.ctx, (int64_t)p.H * PD, rows, p.ld_ctx, NP);
  if (st != CLG_OK) return st;
  st = make_tmap_rows(&tdq, p.dqkv, 3LL * p.H * PD, rows, p.ld_dqkv, p.S);
  if (st != CLG_OK) return st;
  AttnPipeBwdArgs ab;
  memset(&ab, 0, sizeof(ab));
  AttnPipeArgs& a = ab.c;
  a.B = p.B; a.S = p.S; a.H = p.H; a.NP = NP; a.items = p.B * p.H; a.nst = nst;
  a.scale = p.scale; a.key_bias = p.key_bias; a.lse = const_cast<float*>(p.lse);
  a.drop = make_drop(p.dropout);
  a.cu = p.cu_rows; a.out = (bf16*)p.dqkv; a.ld_out = p.ld_dqkv;
  ab.Hd = p.H * PD;
  const int smem = bwd_pipe_smem(NP, nst, pdp);
  switch (NP / 16) {
    case 1: return launch_pipe_bwd<1>(pdp, tq, tdo, to, tdq, ab, smem, stream);
    case 2: return launch_pipe_bwd<2>(pdp, tq, tdo, to, tdq, ab, smem, stream);
    case 3: return launch_pipe_bwd<3>(pdp, tq, tdo, to, tdq, ab, smem, stream);
    case 4: return launch_pipe_bwd<4>(pdp, tq, tdo, to, tdq, ab, smem, stream);
    case 5: return launch_pipe_bwd<5>(pdp, tq, tdo, to, tdq, ab, smem, stream);
    case 6: return launch_pipe_bwd<6>(pdp, tq, tdo, to, tdq, ab, smem, stream);
    default: return CLG_ERR_UNSUPPORTED;
  }
}

}  // namespace clg

// debug hook (not part of the C ABI in include/clgvqa.h): device buffer of 5 * 16 * 8 int64 that the next pipelined
// attention forward launches fill with clock64 stamps of CTA 0 (tools/attn_trace.py); NULL switches it off
extern "C" int clg_debug_attn_trace(void* buf) { clg::g_attn_trace = static_cast<long long*>(buf); return 0; }
