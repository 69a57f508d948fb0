// tcgen05 / TMEM / TMA GEMM for sm_100a:  C[M,N] = epi( A(m,k) * B(n,k) ), bf16 operands,
// fp32 accumulation in tensor memory.
//
// Persistent, warp-specialised: warp 0 = TMA producer, warp 1 = MMA issuer (one elected lane issues
// tcgen05.mma; the warp also owns the TMEM allocation), warps 2-5 = epilogue (tcgen05.ld -> registers
// -> fused epilogue -> global).  Two accumulator stages in TMEM (2 x BN columns) let the epilogue of
// tile i overlap the MMAs of tile i+1; a KSTAGES-deep smem ring (128B-swizzled TMA boxes) feeds the
// tensor core.  Both operands may be K-major (forward / dgrad through a transposed copy) or MN-major
// (dgrad on W[N,K] as stored, wgrad on dY / X as stored) - selected by the UMMA descriptors, no
// transposes in memory.  Optional split-K (atomic fp32 accumulate) for the wgrad shapes whose output
// has too few tiles to fill 148 SMs.
#include <mutex>
#include <stdlib.h>
#include "gemm_epilogue_tc.cuh"
#include "gemm_epilogue_fast.cuh"
#include "tc_common.cuh"

namespace clg {

static constexpr int BM = 128;
static constexpr int BK = 64;            // 64 bf16 = 128 B = one swizzle row
static constexpr int UMMA_K = 16;
// warps: TMA producer, MMA issuer, then the epilogue warps (8; 12 for the ALU-heavy GeLU-pair epilogue, whose per-tile
// work must fit under a K = 768 main loop)
// epilogue warps per CTA (multiples of 4: one set per TMEM lane quarter); the math-heavy forms get more
#ifndef CLG_GELU_EPI_WARPS
#define CLG_GELU_EPI_WARPS 12
#endif
#ifndef CLG_MULAUX_EPI_WARPS
#define CLG_MULAUX_EPI_WARPS 8
#endif
template <int EPI> __host__ __device__ constexpr int epi_warps() { return EPI == EPI_GELU_PAIR ? CLG_GELU_EPI_WARPS : (EPI == EPI_MUL_AUX ? CLG_MULAUX_EPI_WARPS : 8); }
template <int EPI> __host__ __device__ constexpr int num_threads() { return 64 + 32 * epi_warps<EPI>(); }
static constexpr int A_TILE_BYTES = BM * BK * 2;   // 16 KiB

static constexpr int SMEM_MAX = 232448;            // 227 KiB per CTA
static constexpr int BAR_BYTES = 512;

// MC (pairing mode): 0 = one CTA per tile; 1 = CTA pair, each computes its own 128 x BN tile and the shared B tile is
// TMA-multicast; 2 = CTA pair issuing ONE tcgen05.mma.cta_group::2 (M = 256): every CTA stages its 128 A rows and only
// HALF of the B tile - per-SM shared-memory fill per FLOP halves, the ring gets deeper, latency tolerance doubles.
template <int BN, int EPI, int MC = 0> struct TileCfg {
  static constexpr int B_TILE_BYTES = (MC == 2 ? BN / 2 : BN) * BK * 2;
  static constexpr int STAGE_BYTES = A_TILE_BYTES + B_TILE_BYTES;
  // epilogue scratch: generic = one [32][33] fp32 staging tile per warp; fast = per-warp TMA boxes
  static constexpr int EPI_BYTES = EPI == EPI_GENERIC ? 8 * STG_FLOATS * 4 : epi_warps<EPI>() * epi_warp_bytes<EPI>();
  static constexpr int FIT = (SMEM_MAX - 1024 - BAR_BYTES - EPI_BYTES) / STAGE_BYTES;
  static constexpr int KSTAGES = FIT > 8 ? 8 : FIT;
  static constexpr int TMEM_COLS = 2 * BN;          // power of two >= 32 for BN in {64,128,256}
  static constexpr int SMEM_BYTES = KSTAGES * STAGE_BYTES + 1024 /*align slack*/ + BAR_BYTES + EPI_BYTES;
};

__host__ __device__ constexpr uint32_t make_idesc(int n, bool a_mn, bool b_mn) { return make_idesc_mn(BM, n, a_mn, b_mn); }

struct TcArgs {
  int M, N, K;
  int num_m, num_n, splits, kper, kblocks;
  int staged;      // epilogue may use the coalesced shared-memory staged path
  int total_units; // work units: (m block | pair of m blocks when MC) x n block x split
  int m_split_blk; // first 128-row block of the second row group (0: no second group)
};

// MC: clusters of two CTAs work on two vertically adjacent 128-row tiles of the same N block; each CTA loads
// half of the shared B tile and TMA-multicasts it into both CTAs' shared memory (A 16 KiB + B/2 per k-block from
// L2 instead of A + B), so the L2 -> SM traffic of the weight / activation tile is halved.
template <int BN, bool A_K, bool B_K, int MC, int EPI>
__global__ void __launch_bounds__(num_threads<EPI>(), 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
               const __grid_constant__ CUtensorMap tma_c, const __grid_constant__ CUtensorMap tma_x,
               const __grid_constant__ CUtensorMap tma_a2, const __grid_constant__ CUtensorMap tma_b2,
               const TcArgs t, const EpiArgs e) {
  using Cfg = TileCfg<BN, EPI, MC>;
  constexpr int KSTAGES = Cfg::KSTAGES;
  static_assert(KSTAGES >= 3, "pipeline too shallow");
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  // layout: [KSTAGES x (A | B)] [epilogue scratch] [barriers]
  const uint32_t epi_base = smem_base + KSTAGES * Cfg::STAGE_BYTES;
  const uint32_t bar_base = epi_base + Cfg::EPI_BYTES;
  // barrier layout (8 B each): full[KSTAGES], empty[KSTAGES], tmem_full[2], tmem_empty[2], tmem_ptr, epi_in[8][2]
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (KSTAGES + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * KSTAGES + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * KSTAGES + 2 + s); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * KSTAGES + 4);
  const uint32_t epi_in_bar = bar_base + 8u * (2 * KSTAGES + 5);
  static_assert(8 * (2 * KSTAGES + 5 + 2 * 12) <= BAR_BYTES, "barrier area too small");
  volatile uint32_t* tmem_slot_gen =
      reinterpret_cast<volatile uint32_t*>(smem_gen + KSTAGES * Cfg::STAGE_BYTES + Cfg::EPI_BYTES + 8 * (2 * KSTAGES + 4));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  pdl_trigger();

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tma_a)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tma_b)) : "memory");
    if (t.m_split_blk > 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tma_a2)) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tma_b2)) : "memory");
    }
    for (int s = 0; s < KSTAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), MC == 1 ? 2 : 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(tfull_bar(s), 1); mbar_init(tempty_bar(s), (MC == 2 ? 2 : 1) * epi_warps<EPI>()); }
    for (int s = 0; s < 2 * epi_warps<EPI>(); ++s) mbar_init(epi_in_bar + 8u * s, 1);
    if (EPI != EPI_GENERIC) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tma_c)) : "memory");
      if (EPI != EPI_PLAIN) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tma_x)) : "memory");
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    if (MC == 2) tmem_alloc_2cta(tmem_slot, (uint32_t)Cfg::TMEM_COLS);
    else tmem_alloc(tmem_slot, (uint32_t)Cfg::TMEM_COLS);
  }
  tcgen05_fence_before();
  __syncthreads();
  if (MC) cluster_sync_all();          // the peer's barriers must be initialised before any multicast arrives
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;
  pdl_wait();                          // prologue done; everything below touches memory of the preceding kernels

  const int crank = MC ? (int)cluster_ctarank() : 0;
  const int total_work = t.total_units;
  const int w_first = MC ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int w_step = MC ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  auto m_block_of = [&](int tile) { const int mi = tile / t.num_n; return MC ? 2 * mi + crank : mi; };

  if (warp == 0) {
    // TMA producer: the whole warp walks the schedule (uniform control flow), one elected lane issues
    int stage = 0; uint32_t phase = 0;
    for (int w = w_first; w < total_work; w += w_step) {
      const int split = w % t.splits;
      const int tile = w / t.splits;
      const int n_blk = tile % t.num_n;
      int m_blk = m_block_of(tile);
      // second row group: its own operands, rows counted from its first block
      const bool g2 = t.m_split_blk > 0 && m_blk >= t.m_split_blk;
      const CUtensorMap* pa = g2 ? &tma_a2 : &tma_a;
      const CUtensorMap* pb = g2 ? &tma_b2 : &tma_b;
      if (g2) m_blk -= t.m_split_blk;
      const int kb0 = split * t.kper;
      const int kb1 = min(t.kblocks, kb0 + t.kper);
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(empty_bar(stage), phase ^ 1u);
        if (MC == 2) {
          // pair mode: both CTAs' boxes complete on the LEADER's full barrier (the leader issues the MMA for both)
          if (elect_one()) {
            const uint32_t fb = mapa_shared(full_bar(stage), 0);
            if (crank == 0) mbar_expect_tx(full_bar(stage), 2 * Cfg::STAGE_BYTES);
            const uint32_t sa = smem_base + stage * Cfg::STAGE_BYTES;
            const uint32_t sb = sa + A_TILE_BYTES;
            if (A_K) {
              tma_load_2d_2cta(sa, pa, fb, kb * BK, m_blk * BM);
            } else {
#pragma unroll
              for (int j = 0; j < BM / 64; ++j)
                tma_load_2d_2cta(sa + j * (64 * BK * 2), pa, fb, m_blk * BM + j * 64, kb * BK);
            }
            if (B_K) {
              tma_load_2d_2cta(sb, pb, fb, kb * BK, n_blk * BN + crank * (BN / 2));
            } else {
#pragma unroll
              for (int j = 0; j < BN / 128; ++j)
                tma_load_2d_2cta(sb + j * (64 * BK * 2), pb, fb, n_blk * BN + (crank * (BN / 128) + j) * 64, kb * BK);
            }
          }
        } else if (elect_one()) {
          mbar_expect_tx(full_bar(stage), Cfg::STAGE_BYTES);
          const uint32_t sa = smem_base + stage * Cfg::STAGE_BYTES;
          const uint32_t sb = sa + A_TILE_BYTES;
          if (A_K) {
            tma_load_2d(sa, pa, full_bar(stage), kb * BK, m_blk * BM);
          } else {
#pragma unroll
            for (int j = 0; j < BM / 64; ++j)
              tma_load_2d(sa + j * (64 * BK * 2), pa, full_bar(stage), m_blk * BM + j * 64, kb * BK);
          }
          if (MC == 1) {
            // this CTA fetches its half of the B tile and multicasts it to both CTAs of the pair
            if (B_K) {
              tma_load_2d_mc(sb + crank * (BN / 2) * 128, pb, full_bar(stage), kb * BK, n_blk * BN + crank * (BN / 2), 3);
            } else {
#pragma unroll
              for (int j = 0; j < BN / 128; ++j) {
                const int jj = crank * (BN / 128) + j;
                tma_load_2d_mc(sb + jj * (64 * BK * 2), pb, full_bar(stage), n_blk * BN + jj * 64, kb * BK, 3);
              }
            }
          } else if (B_K) {
            tma_load_2d(sb, pb, full_bar(stage), kb * BK, n_blk * BN);
          } else {
#pragma unroll
            for (int j = 0; j < BN / 64; ++j)
              tma_load_2d(sb + j * (64 * BK * 2), pb, full_bar(stage), n_blk * BN + j * 64, kb * BK);
          }
        }
        __syncwarp();
        if (++stage == KSTAGES) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // MMA issuer (pair mode: the leader CTA only): uniform control flow, one elected lane issues tcgen05.mma / commit
    if (MC != 2 || crank == 0) {
    constexpr uint32_t idesc = MC == 2 ? make_idesc_mn(2 * BM, BN, !A_K, !B_K) : make_idesc(BN, !A_K, !B_K);
    int stage = 0; uint32_t phase = 0;
    int local = 0;
    for (int w = w_first; w < total_work; w += w_step, ++local) {
      const int split = w % t.splits;
      const int kb0 = split * t.kper;
      const int kb1 = min(t.kblocks, kb0 + t.kper);
      const int as = local & 1;
      const uint32_t aphase = (local >> 1) & 1;
      mbar_wait(tempty_bar(as), aphase ^ 1u);
      tcgen05_fence_after();
      const uint32_t tmem_d = tmem_base + (uint32_t)(as * BN);
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(full_bar(stage), phase);
        tcgen05_fence_after();
        if (elect_one()) {
          const uint32_t sa = smem_base + stage * Cfg::STAGE_BYTES;
          const uint32_t sb = sa + A_TILE_BYTES;
          // K-major: 8-row groups 1024 B apart (SBO), advance 32 B per UMMA_K inside the swizzle row.
          // MN-major: 64-wide MN blocks 8192 B apart (LBO), 8-k groups 1024 B apart (SBO),
          //           advance 16 k-rows = 2048 B per UMMA_K.
          const uint64_t adesc0 = A_K ? umma_desc(sa, 16, 1024) : umma_desc(sa, 64 * BK * 2, 1024);
          const uint64_t bdesc0 = B_K ? umma_desc(sb, 16, 1024) : umma_desc(sb, 64 * BK * 2, 1024);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            // the start-address field (bits 0-13, 16 B units) never carries out of its field inside one stage
            const uint64_t adesc = adesc0 + (uint64_t)((A_K ? k * (UMMA_K * 2) : k * (UMMA_K * 128)) >> 4);
            const uint64_t bdesc = bdesc0 + (uint64_t)((B_K ? k * (UMMA_K * 2) : k * (UMMA_K * 128)) >> 4);
            if (MC == 2) umma_bf16_2cta(tmem_d, adesc, bdesc, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
            else umma_bf16(tmem_d, adesc, bdesc, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          // frees the smem slot once these MMAs have read it (in both CTAs of a pair: multicast mode - the peer's
          // next TMA writes into this CTA's slot as well; cta_group::2 - the MMA read the peer's slot too)
          if (MC == 2) umma_commit_2cta(empty_bar(stage), 3);
          else if (MC == 1) umma_commit_mc(empty_bar(stage), 3);
          else umma_commit(empty_bar(stage));
          if (kb == kb1 - 1) {                                  // accumulator complete -> epilogue (of both CTAs)
            if (MC == 2) umma_commit_2cta(tfull_bar(as), 3); else umma_commit(tfull_bar(as));
          }
        }
        __syncwarp();
        if (++stage == KSTAGES) { stage = 0; phase ^= 1u; }
      }
    }
    }
  } else {
    // 8 epilogue warps: two per TMEM lane quarter
    const int q = warp & 3;                // TMEM lane quarter this warp may access
    const int half = (warp - 2) >> 2;      // 0: warps 2-5, 1: warps 6-9
    if constexpr (EPI == EPI_GENERIC) {
      // generic form: interleaved over the 32-column chunks of the tile, staged through a [32][33] fp32 tile
      float* stg = reinterpret_cast<float*>(smem_gen + KSTAGES * Cfg::STAGE_BYTES) + (warp - 2) * STG_FLOATS;
      EpiArgs e2 = e;                      // the second row group's view of the epilogue
      e2.C = e.C2; e2.ldc = e.ldc2; e2.mask_bits = e.mask_bits2; e2.M = e.M2;
      int local = 0;
      for (int w = w_first; w < total_work; w += w_step, ++local) {
        const int tile = w / t.splits;
        const int n_blk = tile % t.num_n;
        int m_blk = m_block_of(tile);
        const bool g2 = t.m_split_blk > 0 && m_blk >= t.m_split_blk;
        if (g2) m_blk -= t.m_split_blk;
        const EpiArgs& eg = g2 ? e2 : e;
        const int as = local & 1;
        const uint32_t aphase = (local >> 1) & 1;
        mbar_wait(tfull_bar(as), aphase);
        __syncwarp();
        tcgen05_fence_after();
        const int r0 = m_blk * BM + q * 32;
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * BN);
#pragma unroll 1
        for (int c = half; c < BN / 32; c += 2) {
          float v[32];
          tmem_ld32(taddr + (uint32_t)(c * 32), v);
          tc_epi_chunk(eg, stg, r0, n_blk * BN + c * 32, lane, t.staged, v);
        }
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) { if (MC == 2) mbar_arrive_cluster_relaxed(mapa_shared(tempty_bar(as), 0)); else mbar_arrive_relaxed(tempty_bar(as)); }   // TMEM-only hand-off
      }
    } else {
      // fast forms (splits == 1): this warp owns columns [half*BN/2, (half+1)*BN/2) of its 32-row slab and walks
      // them in 64 B chunks; `pos` numbers its chunks over all tiles (ring slots, barrier parities, prefetch)
      using FE = FastEpi<EPI, BN>;
      constexpr int CW = FE::CW, NIN = EpiCfg<EPI>::NIN;
      constexpr int NEW = epi_warps<EPI>(), PARTS = NEW / 4, NCHT = BN / CW;      // column parts per lane quarter, chunks per slab
      const int ew = warp - 2;
      const int part = ew >> 2;
      // chunks [c_first, c_first + nch) of the 32-row slab belong to this warp (parts differ by at most one chunk)
      const int c_first = (part * NCHT) / PARTS;
      const int nch = ((part + 1) * NCHT) / PARTS - c_first;
      const uint32_t wbase = epi_base + ew * epi_warp_bytes<EPI>();
      FE fe{e, &tma_c, &tma_x, wbase, wbase + NIN * EPI_BOX_BYTES, epi_in_bar + 16u * ew, lane, q, half};
      auto coords = [&](uint32_t pos, int& row0, int& col) -> bool {
        const int lt = (int)(pos / (uint32_t)nch), ci = (int)(pos % (uint32_t)nch);
        const int w = w_first + lt * w_step;
        if (w >= total_work) return false;
        const int n_blk = w % t.num_n, m_blk = m_block_of(w);
        row0 = m_blk * BM + q * 32;
        col = n_blk * BN + (c_first + ci) * CW;
        return true;
      };
      if (NIN > 0 && lane == 0) {
        for (int i = 0; i < NIN; ++i) { int r, c; if (coords(i, r, c)) fe.issue_in(i, r, c); }
      }
      uint32_t pos = 0;
      int local = 0;
      for (int w = w_first; w < total_work; w += w_step, ++local) {
        const int as = local & 1;
        const uint32_t aphase = (local >> 1) & 1;
        mbar_wait(tfull_bar(as), aphase);
        __syncwarp();
        tcgen05_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * BN + c_first * CW);
        const int n_blk = w % t.num_n, m_blk = m_block_of(w);
        const int row0 = m_blk * BM + q * 32;
        const int col0 = n_blk * BN + c_first * CW;
#pragma unroll 1
        for (int ci = 0; ci < nch; ++ci, ++pos) {
          int nrow0 = 0, ncol = 0;
          const bool nv = NIN > 0 ? coords(pos + NIN, nrow0, ncol) : false;
          fe.chunk(pos, taddr + (uint32_t)(ci * CW), row0, col0 + ci * CW, nv, nrow0, ncol);
        }
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) { if (MC == 2) mbar_arrive_cluster_relaxed(mapa_shared(tempty_bar(as), 0)); else mbar_arrive_relaxed(tempty_bar(as)); }   // TMEM-only hand-off
      }
      if (lane == 0) bulk_wait_all();      // shared-memory boxes must outlive their TMA stores
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (MC) cluster_sync_all();          // the peer may still multicast into / arrive on this CTA's shared memory
  if (warp == 1) {
    __syncwarp();
    tcgen05_fence_after();
    if (MC == 2) tmem_dealloc_2cta(tmem_base, (uint32_t)Cfg::TMEM_COLS);
    else tmem_dealloc(tmem_base, (uint32_t)Cfg::TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

// 2-D bf16 tensor [outer, inner] with row pitch ld elements; box = 64 x box_outer, 128B swizzle.
int make_tmap_bf16(CUtensorMap* m, const void* ptr, int64_t inner, int64_t outer, int64_t ld, int box_outer) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return CLG_ERR_CUDA;
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) || ((ld * 2) & 15)) return CLG_ERR_UNSUPPORTED;
  cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
  cuuint64_t strides[1] = {(cuuint64_t)(ld * 2)};
  cuuint32_t box[2] = {64u, (cuuint32_t)box_outer};
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_cuda_error(cudaErrorInvalidValue, "cuTensorMapEncodeTiled"); return CLG_ERR_CUDA; }
  return CLG_OK;
}

// epilogue tile map: box = 64 B x 32 rows, 64B swizzle (matches sw64() in gemm_epilogue_fast.cuh)
int make_tmap_epi(CUtensorMap* m, const void* ptr, int dtype, int64_t inner, int64_t outer, int64_t ld) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return CLG_ERR_CUDA;
  const int esz = dtype == CLG_BF16 ? 2 : 4;
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) || ((ld * esz) & 15)) return CLG_ERR_UNSUPPORTED;
  cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
  cuuint64_t strides[1] = {(cuuint64_t)(ld * esz)};
  cuuint32_t box[2] = {(cuuint32_t)(64 / esz), 32u};
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = fn(m, dtype == CLG_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                  const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_cuda_error(cudaErrorInvalidValue, "cuTensorMapEncodeTiled(epilogue)"); return CLG_ERR_CUDA; }
  return CLG_OK;
}

// SMs the persistent GEMM grids occupy.  Under data parallelism NCCL's all-reduce kernels run concurrently with the
// backward GEMMs: with one persistent CTA per SM and a static tile schedule, a CTA whose SM is held by a NCCL CTA starts
// late and the whole GEMM waits for it.  clg_set_gemm_sm_limit(n) (parallel.attach) leaves sms - n SMs to NCCL.
static int g_gemm_sm_limit = 0;
static int gemm_sms() {
  static const int env = [] { const char* s = getenv("CLG_GEMM_SMS"); return s ? atoi(s) : 0; }();
  const int lim = env > 0 ? env : g_gemm_sm_limit;
  const int sms = num_sms();
  return (lim > 0 && lim < sms) ? (lim & ~1) : sms;
}

static bool env_flag(const char* name, bool dflt) {
  const char* s = getenv(name);
  if (!s || !s[0]) return dflt;
  return s[0] != '0';
}
static bool multicast_enabled() { static const bool v = env_flag("CLG_GEMM_MC", true); return v; }
static bool fast_epilogue_enabled() { static const bool v = env_flag("CLG_GEMM_FAST_EPI", true); return v; }

#ifndef CLG_PAIR_MODE
#define CLG_PAIR_MODE 2      // 1: multicast pair, 2: cta_group::2 pair (see TileCfg)
#endif
static constexpr int PAIR = CLG_PAIR_MODE;

template <int BN, bool A_K, bool B_K, int MC, int EPI>
static int launch_tc(const clgGemmParams& p, const EpiArgs& e0, cudaStream_t stream) {
  using Cfg = TileCfg<BN, EPI, MC>;
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] {
    attr_err = cudaFuncSetAttribute(gemm_tc_kernel<BN, A_K, B_K, MC, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    Cfg::SMEM_BYTES);
  });
  CLG_CUDA_OK(attr_err, "cudaFuncSetAttribute(gemm_tc)");

  CUtensorMap ta, tb, tc, tx;
  int st;
  if (A_K) st = make_tmap_bf16(&ta, p.A, p.K, p.M, p.lda, BM);
  else     st = make_tmap_bf16(&ta, p.A, p.M, p.K, p.lda, BK);
  if (st != CLG_OK) return st;
  if (B_K) st = make_tmap_bf16(&tb, p.B, p.K, p.N, p.ldb, MC ? BN / 2 : BN);
  else     st = make_tmap_bf16(&tb, p.B, p.N, p.K, p.ldb, BK);
  if (st != CLG_OK) return st;
  CUtensorMap ta2 = ta, tb2 = tb;
  const bool grouped = p.M2 > 0;
  if (grouped) {
    if (EPI != EPI_GENERIC || !MC || (p.M % (2 * BM)) != 0 || !p.A2 || !p.B2 || !p.C2) return CLG_ERR_ARG;
    if (A_K) st = make_tmap_bf16(&ta2, p.A2, p.K, p.M2, p.lda2, BM);
    else     st = make_tmap_bf16(&ta2, p.A2, p.M2, p.K, p.lda2, BK);
    if (st != CLG_OK) return st;
    if (B_K) st = make_tmap_bf16(&tb2, p.B2, p.K, p.N, p.ldb2, MC ? BN / 2 : BN);
    else     st = make_tmap_bf16(&tb2, p.B2, p.N, p.K, p.ldb2, BK);
    if (st != CLG_OK) return st;
  }
  if (EPI != EPI_GENERIC) {
    st = make_tmap_epi(&tc, p.C, p.c_dtype, p.N, p.M, p.ldc);
    if (st != CLG_OK) return st;
    if (EPI == EPI_RESID) st = make_tmap_epi(&tx, p.residual, CLG_F32, p.N, p.M, p.ld_res);
    else if (EPI == EPI_GELU_PAIR) { if (p.aux_out) st = make_tmap_epi(&tx, p.aux_out, CLG_BF16, p.N, p.M, p.ld_aux_out); else tx = tc; }
    else if (EPI == EPI_MUL_AUX) st = make_tmap_epi(&tx, p.aux_in, CLG_BF16, p.N, p.M, p.ld_aux_in);
    else tx = tc;
    if (st != CLG_OK) return st;
  } else {
    tc = ta; tx = ta;      // unused by the generic epilogue
  }

  TcArgs t;
  t.M = p.M; t.N = p.N; t.K = p.K;
  t.m_split_blk = grouped ? p.M / BM : 0;
  t.num_m = (p.M + BM - 1) / BM + (grouped ? (p.M2 + BM - 1) / BM : 0);
  t.num_n = (p.N + BN - 1) / BN;
  t.kblocks = (p.K + BK - 1) / BK;
  const int sms = gemm_sms();
  int splits = p.split_k;
  const bool can_split = (EPI == EPI_GENERIC && p.act == CLG_ACT_NONE && p.bias == nullptr && p.residual == nullptr &&
                          p.c_dtype == CLG_F32 && !p.dropout.p);
  if (splits <= 0) {
    splits = 1;
    const int tiles = (MC ? 2 * ((t.num_m + 1) / 2) : t.num_m) * t.num_n;
    if (can_split && tiles < sms && t.kblocks >= 16) {
      // pick the split count that best fills whole waves of SMs, >= 8 k-blocks (K = 512) per split
      const int max_s = t.kblocks / 8 < 32 ? t.kblocks / 8 : 32;
      double best = (double)tiles / sms;
      for (int s = 2; s <= max_s; ++s) {
        const int tot = tiles * s;
        const int waves = (tot + sms - 1) / sms;
        const double ef = (double)tot / ((double)waves * sms);
        if (ef > best + 0.02) { best = ef; splits = s; }
      }
    }
  }
  if (splits > 1 && !can_split) return CLG_ERR_ARG;
  t.kper = (t.kblocks + splits - 1) / splits;
  t.splits = (t.kblocks + t.kper - 1) / t.kper;
  EpiArgs e = e0;
  if (t.splits > 1) {
    if (!p.accumulate) {
      // split-K accumulates atomically: C must start from zero
      CLG_CUDA_OK(cudaMemset2DAsync(p.C, p.ldc * sizeof(float), 0, (size_t)p.N * sizeof(float), p.M, stream),
                  "memset(split-K C)");
      if (grouped)
        CLG_CUDA_OK(cudaMemset2DAsync(p.C2, p.ldc2 * sizeof(float), 0, (size_t)p.N * sizeof(float), p.M2, stream),
                    "memset(split-K C2)");
    }
    e.accumulate = 2;
  }
  t.staged = (p.N % 2 == 0) ? epi_stage_mode(e) : 0;
  if (grouped && t.staged) {
    EpiArgs e2 = e;
    e2.C = e.C2; e2.ldc = e.ldc2;
    const int m2 = epi_stage_mode(e2);
    if (m2 != t.staged) t.staged = (m2 == 0 || e.c_dtype != CLG_F32) ? 0 : 2;
  }
  if (MC) {
    t.total_units = ((t.num_m + 1) / 2) * t.num_n * t.splits;
    const int max_clusters = sms / 2;
    const int clusters = t.total_units < max_clusters ? t.total_units : max_clusters;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * clusters);
    cfg.blockDim = dim3(num_threads<EPI>());
    cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 2 : 1;
    CLG_CUDA_OK(cudaLaunchKernelEx(&cfg, gemm_tc_kernel<BN, A_K, B_K, MC, EPI>, ta, tb, tc, tx, ta2, tb2, t, e),
                "cudaLaunchKernelEx(gemm_tc, cluster 2)");
    return check_launch("gemm_tc_kernel(mc)");
  }
  t.total_units = t.num_m * t.num_n * t.splits;
  const int grid = t.total_units < sms ? t.total_units : sms;
  CLG_CUDA_OK(launch_pdl(gemm_tc_kernel<BN, A_K, B_K, MC, EPI>, dim3(grid), dim3(num_threads<EPI>()), Cfg::SMEM_BYTES, stream, ta, tb, tc, tx, ta2, tb2, t, e),
              "launch(gemm_tc)");
  return check_launch("gemm_tc_kernel");
}

// generic epilogue: every operand major-ness
template <int BN, int MC>
static int dispatch_generic(const clgGemmParams& p, const EpiArgs& e, cudaStream_t s) {
  if (p.a_kmajor && p.b_kmajor) return launch_tc<BN, true, true, MC, EPI_GENERIC>(p, e, s);
  if (p.a_kmajor && !p.b_kmajor) return launch_tc<BN, true, false, MC, EPI_GENERIC>(p, e, s);
  if (!p.a_kmajor && !p.b_kmajor) return launch_tc<BN, false, false, MC, EPI_GENERIC>(p, e, s);
  return launch_tc<BN, false, true, MC, EPI_GENERIC>(p, e, s);
}

// fast epilogues exist for the shapes of the step: A K-major (activations / their gradients as stored), B K-major
// (forward, W[N,K]) or MN-major (dgrad on W as stored), BN >= 128, multicast pairs
template <int BN>
static int dispatch_fast(int kind, const clgGemmParams& p, const EpiArgs& e, cudaStream_t s) {
  if (p.b_kmajor) {
    switch (kind) {
      case EPI_PLAIN: return launch_tc<BN, true, true, PAIR, EPI_PLAIN>(p, e, s);
      case EPI_RESID: return launch_tc<BN, true, true, PAIR, EPI_RESID>(p, e, s);
      case EPI_GELU_PAIR: return launch_tc<BN, true, true, PAIR, EPI_GELU_PAIR>(p, e, s);
      default: break;
    }
  } else {
    switch (kind) {
      case EPI_PLAIN: return launch_tc<BN, true, false, PAIR, EPI_PLAIN>(p, e, s);
      case EPI_RESID: return launch_tc<BN, true, false, PAIR, EPI_RESID>(p, e, s);
      case EPI_MUL_AUX: return launch_tc<BN, true, false, PAIR, EPI_MUL_AUX>(p, e, s);
      default: break;
    }
  }
  return dispatch_generic<BN, PAIR>(p, e, s);
}

template <int BN>
static int dispatch_major(const clgGemmParams& p, const EpiArgs& e, cudaStream_t s) {
  // pairs of vertically adjacent tiles share the B tile through TMA multicast whenever there are at least two
  // row blocks (BN >= 128 so that each CTA's half of an MN-major B tile is a whole 64-wide box)
  const bool mc = BN >= 128 && p.M > BM && multicast_enabled();
  if (p.M2 > 0 && !mc) return CLG_ERR_UNSUPPORTED;
  if (mc) {
    constexpr int BNF = BN >= 128 ? BN : 128;
    if (p.M2 <= 0 && p.a_kmajor && p.split_k <= 1 && fast_epilogue_enabled()) {
      const int kind = classify_epilogue(e);
      if (kind != EPI_GENERIC) return dispatch_fast<BNF>(kind, p, e, s);
    }
    return dispatch_generic<BNF, PAIR>(p, e, s);
  }
  return dispatch_generic<BN, 0>(p, e, s);
}

// entry used by clg_gemm
int gemm_tc(const clgGemmParams& p, cudaStream_t stream) {
  if (p.a_dtype != CLG_BF16 || p.b_dtype != CLG_BF16) return CLG_ERR_UNSUPPORTED;
  if (p.M <= 0 || p.N <= 0 || p.K <= 0) return CLG_ERR_ARG;
  const EpiArgs e = make_epi(p);
  // Tile choice: BN = 256 when the tile count still fills the machine in whole waves reasonably,
  // BN = 128 for the N = 768 family (fewer idle SMs in the last wave).
  const int sms = gemm_sms();
  if (p.M2 > 0 && ((p.M % (2 * BM)) != 0 || p.N < 128)) return CLG_ERR_UNSUPPORTED;
  const int num_m = (p.M + BM - 1) / BM + (p.M2 > 0 ? (p.M2 + BM - 1) / BM : 0);
  auto eff = [&](int bn) {
    const int tiles = num_m * ((p.N + bn - 1) / bn);
    const int waves = (tiles + sms - 1) / sms;
    return (double)tiles / ((double)waves * sms);
  };
  // BN = 256 whenever a CTA pair still gets at least one 256 x 256 unit: the cta_group::2 MMA at N = 256 has the
  // lowest shared-memory traffic per FLOP (measured: the tensor pipe is shared-memory-bandwidth bound at N = 128);
  // otherwise the tile width with fewer idle SMs in the last wave.
  int bn = 256;
  if (p.N <= 64) bn = 64;
  else if (p.N <= 128) bn = 128;
  else {
    const int units256 = ((num_m + 1) / 2) * ((p.N + 255) / 256);
    if (!(p.N % 256 == 0 && units256 >= sms / 2) && eff(128) > eff(256) + 0.08) bn = 128;
    // weight-gradient shapes (few output tiles, K = tokens): 256-wide units filled up by split-K when a split count
    // exists that keeps >= 90% of the CTA pairs busy in whole rounds (FFN wgrad: 36 units x 2 splits on 74 pairs)
    const bool can_split = p.act == CLG_ACT_NONE && !p.bias && !p.residual && p.c_dtype == CLG_F32 && !p.dropout.p && !p.colsum &&
                           p.split_k <= 0;
    if (bn == 128 && can_split && p.N % 256 == 0 && p.M > BM && (p.K + BK - 1) / BK >= 32) {
      const int clusters = sms / 2, kblocks = (p.K + BK - 1) / BK;
      for (int sp = 2; sp <= 3 && kblocks / sp >= 8; ++sp) {      // more splits: measured slower than 128-wide units (QKV wgrad)
        const int tot = units256 * sp;
        const int rounds = (tot + clusters - 1) / clusters;
        if ((double)tot / ((double)rounds * clusters) >= 0.9) { bn = 256; break; }
      }
    }
  }
  static const int force_bn = [] { const char* s = getenv("CLG_GEMM_FORCE_BN"); return s ? atoi(s) : 0; }();
  if (force_bn == 256 && p.N > 128) bn = 256;
  if (force_bn == 128 && p.N > 64) bn = 128;
  if (force_bn == 64 && p.M2 <= 0) bn = 64;
  // (A second launch with narrower units for the rows of the last, partial round was measured and dropped: the step is
  // bound by L2 -> SM operand traffic when all SMs are busy, so the few clusters of a partial round run their units
  // ~1.7x faster than a full round and the split gains nothing.)
  if (bn == 256) return dispatch_major<256>(p, e, stream);
  if (bn == 128) return dispatch_major<128>(p, e, stream);
  return dispatch_major<64>(p, e, stream);
}

}  // namespace clg

extern "C" int clg_set_gemm_sm_limit(int n) { clg::g_gemm_sm_limit = n; return CLG_OK; }
