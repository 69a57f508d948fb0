// TMA-fed epilogues of the tcgen05 GEMM for the four hot forms of the step.  Each epilogue warp owns a
// 32-row slab of the accumulator tile (its TMEM lane quarter, lane = row) and walks it in chunks of
// 64 B per row (16 fp32 or 32 bf16 columns).  Elementwise inputs (fp32 residual / bf16 saved gelu') are
// TMA-prefetched into per-warp shared-memory boxes one or two chunks ahead; results are written to
// per-warp boxes (64B-swizzled, conflict-free 16 B accesses) and leave with one TMA store per chunk -
// no per-lane global addressing, every HBM transaction a full 64 B row segment, latency hidden.
//
//   EPI_PLAIN     C(bf16)  = acc + bias                                  QKV fwd, out-proj dgrad
//   EPI_RESID     C(fp32)  = dropout(acc + bias) + residual(fp32)        out-proj / FFN2 fwd, QKV / FFN1 dgrad
//   EPI_GELU_PAIR C(bf16)  = gelu(acc + bias), aux_out(bf16) = gelu'     FFN1 fwd
//   EPI_MUL_AUX   C(bf16)  = acc * aux_in(bf16)                          FFN2 dgrad
#pragma once
#include "gemm_epilogue.cuh"
#include "tc_common.cuh"

namespace clg {

enum { EPI_GENERIC = 0, EPI_PLAIN = 1, EPI_RESID = 2, EPI_GELU_PAIR = 3, EPI_MUL_AUX = 4 };

static constexpr int EPI_BOX_BYTES = 2048;          // 32 rows x 64 B

template <int EPI> struct EpiCfg;
template <> struct EpiCfg<EPI_GENERIC>   { static constexpr int CW = 32, NIN = 0, NOUT = 0; };
template <> struct EpiCfg<EPI_PLAIN>     { static constexpr int CW = 32, NIN = 0, NOUT = 2; };
template <> struct EpiCfg<EPI_RESID>     { static constexpr int CW = 16, NIN = 2, NOUT = 2; };
template <> struct EpiCfg<EPI_GELU_PAIR> { static constexpr int CW = 32, NIN = 0, NOUT = 2; };   // one (h, g') pair
template <> struct EpiCfg<EPI_MUL_AUX>   { static constexpr int CW = 32, NIN = 1, NOUT = 1; };

template <int EPI> __host__ __device__ constexpr int epi_warp_bytes() { return (EpiCfg<EPI>::NIN + EpiCfg<EPI>::NOUT) * EPI_BOX_BYTES; }

// byte offset of 16 B unit j (0..3) of row `lane` inside a 64B-swizzled 32 x 64 B box
__device__ __forceinline__ uint32_t sw64(int lane, int j) {
  return (uint32_t)(lane * 64 + ((j ^ ((lane >> 1) & 3)) << 4));
}

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

// write 32 fp32 values of this lane's row as bf16 into a swizzled box
__device__ __forceinline__ void box_store_bf16(uint32_t box, int lane, const float* v) {
#pragma unroll
  for (int j = 0; j < 4; ++j)
    sts128(box + sw64(lane, j), pack_bf16x2(v[8 * j + 0], v[8 * j + 1]), pack_bf16x2(v[8 * j + 2], v[8 * j + 3]),
           pack_bf16x2(v[8 * j + 4], v[8 * j + 5]), pack_bf16x2(v[8 * j + 6], v[8 * j + 7]));
}
__device__ __forceinline__ void box_store_f32(uint32_t box, int lane, const float* v) {
#pragma unroll
  for (int j = 0; j < 4; ++j)
    sts128(box + sw64(lane, j), __float_as_uint(v[4 * j + 0]), __float_as_uint(v[4 * j + 1]), __float_as_uint(v[4 * j + 2]),
           __float_as_uint(v[4 * j + 3]));
}
__device__ __forceinline__ void box_load_bf16(uint32_t box, int lane, float* v) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const uint4 u = lds128(box + sw64(lane, j));
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      v[8 * j + 2 * t + 0] = __uint_as_float(w[t] << 16);
      v[8 * j + 2 * t + 1] = __uint_as_float(w[t] & 0xffff0000u);
    }
  }
}
__device__ __forceinline__ void box_load_f32(uint32_t box, int lane, float* v) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const uint4 u = lds128(box + sw64(lane, j));
    v[4 * j + 0] = __uint_as_float(u.x); v[4 * j + 1] = __uint_as_float(u.y);
    v[4 * j + 2] = __uint_as_float(u.z); v[4 * j + 3] = __uint_as_float(u.w);
  }
}

// The shared-memory reads of a box must have RETURNED before the TMA engine may overwrite it: __syncwarp orders
// instruction issue, not the completion of queued ld.shared (observed: under tensor-core operand traffic a
// queued read saw the next chunk's data).  Tying the registers to an empty asm forces the scoreboard wait; the
// proxy fence orders the generic-proxy reads before the async-proxy write.
template <int N> __device__ __forceinline__ void consume_regs(float* v) {
#pragma unroll
  for (int j = 0; j < N; j += 8) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v) + j;
    asm volatile("" : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]) :: "memory");
  }
}

// host: which fast form (if any) computes this epilogue exactly
inline int classify_epilogue(const EpiArgs& e) {
  auto al16 = [](const void* p, int64_t ld, int esz) {
    return p && (reinterpret_cast<uintptr_t>(p) & 15) == 0 && ((ld * esz) & 15) == 0;
  };
  if (e.alpha != 1.0f || e.mask_bits || e.accumulate || (e.N % 32) != 0) return EPI_GENERIC;
  if (e.colsum && e.act != CLG_ACT_MUL_AUX) return EPI_GENERIC;
  if (e.bias && (reinterpret_cast<uintptr_t>(e.bias) & 15)) return EPI_GENERIC;
  const bool c_bf16 = e.c_dtype == CLG_BF16 && al16(e.C, e.ldc, 2);
  const bool c_f32 = e.c_dtype == CLG_F32 && al16(e.C, e.ldc, 4);
  if (e.act == CLG_ACT_NONE && !e.aux_in && !e.aux_out) {
    if (c_bf16 && !e.residual && !e.drop.on) return EPI_PLAIN;
    if (c_f32 && e.residual && al16(e.residual, e.ld_res, 4) && !(e.drop.on && e.drop_before_act)) return EPI_RESID;
    return EPI_GENERIC;
  }
  // aux_out == NULL: forward-only runs (eval) do not need gelu' - the same epilogue with one output
  if (e.act == CLG_ACT_GELU_GRAD && c_bf16 && e.bias && !e.residual && !e.drop.on && !e.aux_in &&
      (!e.aux_out || (e.aux_out_dtype == CLG_BF16 && al16(e.aux_out, e.ld_aux_out, 2))))
    return EPI_GELU_PAIR;
  if (e.act == CLG_ACT_MUL_AUX && c_bf16 && !e.bias && !e.residual && !e.drop.on && !e.aux_out &&
      e.aux_in_dtype == CLG_BF16 && al16(e.aux_in, e.ld_aux_in, 2))
    return EPI_MUL_AUX;
  return EPI_GENERIC;
}

// Per-warp streaming state of the fast epilogue: `pos` counts the chunks this warp has processed since
// kernel start (over all of its tiles), which fixes the ring slots and mbarrier parities.
template <int EPI, int BN>
struct FastEpi {
  using Cfg = EpiCfg<EPI>;
  static constexpr int CW = Cfg::CW;
  static_assert(BN / CW >= 3, "tile too narrow for the fast epilogue");

  const EpiArgs& e;
  const CUtensorMap* tm_c;
  const CUtensorMap* tm_x;                            // aux_out (GELU_PAIR) | residual (RESID) | aux_in (MUL_AUX)
  uint32_t in_box, out_box, in_bar;                   // shared-memory addresses of this warp's boxes / barriers
  int lane, q, half;

  __device__ __forceinline__ uint32_t ibox(int s) const { return in_box + s * EPI_BOX_BYTES; }
  __device__ __forceinline__ uint32_t obox(int s) const { return out_box + s * EPI_BOX_BYTES; }

  // lane 0: prefetch the elementwise input of chunk `ci` of the tile at (row0, col0)
  __device__ __forceinline__ void issue_in(uint32_t pos, int row0, int col) const {
    if (Cfg::NIN == 0) return;
    const int s = pos % (Cfg::NIN ? Cfg::NIN : 1);
    const uint32_t bar = in_bar + 8u * s;
    mbar_expect_tx(bar, EPI_BOX_BYTES);
    tma_load_2d(ibox(s), tm_x, bar, col, row0);
  }
  __device__ __forceinline__ void wait_in(uint32_t pos) const {
    const int n = Cfg::NIN ? Cfg::NIN : 1;
    mbar_wait(in_bar + 8u * (pos % n), (pos / n) & 1u);
  }

  // One chunk: accumulators of 32 rows x CW columns at TMEM address taddr; global coordinates (row0, col).
  // `next_valid`/(nrow0, ncol): coordinates of the chunk NIN positions ahead (prefetched here).
  __device__ __forceinline__ void chunk(uint32_t pos, uint32_t taddr, int row0, int col, bool next_valid, int nrow0, int ncol) {
    const int row = row0 + lane;
    if (EPI == EPI_RESID) {
      float acc[16], rin[16];
      tmem_ld16_nowait(taddr, acc);
      wait_in(pos);
      box_load_f32(ibox(pos % 2), lane, rin);
      consume_regs<16>(rin);
      fence_proxy_async();
      __syncwarp();
      if (lane == 0 && next_valid) issue_in(pos + 2, nrow0, ncol);
      tmem_wait_ld_dep16(acc);
      if (e.bias) {
        const float4* b4 = reinterpret_cast<const float4*>(e.bias + col);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 b = __ldg(b4 + j);
          acc[4 * j + 0] += b.x; acc[4 * j + 1] += b.y; acc[4 * j + 2] += b.z; acc[4 * j + 3] += b.w;
        }
      }
      if (e.drop.on) {
        const uint32_t eidx = (uint32_t)((uint64_t)row * (uint64_t)e.N + (uint64_t)col);
#pragma unroll
        for (int j = 0; j < 16; j += 2) {
          const uint32_t h = lowbias32(((eidx + j) >> 1) ^ e.drop.key);
          acc[j] = ((h & 0xffffu) >= e.drop.thresh) ? acc[j] * e.drop.scale : 0.f;
          acc[j + 1] = ((h >> 16) >= e.drop.thresh) ? acc[j + 1] * e.drop.scale : 0.f;
        }
      }
#pragma unroll
      for (int j = 0; j < 16; ++j) acc[j] += rin[j];
      const int s = pos % 2;
      if (lane == 0) bulk_wait_read<1>();                // the store that last used this slot has read it
      __syncwarp();
      box_store_f32(obox(s), lane, acc);
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) { tma_store_2d(tm_c, obox(s), col, row0); bulk_commit(); }
    } else if (EPI == EPI_PLAIN) {
      float acc[32];
      tmem_ld32(taddr, acc);
      if (e.bias) {
        const float4* b4 = reinterpret_cast<const float4*>(e.bias + col);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 b = __ldg(b4 + j);
          acc[4 * j + 0] += b.x; acc[4 * j + 1] += b.y; acc[4 * j + 2] += b.z; acc[4 * j + 3] += b.w;
        }
      }
      const int s = pos % 2;
      if (lane == 0) bulk_wait_read<1>();
      __syncwarp();
      box_store_bf16(obox(s), lane, acc);
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) { tma_store_2d(tm_c, obox(s), col, row0); bulk_commit(); }
    } else if (EPI == EPI_GELU_PAIR) {
      float acc[32], dg[32];
      tmem_ld32(taddr, acc);
      {
        const float4* b4 = reinterpret_cast<const float4*>(e.bias + col);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 b = __ldg(b4 + j);
          acc[4 * j + 0] += b.x; acc[4 * j + 1] += b.y; acc[4 * j + 2] += b.z; acc[4 * j + 3] += b.w;
        }
      }
#pragma unroll
      for (int j = 0; j < 32; ++j) gelu_pair_fast(acc[j], acc[j], dg[j]);
      if (e.aux_out) {
        if (lane == 0) bulk_wait_read<0>();              // previous chunk's pair of stores has read both boxes
        __syncwarp();
        box_store_bf16(obox(0), lane, acc);
        box_store_bf16(obox(1), lane, dg);
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          tma_store_2d(tm_c, obox(0), col, row0);
          tma_store_2d(tm_x, obox(1), col, row0);
          bulk_commit();
        }
      } else {                                           // one output: the two boxes alternate
        const int s = pos % 2;
        if (lane == 0) bulk_wait_read<1>();
        __syncwarp();
        box_store_bf16(obox(s), lane, acc);
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) { tma_store_2d(tm_c, obox(s), col, row0); bulk_commit(); }
      }
    } else if (EPI == EPI_MUL_AUX) {
      float acc[32], ax[32];
      tmem_ld32_nowait(taddr, acc);
      wait_in(pos);
      box_load_bf16(ibox(0), lane, ax);
      consume_regs<32>(ax);
      fence_proxy_async();
      __syncwarp();
      if (lane == 0 && next_valid) issue_in(pos + 1, nrow0, ncol);
      tmem_wait_ld_dep32(acc);
#pragma unroll
      for (int j = 0; j < 32; ++j) acc[j] *= ax[j];
      if (lane == 0) bulk_wait_read<0>();
      __syncwarp();
      box_store_bf16(obox(0), lane, acc);
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) { tma_store_2d(tm_c, obox(0), col, row0); bulk_commit(); }
      if (e.colsum) {           // fused bias gradient of the layer below (rows beyond M are exact zeros: OOB boxes)
        const float cs = warp_colsum32(acc, lane);
        atomicAdd(e.colsum + col + lane, cs);
      }
    }
  }
};

}  // namespace clg
