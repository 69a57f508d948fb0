// LayerNorm forward / backward: one warp per row, the row lives in registers as float4 chunks
// (lane owns columns 4*(lane + 32*i) .. +3), two-pass variance, all global accesses 128-bit.
// HBM-bound by design: forward reads x once and writes y (+ the bf16 GEMM-operand copy); backward
// reads dy and x once and writes dx (+ operand copy).  The column sums of the backward (dgamma,
// dbeta, dbias) are accumulated in per-warp PRIVATE shared-memory rows (no atomics, no 72-register
// accumulator set -> 2 CTAs / SM), reduced across the warps of the CTA at the end and added to
// global memory with one atomicAdd per column per CTA.
// A scalar generic kernel covers row lengths that are not a multiple of 128.
#include "common.cuh"

namespace clg {

static constexpr int LN_WARPS = 8;
static constexpr int LN_THREADS = LN_WARPS * 32;
static constexpr int LN_MAX_N = 1536;

__device__ __forceinline__ void store_lowp4(void* p, int dtype, int64_t i, const float4 v) {
  if (dtype == CLG_BF16) {
    __nv_bfloat162 h0 = __floats2bfloat162_rn(v.x, v.y), h1 = __floats2bfloat162_rn(v.z, v.w);
    uint2 u;
    u.x = *reinterpret_cast<uint32_t*>(&h0); u.y = *reinterpret_cast<uint32_t*>(&h1);
    *reinterpret_cast<uint2*>((bf16*)p + i) = u;
  } else {
    *reinterpret_cast<float4*>((float*)p + i) = v;
  }
}

__device__ __forceinline__ float4 drop4(const Drop& d, uint64_t idx, float4 v) {
  if (!d.on) return v;
  // idx is a multiple of 4: two hashes cover the four elements
  const uint32_t i = (uint32_t)idx;
  const uint32_t h0 = lowbias32((i >> 1) ^ d.key), h1 = lowbias32(((i >> 1) + 1) ^ d.key);
  v.x = ((h0 & 0xffffu) >= d.thresh) ? v.x * d.scale : 0.f;
  v.y = ((h0 >> 16) >= d.thresh) ? v.y * d.scale : 0.f;
  v.z = ((h1 & 0xffffu) >= d.thresh) ? v.z * d.scale : 0.f;
  v.w = ((h1 >> 16) >= d.thresh) ? v.w * d.scale : 0.f;
  return v;
}

// ------------------------------------------------------------------------------------------------
// forward, N = 128 * C4
// ------------------------------------------------------------------------------------------------
// gamma / beta are staged in shared memory (not in 2 x C4 float4 registers per lane): a warp normalises one or two rows,
// so registers spent on them only cost occupancy - at 64 registers (N = 768, MINB = 4) four CTAs are resident per SM and
// twice as many row loads are in flight as with the 120-register version (measured: 28.7 -> 26.7 us per launch on 15.4 k
// rows from DRAM, 37.9 -> 34.1 us incl. launch latency with the input in L2; 3 CTAs / 80 registers: 26.7 / 35.6).  The
// first row's loads are issued before the staging barrier.
template <int C4, int MINB>
__global__ void __launch_bounds__(LN_THREADS, MINB)
ln_fwd_vec_kernel(const clgLayerNormFwdParams p, const Drop drop) {
  constexpr int N = 128 * C4;
  __shared__ float4 s_gam[N / 4], s_bet[N / 4];
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float invN = 1.0f / (float)N;
  float4 v[C4];
  int64_t r = (int64_t)blockIdx.x * LN_WARPS + warp;
  bool loaded = false;
  if (r < p.rows) {
    const float* xr = p.x + r * p.ldx;
#pragma unroll
    for (int i = 0; i < C4; ++i) v[i] = *reinterpret_cast<const float4*>(xr + 4 * (lane + 32 * i));
    loaded = true;
  }
  for (int i = threadIdx.x; i < N / 4; i += LN_THREADS) {
    s_gam[i] = *reinterpret_cast<const float4*>(p.gamma + 4 * i);
    s_bet[i] = *reinterpret_cast<const float4*>(p.beta + 4 * i);
  }
  __syncthreads();
  for (; r < p.rows; r += (int64_t)gridDim.x * LN_WARPS) {
    if (!loaded) {
      const float* xr = p.x + r * p.ldx;
#pragma unroll
      for (int i = 0; i < C4; ++i) v[i] = *reinterpret_cast<const float4*>(xr + 4 * (lane + 32 * i));
    }
    loaded = false;
    const float pre = (p.row_valid && p.valid_pre) ? (p.row_valid[r] ? 1.f : 0.f) : 1.f;
    const float post = (p.row_valid && p.valid_post) ? (p.row_valid[r] ? 1.f : 0.f) : 1.f;
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < C4; ++i) {
      v[i].x *= pre; v[i].y *= pre; v[i].z *= pre; v[i].w *= pre;
      s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
    const float mean = warp_sum(s) * invN;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < C4; ++i) {
      const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
      q += (a * a + b * b) + (c * c + d * d);
    }
    const float rstd = 1.0f / sqrtf(warp_sum(q) * invN + p.eps);
    if (lane == 0) {
      if (p.mean) p.mean[r] = mean;
      if (p.rstd) p.rstd[r] = rstd;
    }
    const int64_t orow = map_row(p.out_map, r);
    const float* addr = p.add_mat ? p.add_mat + map_row(p.add_map, r) * p.ld_add : nullptr;
#pragma unroll
    for (int i = 0; i < C4; ++i) {
      const int c = 4 * (lane + 32 * i);
      const float4 gam = s_gam[lane + 32 * i], bet = s_bet[lane + 32 * i];
      float4 y;
      y.x = ((v[i].x - mean) * rstd * gam.x + bet.x) * post;
      y.y = ((v[i].y - mean) * rstd * gam.y + bet.y) * post;
      y.z = ((v[i].z - mean) * rstd * gam.z + bet.z) * post;
      y.w = ((v[i].w - mean) * rstd * gam.w + bet.w) * post;
      if (addr) { const float4 a = *reinterpret_cast<const float4*>(addr + c); y.x += a.x; y.y += a.y; y.z += a.z; y.w += a.w; }
      if (p.add_vec) { const float4 a = *reinterpret_cast<const float4*>(p.add_vec + c); y.x += a.x; y.y += a.y; y.z += a.z; y.w += a.w; }
      y = drop4(drop, (uint64_t)r * (uint64_t)N + c, y);
      if (p.y) *reinterpret_cast<float4*>(p.y + orow * p.ldy + c) = y;
      if (p.y_lowp) store_lowp4(p.y_lowp, p.lowp_dtype, orow * p.ldy_lowp + c, y);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// backward, N = 128 * C4
// ------------------------------------------------------------------------------------------------
template <int C4>
__global__ void __launch_bounds__(LN_THREADS, 2)
ln_bwd_vec_kernel(const clgLayerNormBwdParams p, const Drop drop_out, const Drop drop_in) {
  extern __shared__ float4 sm4[];
  pdl_trigger();
  pdl_wait();
  constexpr int N = 128 * C4;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // smem: gamma [N] | per-warp accumulators [LN_WARPS][3][N] (only the arrays that are needed)
  float4* sgam = sm4;
  const int n_acc = (p.dgamma ? 1 : 0) + (p.dbeta ? 1 : 0) + (p.dbias ? 1 : 0);
  float4* acc = sm4 + N / 4 + (size_t)warp * n_acc * (N / 4);
  float4* a_g = acc;
  float4* a_b = a_g + (p.dgamma ? N / 4 : 0);
  float4* a_bias = a_b + (p.dbeta ? N / 4 : 0);
  for (int i = threadIdx.x; i < N / 4; i += LN_THREADS) sgam[i] = *reinterpret_cast<const float4*>(p.gamma + 4 * i);
  for (int i = lane; i < n_acc * (N / 4); i += 32) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  __syncthreads();
  const float invN = 1.0f / (float)N;
  for (int64_t r = (int64_t)blockIdx.x * LN_WARPS + warp; r < p.rows; r += (int64_t)gridDim.x * LN_WARPS) {
    const float pre = (p.row_valid && p.valid_pre) ? (p.row_valid[r] ? 1.f : 0.f) : 1.f;
    const float post = (p.row_valid && p.valid_post) ? (p.row_valid[r] ? 1.f : 0.f) : 1.f;
    const float* dyr = p.dy + map_row(p.dy_map, r) * p.lddy;
    const float* xr = p.x + r * p.ldx;
    float4 d[C4], xh[C4];
#pragma unroll
    for (int i = 0; i < C4; ++i) {
      d[i] = *reinterpret_cast<const float4*>(dyr + 4 * (lane + 32 * i));
      xh[i] = *reinterpret_cast<const float4*>(xr + 4 * (lane + 32 * i));
    }
    {
      // the warp's next row: one 128 B line per lane towards L2 while this row is being reduced (the saved forward
      // activation x comes from DRAM; both register sets stay live to the end of the row, so there is no room to
      // load the next row early).  Step-level A/B on one box: 12.03 / 12.06 ms with, 12.11 / 12.17 ms without; the
      // kernel timed alone (both inputs in L2, or both in DRAM) does not change (profiles/r02_ln_micro_smem_gamma.txt)
      const int64_t rn = r + (int64_t)gridDim.x * LN_WARPS;
      if (rn < p.rows) {
        const float* xn = p.x + rn * p.ldx;
        const float* dn = p.dy + map_row(p.dy_map, rn) * p.lddy;
        for (int l = lane; l < C4 * 4; l += 32) {
          asm volatile("prefetch.global.L2 [%0];" ::"l"(xn + 32 * l));
          asm volatile("prefetch.global.L2 [%0];" ::"l"(dn + 32 * l));
        }
      }
    }
    const float mean = p.mean[r], rstd = p.rstd[r];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < C4; ++i) {
      const int c4 = lane + 32 * i;
      d[i] = drop4(drop_out, (uint64_t)r * (uint64_t)N + 4 * c4, d[i]);
      d[i].x *= post; d[i].y *= post; d[i].z *= post; d[i].w *= post;
      xh[i].x = (xh[i].x * pre - mean) * rstd; xh[i].y = (xh[i].y * pre - mean) * rstd;
      xh[i].z = (xh[i].z * pre - mean) * rstd; xh[i].w = (xh[i].w * pre - mean) * rstd;
      if (p.dgamma) {
        float4 a = a_g[c4];
        a.x += d[i].x * xh[i].x; a.y += d[i].y * xh[i].y; a.z += d[i].z * xh[i].z; a.w += d[i].w * xh[i].w;
        a_g[c4] = a;
      }
      if (p.dbeta) {
        float4 a = a_b[c4];
        a.x += d[i].x; a.y += d[i].y; a.z += d[i].z; a.w += d[i].w;
        a_b[c4] = a;
      }
      const float4 g4 = sgam[c4];
      d[i].x *= g4.x; d[i].y *= g4.y; d[i].z *= g4.z; d[i].w *= g4.w;     // d now holds g = dy * gamma
      s1 += (d[i].x + d[i].y) + (d[i].z + d[i].w);
      s2 += (d[i].x * xh[i].x + d[i].y * xh[i].y) + (d[i].z * xh[i].z + d[i].w * xh[i].w);
    }
    s1 = warp_sum(s1) * invN;
    s2 = warp_sum(s2) * invN;
    const float k = rstd * pre;
#pragma unroll
    for (int i = 0; i < C4; ++i) {
      const int c4 = lane + 32 * i;
      float4 dx;
      dx.x = k * (d[i].x - s1 - xh[i].x * s2); dx.y = k * (d[i].y - s1 - xh[i].y * s2);
      dx.z = k * (d[i].z - s1 - xh[i].z * s2); dx.w = k * (d[i].w - s1 - xh[i].w * s2);
      if (p.dx) *reinterpret_cast<float4*>(p.dx + r * p.lddx + 4 * c4) = dx;
      const float4 din = drop4(drop_in, (uint64_t)r * (uint64_t)N + 4 * c4, dx);
      if (p.dx_lowp) store_lowp4(p.dx_lowp, p.lowp_dtype, r * p.lddx_lowp + 4 * c4, din);
      if (p.dbias) {
        float4 a = a_bias[c4];
        a.x += din.x; a.y += din.y; a.z += din.z; a.w += din.w;
        a_bias[c4] = a;
      }
    }
  }
  __syncthreads();
  // reduce the per-warp rows and add to global: thread t owns columns t, t + 256, ...
  const float* accf = reinterpret_cast<const float*>(sm4 + N / 4);
  float* outs[3] = {p.dgamma, p.dbeta, p.dbias};
  int slot = 0;
#pragma unroll
  for (int which = 0; which < 3; ++which) {
    if (!outs[which]) continue;
    for (int c = threadIdx.x; c < N; c += LN_THREADS) {
      float s = 0.f;
#pragma unroll
      for (int w = 0; w < LN_WARPS; ++w) s += accf[((size_t)w * n_acc + slot) * N + c];
      atomicAdd(outs[which] + c, s);
    }
    ++slot;
  }
}

// ------------------------------------------------------------------------------------------------
// generic scalar kernels (any N <= 1536)
// ------------------------------------------------------------------------------------------------
static constexpr int PERG = LN_MAX_N / 32;

__global__ void __launch_bounds__(LN_THREADS)
ln_fwd_gen_kernel(const clgLayerNormFwdParams p, const Drop drop) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int N = p.N;
  const float invN = 1.0f / (float)N;
  for (int64_t r = (int64_t)blockIdx.x * LN_WARPS + warp; r < p.rows; r += (int64_t)gridDim.x * LN_WARPS) {
    const float* xr = p.x + r * p.ldx;
    const float pre = (p.row_valid && p.valid_pre) ? (p.row_valid[r] ? 1.f : 0.f) : 1.f;
    const float post = (p.row_valid && p.valid_post) ? (p.row_valid[r] ? 1.f : 0.f) : 1.f;
    float s = 0.f;
    for (int c = lane; c < N; c += 32) s += xr[c] * pre;
    const float mean = warp_sum(s) * invN;
    float q = 0.f;
    for (int c = lane; c < N; c += 32) { const float d = xr[c] * pre - mean; q += d * d; }
    const float rstd = 1.0f / sqrtf(warp_sum(q) * invN + p.eps);
    if (lane == 0) {
      if (p.mean) p.mean[r] = mean;
      if (p.rstd) p.rstd[r] = rstd;
    }
    const int64_t orow = map_row(p.out_map, r);
    const float* addr = p.add_mat ? p.add_mat + map_row(p.add_map, r) * p.ld_add : nullptr;
    for (int c = lane; c < N; c += 32) {
      float y = ((xr[c] * pre - mean) * rstd * p.gamma[c] + p.beta[c]) * post;
      if (addr) y += addr[c];
      if (p.add_vec) y += p.add_vec[c];
      y = drop_apply(drop, (uint64_t)r * (uint64_t)N + c, y);
      if (p.y) p.y[orow * p.ldy + c] = y;
      if (p.y_lowp) store_from_f32(p.y_lowp, p.lowp_dtype, orow * p.ldy_lowp + c, y);
    }
  }
}

__global__ void __launch_bounds__(LN_THREADS)
ln_bwd_gen_kernel(const clgLayerNormBwdParams p, const Drop drop_out, const Drop drop_in) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int N = p.N;
  const float invN = 1.0f / (float)N;
  for (int64_t r = (int64_t)blockIdx.x * LN_WARPS + warp; r < p.rows; r += (int64_t)gridDim.x * LN_WARPS) {
    const float pre = (p.row_valid && p.valid_pre) ? (p.row_valid[r] ? 1.f : 0.f) : 1.f;
    const float post = (p.row_valid && p.valid_post) ? (p.row_valid[r] ? 1.f : 0.f) : 1.f;
    const float* dyr = p.dy + map_row(p.dy_map, r) * p.lddy;
    const float* xr = p.x + r * p.ldx;
    const float mean = p.mean[r], rstd = p.rstd[r];
    float s1 = 0.f, s2 = 0.f;
    for (int c = lane; c < N; c += 32) {
      const float d = drop_apply(drop_out, (uint64_t)r * (uint64_t)N + c, dyr[c]) * post;
      const float xh = (xr[c] * pre - mean) * rstd;
      const float g = d * p.gamma[c];
      s1 += g; s2 += g * xh;
      if (p.dgamma) atomicAdd(p.dgamma + c, d * xh);
      if (p.dbeta) atomicAdd(p.dbeta + c, d);
    }
    s1 = warp_sum(s1) * invN;
    s2 = warp_sum(s2) * invN;
    for (int c = lane; c < N; c += 32) {
      const float d = drop_apply(drop_out, (uint64_t)r * (uint64_t)N + c, dyr[c]) * post;
      const float xh = (xr[c] * pre - mean) * rstd;
      const float dx = rstd * (d * p.gamma[c] - s1 - xh * s2) * pre;
      if (p.dx) p.dx[r * p.lddx + c] = dx;
      const float din = drop_apply(drop_in, (uint64_t)r * (uint64_t)N + c, dx);
      if (p.dx_lowp) store_from_f32(p.dx_lowp, p.lowp_dtype, r * p.lddx_lowp + c, din);
      if (p.dbias) atomicAdd(p.dbias + c, din);
    }
  }
}

static bool aligned16(const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; }

template <int C4>
static int launch_bwd_vec(const clgLayerNormBwdParams& p, const Drop& dout, const Drop& din, int grid, cudaStream_t s) {
  constexpr int N = 128 * C4;
  const int n_acc = (p.dgamma ? 1 : 0) + (p.dbeta ? 1 : 0) + (p.dbias ? 1 : 0);
  const size_t smem = sizeof(float) * ((size_t)N + (size_t)LN_WARPS * n_acc * N);
  static size_t configured = 0;
  if (smem > configured) {
    CLG_CUDA_OK(cudaFuncSetAttribute(ln_bwd_vec_kernel<C4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "ln_bwd smem");
    configured = smem;
  }
  CLG_CUDA_OK(launch_pdl(ln_bwd_vec_kernel<C4>, dim3(grid), dim3(LN_THREADS), smem, s, p, dout, din), "launch(ln_bwd_vec)");
  return check_launch("ln_bwd_vec_kernel");
}

}  // namespace clg

extern "C" int clg_layernorm_fwd(const clgLayerNormFwdParams* p, clgStream_t stream) {
  using namespace clg;
  if (!p || !p->x || !p->gamma || !p->beta || p->rows < 0 || p->N <= 0 || p->N > LN_MAX_N) return CLG_ERR_ARG;
  int st = require_sm100();
  if (st != CLG_OK) return st;
  if (p->rows == 0) return CLG_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const Drop d = make_drop(p->dropout);
  int64_t blocks = (p->rows + LN_WARPS - 1) / LN_WARPS;
  const int64_t cap = (int64_t)num_sms() * 8;
  if (blocks > cap) blocks = cap;
  const int grid = (int)blocks;
  const bool vec = (p->N % 128 == 0) && (p->ldx % 4 == 0) && aligned16(p->x) && aligned16(p->gamma) && aligned16(p->beta) &&
                   (!p->y || ((p->ldy % 4 == 0) && aligned16(p->y))) &&
                   (!p->y_lowp || ((p->ldy_lowp % 4 == 0) && aligned16(p->y_lowp))) &&
                   (!p->add_mat || ((p->ld_add % 4 == 0) && aligned16(p->add_mat))) && (!p->add_vec || aligned16(p->add_vec));
  if (vec && p->N == 768) CLG_CUDA_OK(launch_pdl(ln_fwd_vec_kernel<6, 4>, dim3(grid), dim3(LN_THREADS), 0, s, *p, d), "launch(ln_fwd)");
  else if (vec && p->N == 1536) CLG_CUDA_OK(launch_pdl(ln_fwd_vec_kernel<12, 2>, dim3(grid), dim3(LN_THREADS), 0, s, *p, d), "launch(ln_fwd)");
  else if (vec && p->N == 128) CLG_CUDA_OK(launch_pdl(ln_fwd_vec_kernel<1, 4>, dim3(grid), dim3(LN_THREADS), 0, s, *p, d), "launch(ln_fwd)");
  else ln_fwd_gen_kernel<<<grid, LN_THREADS, 0, s>>>(*p, d);
  return check_launch("ln_fwd_kernel");
}

extern "C" int clg_layernorm_bwd(const clgLayerNormBwdParams* p, clgStream_t stream) {
  using namespace clg;
  if (!p || !p->dy || !p->x || !p->gamma || !p->mean || !p->rstd || p->rows < 0 || p->N <= 0 || p->N > LN_MAX_N)
    return CLG_ERR_ARG;
  int st = require_sm100();
  if (st != CLG_OK) return st;
  if (p->rows == 0) return CLG_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const Drop dout = make_drop(p->dropout_out), din = make_drop(p->dropout_in);
  int64_t blocks = (p->rows + LN_WARPS - 1) / LN_WARPS;
  const int64_t cap = (int64_t)num_sms() * 2;          // 2 resident CTAs / SM; each ends with N atomics per vector
  if (blocks > cap) blocks = cap;
  const int grid = (int)blocks;
  const bool vec = (p->N % 128 == 0) && (p->ldx % 4 == 0) && (p->lddy % 4 == 0) && aligned16(p->x) && aligned16(p->dy) &&
                   aligned16(p->gamma) && (!p->dx || ((p->lddx % 4 == 0) && aligned16(p->dx))) &&
                   (!p->dx_lowp || ((p->lddx_lowp % 4 == 0) && aligned16(p->dx_lowp)));
  if (vec && p->N == 768) return launch_bwd_vec<6>(*p, dout, din, grid, s);
  if (vec && p->N == 1536) return launch_bwd_vec<12>(*p, dout, din, grid, s);
  if (vec && p->N == 128) return launch_bwd_vec<1>(*p, dout, din, grid, s);
  ln_bwd_gen_kernel<<<grid, LN_THREADS, 0, s>>>(*p, dout, din);
  return check_launch("ln_bwd_gen_kernel");
}
