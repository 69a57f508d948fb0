// LayerNorm forward / backward (one warp per row, rows kept in registers, two-pass variance).
// HBM-bound: forward reads x once and writes y (+ the bf16 GEMM-operand copy); backward reads dy and x
// once and writes dx (+ operand copy) while accumulating dgamma / dbeta / dbias column sums in
// registers across the rows a warp owns, then one smem reduction and one atomicAdd per column per CTA.
#include "common.cuh"

namespace clg {

static constexpr int LN_MAX_PER_LANE = 48;   // N <= 32 * 48 = 1536 (M3P classifier hidden)
static constexpr int LN_WARPS = 8;

template <int PER>   // PER = ceil(N / 32) rounded to the instantiated sizes
__global__ void __launch_bounds__(LN_WARPS * 32)
ln_fwd_kernel(const clgLayerNormFwdParams p, const Drop drop) {
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int N = p.N;
  const float invN = 1.0f / (float)N;
  for (int64_t r = (int64_t)blockIdx.x * LN_WARPS + warp; r < p.rows; r += (int64_t)gridDim.x * LN_WARPS) {
    const float* xr = p.x + r * p.ldx;
    const float pre = (p.row_valid && p.valid_pre) ? (p.row_valid[r] ? 1.f : 0.f) : 1.f;
    const float post = (p.row_valid && p.valid_post) ? (p.row_valid[r] ? 1.f : 0.f) : 1.f;
    float v[PER];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < PER; ++i) {
      const int c = lane + i * 32;
      v[i] = (c < N) ? xr[c] * pre : 0.f;
      s += v[i];
    }
    const float mean = warp_sum(s) * invN;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < PER; ++i) {
      const int c = lane + i * 32;
      const float d = (c < N) ? v[i] - mean : 0.f;
      q += d * d;
    }
    const float var = warp_sum(q) * invN;
    const float rstd = 1.0f / sqrtf(var + p.eps);
    if (lane == 0) {
      if (p.mean) p.mean[r] = mean;
      if (p.rstd) p.rstd[r] = rstd;
    }
    const int64_t orow = map_row(p.out_map, r);
    const float* addr = p.add_mat ? p.add_mat + map_row(p.add_map, r) * p.ld_add : nullptr;
#pragma unroll
    for (int i = 0; i < PER; ++i) {
      const int c = lane + i * 32;
      if (c < N) {
        float y = ((v[i] - mean) * rstd) * p.gamma[c] + p.beta[c];
        y *= post;
        if (addr) y += addr[c];
        if (p.add_vec) y += p.add_vec[c];
        y = drop_apply(drop, (uint64_t)r * (uint64_t)N + c, y);
        if (p.y) p.y[orow * p.ldy + c] = y;
        if (p.y_lowp) store_from_f32(p.y_lowp, p.lowp_dtype, orow * p.ldy_lowp + c, y);
      }
    }
  }
}

template <int PER>
__global__ void __launch_bounds__(LN_WARPS * 32)
ln_bwd_kernel(const clgLayerNormBwdParams p, const Drop drop_out, const Drop drop_in) {
  __shared__ float red[LN_WARPS][32 * PER];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int N = p.N;
  const float invN = 1.0f / (float)N;
  float ag[PER], ab[PER], abias[PER];
#pragma unroll
  for (int i = 0; i < PER; ++i) { ag[i] = 0.f; ab[i] = 0.f; abias[i] = 0.f; }
  float gam[PER];
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    const int c = lane + i * 32;
    gam[i] = (c < N) ? p.gamma[c] : 0.f;
  }
  for (int64_t r = (int64_t)blockIdx.x * LN_WARPS + warp; r < p.rows; r += (int64_t)gridDim.x * LN_WARPS) {
    const float pre = (p.row_valid && p.valid_pre) ? (p.row_valid[r] ? 1.f : 0.f) : 1.f;
    const float post = (p.row_valid && p.valid_post) ? (p.row_valid[r] ? 1.f : 0.f) : 1.f;
    const float* dyr = p.dy + map_row(p.dy_map, r) * p.lddy;
    const float* xr = p.x + r * p.ldx;
    const float mean = p.mean[r], rstd = p.rstd[r];
    float g[PER], xh[PER];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < PER; ++i) {
      const int c = lane + i * 32;
      if (c < N) {
        float d = drop_apply(drop_out, (uint64_t)r * (uint64_t)N + c, dyr[c]) * post;
        xh[i] = (xr[c] * pre - mean) * rstd;
        ag[i] += d * xh[i];
        ab[i] += d;
        g[i] = d * gam[i];
        s1 += g[i];
        s2 += g[i] * xh[i];
      } else { g[i] = 0.f; xh[i] = 0.f; }
    }
    s1 = warp_sum(s1) * invN;
    s2 = warp_sum(s2) * invN;
#pragma unroll
    for (int i = 0; i < PER; ++i) {
      const int c = lane + i * 32;
      if (c < N) {
        const float dx = rstd * (g[i] - s1 - xh[i] * s2) * pre;
        if (p.dx) p.dx[r * p.lddx + c] = dx;
        const float dxin = drop_apply(drop_in, (uint64_t)r * (uint64_t)N + c, dx);
        if (p.dx_lowp) store_from_f32(p.dx_lowp, p.lowp_dtype, r * p.lddx_lowp + c, dxin);
        abias[i] += dxin;
      }
    }
  }
  // column reductions across the warps of the CTA, then one atomic per column
  float* outs[3] = {p.dgamma, p.dbeta, p.dbias};
#pragma unroll
  for (int which = 0; which < 3; ++which) {
    if (!outs[which]) continue;      // uniform across the CTA
    __syncthreads();
#pragma unroll
    for (int i = 0; i < PER; ++i) red[warp][lane + i * 32] = which == 0 ? ag[i] : (which == 1 ? ab[i] : abias[i]);
    __syncthreads();
    for (int c = threadIdx.x; c < N; c += LN_WARPS * 32) {
      float s = 0.f;
#pragma unroll
      for (int w = 0; w < LN_WARPS; ++w) s += red[w][c];
      atomicAdd(outs[which] + c, s);
    }
  }
}

template <typename P> static int ln_grid(const P& p) {
  int64_t blocks = (p.rows + LN_WARPS - 1) / LN_WARPS;
  const int64_t cap = (int64_t)num_sms() * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

}  // namespace clg

extern "C" int clg_layernorm_fwd(const clgLayerNormFwdParams* p, clgStream_t stream) {
  using namespace clg;
  if (!p || !p->x || !p->gamma || !p->beta || p->rows < 0 || p->N <= 0 || p->N > 32 * LN_MAX_PER_LANE) return CLG_ERR_ARG;
  int st = require_sm100();
  if (st != CLG_OK) return st;
  if (p->rows == 0) return CLG_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const Drop d = make_drop(p->dropout);
  const int grid = ln_grid(*p);
  const int per = (p->N + 31) / 32;
  if (per <= 24) ln_fwd_kernel<24><<<grid, LN_WARPS * 32, 0, s>>>(*p, d);
  else ln_fwd_kernel<48><<<grid, LN_WARPS * 32, 0, s>>>(*p, d);
  return check_launch("ln_fwd_kernel");
}

extern "C" int clg_layernorm_bwd(const clgLayerNormBwdParams* p, clgStream_t stream) {
  using namespace clg;
  if (!p || !p->dy || !p->x || !p->gamma || !p->mean || !p->rstd || p->rows < 0 || p->N <= 0 ||
      p->N > 32 * LN_MAX_PER_LANE)
    return CLG_ERR_ARG;
  int st = require_sm100();
  if (st != CLG_OK) return st;
  if (p->rows == 0) return CLG_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const Drop dout = make_drop(p->dropout_out), din = make_drop(p->dropout_in);
  // fewer, fatter CTAs: every CTA ends with N atomics per accumulated vector
  int64_t blocks = (p->rows + LN_WARPS - 1) / LN_WARPS;
  const int64_t cap = (int64_t)num_sms() * 2;
  if (blocks > cap) blocks = cap;
  const int per = (p->N + 31) / 32;
  if (per <= 24) ln_bwd_kernel<24><<<(int)blocks, LN_WARPS * 32, 0, s>>>(*p, dout, din);
  else ln_bwd_kernel<48><<<(int)blocks, LN_WARPS * 32, 0, s>>>(*p, dout, din);
  return check_launch("ln_bwd_kernel");
}
