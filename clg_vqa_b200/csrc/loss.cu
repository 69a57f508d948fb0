// GQA answer-head loss with the label-prior term (task_utils.py:L413-428), score (L706-711) and
// eval argmax (L831-841) in one pass over the logits: one CTA per sample keeps the 1842-logit row in
// shared memory, does max / argmax / log-sum-exp / top-k / prior term with warp-shuffle reductions
// and writes d(loss)/d(logits) in the same kernel.  HBM traffic = logits + target + distances read
// once, dlogits written once.
#include <limits.h>
#include "common.cuh"

namespace clg {

static constexpr int LOSS_THREADS = 128;
static constexpr int LOSS_WARPS = LOSS_THREADS / 32;
static constexpr int MAX_TOPK = 16;

struct ValIdx { float v; int i; };

__device__ __forceinline__ ValIdx better(ValIdx a, ValIdx b) {
  // larger value wins, smaller index on ties (first occurrence, like torch's CPU argmax / max)
  if (b.v > a.v || (b.v == a.v && b.i < a.i)) return b;
  return a;
}

__device__ __forceinline__ ValIdx block_argmax(ValIdx x, ValIdx* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    ValIdx y;
    y.v = __shfl_xor_sync(0xffffffffu, x.v, o);
    y.i = __shfl_xor_sync(0xffffffffu, x.i, o);
    x = better(x, y);
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) red[warp] = x;
  __syncthreads();
  ValIdx r = red[0];
#pragma unroll
  for (int w = 1; w < LOSS_WARPS; ++w) r = better(r, red[w]);
  return r;
}

__device__ __forceinline__ float block_sum(float x, float* red) {
  x = warp_sum(x);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) red[warp] = x;
  __syncthreads();
  float r = 0.f;
#pragma unroll
  for (int w = 0; w < LOSS_WARPS; ++w) r += red[w];
  return r;
}

__global__ void __launch_bounds__(LOSS_THREADS)
gqa_loss_row_kernel(const clgGqaLossParams p) {
  extern __shared__ float sm[];
  float* z = sm;              // [A] logits
  float* dk = sm + p.A;       // [A] dist restricted to the top-k set (0 elsewhere)
  __shared__ ValIdx red_vi[LOSS_WARPS];
  __shared__ float red_f[LOSS_WARPS];
  __shared__ int topk_idx[MAX_TOPK];
  const int b = blockIdx.x, tid = threadIdx.x, A = p.A;
  const float* zg = p.logits + (int64_t)b * p.ld_logits;
  const float* tg = p.target + (int64_t)b * p.ld_target;
  const float* dg = p.dist ? p.dist + (int64_t)b * p.ld_dist : nullptr;

  ValIdx best = {-INFINITY, 0x7fffffff};
  long long tbest_v = LLONG_MIN; int tbest_i = 0x7fffffff;
  for (int i = tid; i < A; i += LOSS_THREADS) {
    const float v = zg[i];
    z[i] = v;
    dk[i] = 0.f;
    best = better(best, ValIdx{v, i});
    const long long tv = (long long)tg[i];                    // target.long(): truncation
    if (tv > tbest_v || (tv == tbest_v && i < tbest_i)) { tbest_v = tv; tbest_i = i; }
  }
  const ValIdx top1 = block_argmax(best, red_vi);
  // argmax of the truncated target: encode as (value, -index) comparisons through the same reducer
  ValIdx ty = {(float)tbest_v, tbest_i};
  ty = block_argmax(ty, red_vi);
  const int y = ty.i;
  const float zmax = top1.v;

  float se = 0.f;
  for (int i = tid; i < A; i += LOSS_THREADS) se += expf(z[i] - zmax);
  const float sumexp = block_sum(se, red_f);
  const float inv = 1.0f / sumexp;
  const float nll = -(z[y] - zmax - logf(sumexp));

  // top-k by repeated block argmax; a thread owns elements tid + 128*t and masks the taken ones
  float s = 0.f;
  if (dg) {
    uint32_t taken = 0;
    for (int k = 0; k < p.topk; ++k) {
      ValIdx loc = {-INFINITY, 0x7fffffff};
      int t = 0;
      for (int i = tid; i < A; i += LOSS_THREADS, ++t)
        if (!((taken >> t) & 1u)) loc = better(loc, ValIdx{z[i], i});
      const ValIdx w = block_argmax(loc, red_vi);
      if ((w.i % LOSS_THREADS) == tid) {
        taken |= 1u << (w.i / LOSS_THREADS);
        topk_idx[k] = w.i;
      }
    }
    __syncthreads();
    if (tid < p.topk) {
      const int i = topk_idx[tid];
      const float d = dg[i];
      dk[i] = d;
      s = expf(z[i] - zmax) * inv * d;
    }
    s = block_sum(s, red_f);     // also orders the dk writes before the reads below
  }

  if (tid == 0) {
    float* rb = p.row_buf + (int64_t)b * 4;
    rb[0] = nll; rb[1] = s; rb[2] = tg[top1.i]; rb[3] = (float)top1.i;
    if (p.pred) p.pred[b] = top1.i;
  }
  if (p.dlogits) {
    float* og = p.dlogits + (int64_t)b * p.ld_dlogits;
    const float c = p.gscale * (float)A / (float)p.B;
    const float lam = dg ? p.lambda : 0.f;
    for (int i = tid; i < A; i += LOSS_THREADS) {
      const float pi = expf(z[i] - zmax) * inv;
      float g = pi - (i == y ? 1.f : 0.f);
      g += lam * pi * (dk[i] - s);
      og[i] = c * g;
    }
  }
}

__global__ void __launch_bounds__(256)
gqa_loss_reduce_kernel(const clgGqaLossParams p) {
  __shared__ double sh[3][256];
  double a = 0, s = 0, sc = 0;
  for (int b = threadIdx.x; b < p.B; b += 256) {   // fixed order per thread, fixed tree below: deterministic
    const float* rb = p.row_buf + (int64_t)b * 4;
    a += rb[0]; s += rb[1]; sc += rb[2];
  }
  sh[0][threadIdx.x] = a; sh[1][threadIdx.x] = s; sh[2][threadIdx.x] = sc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      sh[0][threadIdx.x] += sh[0][threadIdx.x + o];
      sh[1][threadIdx.x] += sh[1][threadIdx.x + o];
      sh[2][threadIdx.x] += sh[2][threadIdx.x + o];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const double B = (double)p.B, A = (double)p.A;
    const double lam = p.dist ? (double)p.lambda : 0.0;
    p.out[0] = (float)(A * (sh[0][0] / B + lam * sh[1][0] / B));
    p.out[1] = (float)(p.score_mean ? sh[2][0] / B : sh[2][0]);
  }
}

}  // namespace clg

extern "C" int clg_gqa_prior_loss(const clgGqaLossParams* p, clgStream_t stream) {
  using namespace clg;
  if (!p || !p->logits || !p->target || !p->row_buf || !p->out || p->B <= 0 || p->A <= 0 || p->topk < 0 ||
      p->topk > MAX_TOPK || p->topk > p->A || p->A > 32 * LOSS_THREADS)
    return CLG_ERR_ARG;
  int st = require_sm100();
  if (st != CLG_OK) return st;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const size_t smem = sizeof(float) * 2 * (size_t)p->A;
  gqa_loss_row_kernel<<<p->B, LOSS_THREADS, smem, s>>>(*p);
  st = check_launch("gqa_loss_row_kernel");
  if (st != CLG_OK) return st;
  gqa_loss_reduce_kernel<<<1, 256, 0, s>>>(*p);
  return check_launch("gqa_loss_reduce_kernel");
}
