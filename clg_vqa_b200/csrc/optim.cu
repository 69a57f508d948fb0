// Optimiser step over flat fp32 arenas: global gradient norm (two deterministic passes), then one
// multi-tensor kernel doing clip-scale + pytorch-transformers AdamW +
// decoupled decay, and refreshing the masked bf16 shadow copy the GEMMs read (W_eff = W (.) M, the
// torch.nn.utils.prune re-parametrisation of train_task_sft.py:L461 folded into the update).
// HBM-bound: 7 fp32 accesses per parameter (read p, g, m, v; write p, m, v) + 2 B shadow write.
#include <string.h>
#include "common.cuh"

namespace clg {

static constexpr int CHUNK = 2048;
static constexpr int OPT_THREADS = 256;   // 8 elements per thread per chunk

// one warp per 256-element block (8 consecutive floats per lane): the block's "has a non-zero gradient" bit comes for
// free with the norm.  Blocks [sp_lo, sp_hi) (the word-embedding table: 192 M of the 281 M parameters, of which a batch
// touches a few thousand rows) are read only where `touched` marks them (clg_mark_rows, from the token ids of the
// backward): every other block of that range is exactly zero and contributes nothing to the norm.
__device__ __forceinline__ float sumsq_block(const float4* g4, int64_t b, int lane, uint8_t* hot) {
  const float4 v0 = __ldg(g4 + b * 64 + lane * 2);
  const float4 v1 = __ldg(g4 + b * 64 + lane * 2 + 1);
  if (hot) {
    const bool nz = (v0.x != 0.f) | (v0.y != 0.f) | (v0.z != 0.f) | (v0.w != 0.f) | (v1.x != 0.f) | (v1.y != 0.f) |
                    (v1.z != 0.f) | (v1.w != 0.f);
    if (__any_sync(0xffffffffu, nz) && lane == 0) hot[b] = 1;
  }
  return v0.x * v0.x + v0.y * v0.y + v0.z * v0.z + v0.w * v0.w + v1.x * v1.x + v1.y * v1.y + v1.z * v1.z + v1.w * v1.w;
}

__global__ void __launch_bounds__(256)
sumsq_partial_kernel(const float* __restrict__ g, int64_t n, float* __restrict__ partials, uint8_t* __restrict__ hot,
                     int64_t sp_lo, int64_t sp_hi, const uint8_t* __restrict__ touched) {
  __shared__ float red[8];
  float acc = 0.f;
  const int lane = threadIdx.x & 31;
  const int64_t nblk = n >> 8;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const float4* g4 = reinterpret_cast<const float4*>(g);
  if (!touched) sp_lo = sp_hi = 0;
  const int64_t ndense = nblk - (sp_hi - sp_lo);
  for (int64_t i = warp0; i < ndense; i += nwarps) {
    const int64_t b = i < sp_lo ? i : i + (sp_hi - sp_lo);
    acc += sumsq_block(g4, b, lane, hot);
  }
  // sparse range: 32 blocks per warp iteration, one mark per lane
  const int64_t ngroups = (sp_hi - sp_lo + 31) >> 5;
  for (int64_t gi = warp0; gi < ngroups; gi += nwarps) {
    const int64_t b = sp_lo + gi * 32 + lane;
    unsigned live = __ballot_sync(0xffffffffu, b < sp_hi && touched[b] != 0);
    while (live) {
      const int l = __ffs(live) - 1;
      live &= live - 1;
      acc += sumsq_block(g4, sp_lo + gi * 32 + l, lane, hot);
    }
  }
  if (blockIdx.x == 0)
    for (int64_t i = (nblk << 8) + threadIdx.x; i < n; i += blockDim.x) acc += g[i] * g[i];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int w = 0; w < 8; ++w) s += red[w];
    partials[blockIdx.x] = s;
  }
}

// touched[(first_elem + id * row_elems) >> 8 ...] = 1 for every 256-element block the row of token `id` overlaps
__global__ void __launch_bounds__(256)
mark_rows_kernel(const int64_t* __restrict__ ids, int64_t n, int64_t row_elems, int64_t first_elem, int64_t num_rows,
                 int64_t skip_id, uint8_t* __restrict__ touched) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t id = ids[i];
  if (id < 0 || id >= num_rows || id == skip_id) return;
  const int64_t e0 = first_elem + id * row_elems, e1 = e0 + row_elems - 1;
  for (int64_t b = e0 >> 8; b <= (e1 >> 8); ++b) touched[b] = 1;
}

__global__ void __launch_bounds__(256)
sumsq_final_kernel(const float* __restrict__ partials, int n, float* __restrict__ out) {
  __shared__ double sh[256];
  double a = 0;
  for (int i = threadIdx.x; i < n; i += 256) a += (double)partials[i];
  sh[threadIdx.x] = a;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] += (float)sh[0];
}

// One 2048-element chunk of tensor `tm` (8 consecutive elements per thread).  `cold`: the warp's 256-element block has
// g = m = v = 0 (never held a non-zero gradient).
template <bool UPDATE>
__device__ __forceinline__ void adamw_chunk(const clgAdamWParams& p, const int64_t c, const clgTensorMeta& tm, const bool cold,
                                            const float coef) {
  const int64_t base = c * CHUNK + (int64_t)threadIdx.x * 8;
  const int64_t rel = base - tm.start;          // element offset inside the tensor
  if (rel >= tm.numel) return;                  // tail padding of the tensor's last chunk (never written)
  // g = m = v = 0: the update reduces to the decoupled decay p <- fma(-lr*wd, p, p).  Below 2^-25 the product is under
  // half an ulp of p, so the rounded result IS p (bit-identical to the dense update): with the reference's settings
  // (lr <= 1e-4, weight decay 1e-4, train_task.py:L249-267) the untouched word-embedding rows cost no traffic at all
  if (cold && !(tm.lr * tm.wd >= 2.98023224e-8f)) return;
  float pv[8];
  {
    const float4 a = *reinterpret_cast<const float4*>(p.p + base);
    const float4 b = *reinterpret_cast<const float4*>(p.p + base + 4);
    pv[0] = a.x; pv[1] = a.y; pv[2] = a.z; pv[3] = a.w; pv[4] = b.x; pv[5] = b.y; pv[6] = b.z; pv[7] = b.w;
  }
  if (cold) {
    // the dense formulas with g = m = v = 0 reduce to the decoupled decay alone (bit-identical: p - step*(0/eps) == p)
    const float decay = tm.lr * tm.wd;
#pragma unroll
    for (int j = 0; j < 8; ++j) pv[j] = pv[j] - decay * pv[j];
    *reinterpret_cast<float4*>(p.p + base) = make_float4(pv[0], pv[1], pv[2], pv[3]);
    *reinterpret_cast<float4*>(p.p + base + 4) = make_float4(pv[4], pv[5], pv[6], pv[7]);
  } else if (UPDATE) {
    float gv[8], mv[8], vv[8];
    {
      const float4 a = *reinterpret_cast<const float4*>(p.g + base);        // plain loads: g may be cleared below
      const float4 b = *reinterpret_cast<const float4*>(p.g + base + 4);
      gv[0] = a.x; gv[1] = a.y; gv[2] = a.z; gv[3] = a.w; gv[4] = b.x; gv[5] = b.y; gv[6] = b.z; gv[7] = b.w;
      const float4 c0 = *reinterpret_cast<const float4*>(p.m + base);
      const float4 c1 = *reinterpret_cast<const float4*>(p.m + base + 4);
      mv[0] = c0.x; mv[1] = c0.y; mv[2] = c0.z; mv[3] = c0.w; mv[4] = c1.x; mv[5] = c1.y; mv[6] = c1.z; mv[7] = c1.w;
      const float4 d0 = *reinterpret_cast<const float4*>(p.v + base);
      const float4 d1 = *reinterpret_cast<const float4*>(p.v + base + 4);
      vv[0] = d0.x; vv[1] = d0.y; vv[2] = d0.z; vv[3] = d0.w; vv[4] = d1.x; vv[5] = d1.y; vv[6] = d1.z; vv[7] = d1.w;
    }
    const float step_size = tm.lr * sqrtf(p.bias_c2) / p.bias_c1;
    const float decay = tm.lr * tm.wd;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float g = gv[j] * coef;
      mv[j] = mv[j] * p.beta1 + (1.f - p.beta1) * g;
      vv[j] = vv[j] * p.beta2 + (1.f - p.beta2) * g * g;
      const float denom = sqrtf(vv[j]) + p.eps;
      float w = pv[j] - step_size * (mv[j] / denom);
      if (tm.wd > 0.f) w = w - decay * w;
      pv[j] = w;
    }
    *reinterpret_cast<float4*>(p.p + base) = make_float4(pv[0], pv[1], pv[2], pv[3]);
    *reinterpret_cast<float4*>(p.p + base + 4) = make_float4(pv[4], pv[5], pv[6], pv[7]);
    *reinterpret_cast<float4*>(p.m + base) = make_float4(mv[0], mv[1], mv[2], mv[3]);
    *reinterpret_cast<float4*>(p.m + base + 4) = make_float4(mv[4], mv[5], mv[6], mv[7]);
    *reinterpret_cast<float4*>(p.v + base) = make_float4(vv[0], vv[1], vv[2], vv[3]);
    *reinterpret_cast<float4*>(p.v + base + 4) = make_float4(vv[4], vv[5], vv[6], vv[7]);
    if (p.zero_grad) {                                                        // after every load of this thread's g
      float* gz = const_cast<float*>(p.g) + base;
      const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
      *reinterpret_cast<float4*>(gz) = z4;
      *reinterpret_cast<float4*>(gz + 4) = z4;
    }
  }
  if (p.shadow && tm.shadow_off >= 0) {
    if (tm.mask_off >= 0) {
      // 8 consecutive mask bits; rel is a multiple of 8 and mask_off a multiple of 32
      const int64_t bit0 = tm.mask_off + rel;
      const uint32_t bits = (p.mask_bits[bit0 >> 5] >> (bit0 & 31)) & 0xffu;
#pragma unroll
      for (int j = 0; j < 8; ++j) pv[j] *= ((bits >> j) & 1u) ? 1.f : 0.f;
    }
    __nv_bfloat162 h0 = __floats2bfloat162_rn(pv[0], pv[1]);
    __nv_bfloat162 h1 = __floats2bfloat162_rn(pv[2], pv[3]);
    __nv_bfloat162 h2 = __floats2bfloat162_rn(pv[4], pv[5]);
    __nv_bfloat162 h3 = __floats2bfloat162_rn(pv[6], pv[7]);
    uint4 u;
    u.x = *reinterpret_cast<uint32_t*>(&h0); u.y = *reinterpret_cast<uint32_t*>(&h1);
    u.z = *reinterpret_cast<uint32_t*>(&h2); u.w = *reinterpret_cast<uint32_t*>(&h3);
    *reinterpret_cast<uint4*>((bf16*)p.shadow + tm.shadow_off + rel) = u;
  }
}

// A CTA walks the chunks c = blockIdx.x, + gridDim.x, ...  The per-chunk header (tensor index -> clgTensorMeta, the 8
// hot bytes of the chunk's 256-element blocks) is a chain of three dependent loads; two thirds of the chunks of the UC2
// arena are cold word-table blocks for which that chain is all there is.  The headers of HDR chunks are therefore fetched
// at once - lane j of warp 0 takes the j-th chunk of the batch - into shared memory, and the loop over the batch then
// reads no global memory but the arenas themselves.
static constexpr int HDR = 32;

template <bool UPDATE>
__global__ void __launch_bounds__(OPT_THREADS)
adamw_kernel(const clgAdamWParams p) {
  __shared__ clgTensorMeta s_tm[HDR];
  __shared__ unsigned long long s_hot[HDR];
  __shared__ int s_t[HDR];
  const int64_t nchunks = p.n / CHUNK;
  float coef = p.gscale;
  if (UPDATE && p.sumsq) {
    const float total = sqrtf(p.sumsq[0]) * p.gscale;
    const float c = p.max_norm / (total + 1e-6f);
    if (c < 1.f) coef *= c;
  }
  const int warp = threadIdx.x >> 5;
  for (int64_t c0 = blockIdx.x; c0 < nchunks; c0 += (int64_t)HDR * gridDim.x) {
    if (threadIdx.x < HDR) {
      const int64_t c = c0 + (int64_t)threadIdx.x * gridDim.x;
      int t = -2;                                   // -2: past the end of the arena
      unsigned long long h = ~0ull;
      if (c < nchunks) {
        t = __ldg(p.chunk_tensor + c);
        if (UPDATE && p.hot) h = __ldg(reinterpret_cast<const unsigned long long*>(p.hot) + c);   // CHUNK / 256 = 8 bytes
        if (t >= 0) s_tm[threadIdx.x] = p.tensors[t];
      }
      s_t[threadIdx.x] = t;
      s_hot[threadIdx.x] = h;
    }
    __syncthreads();
#pragma unroll 1
    for (int j = 0; j < HDR; ++j) {
      const int t = s_t[j];
      if (t == -2) break;
      const int64_t c = c0 + (int64_t)j * gridDim.x;
      if (t < 0) {                                  // no optimiser entry: its gradient (if any was written) is only cleared
        if (UPDATE && p.zero_grad) {
          float* gz = const_cast<float*>(p.g) + c * CHUNK + (int64_t)threadIdx.x * 8;
          const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
          *reinterpret_cast<float4*>(gz) = z4;
          *reinterpret_cast<float4*>(gz + 4) = z4;
        }
        continue;
      }
      // warp-uniform: a warp covers one 256-element block (byte `warp` of the chunk's hot word)
      const bool cold = UPDATE && ((s_hot[j] >> (8 * warp)) & 0xffull) == 0ull;
      adamw_chunk<UPDATE>(p, c, s_tm[j], cold, coef);
    }
    __syncthreads();                                // the batch is consumed: its header slots may be rewritten
  }
}

}  // namespace clg

extern "C" int clg_sumsq_sparse(const float* g, int64_t n, float* partials, int32_t num_partials, float* out, uint8_t* hot,
                                int64_t sparse_lo, int64_t sparse_hi, const uint8_t* touched, clgStream_t stream) {
  using namespace clg;
  if (!g || !partials || !out || n < 0 || num_partials <= 0) return CLG_ERR_ARG;
  if (touched && (sparse_lo < 0 || sparse_hi < sparse_lo || sparse_hi > (n >> 8))) return CLG_ERR_ARG;
  int st = require_sm100();
  if (st != CLG_OK) return st;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  int grid = num_sms() * 4;
  if (grid > num_partials) grid = num_partials;
  sumsq_partial_kernel<<<grid, 256, 0, s>>>(g, n, partials, hot, sparse_lo, sparse_hi, touched);
  st = check_launch("sumsq_partial_kernel");
  if (st != CLG_OK) return st;
  sumsq_final_kernel<<<1, 256, 0, s>>>(partials, grid, out);
  return check_launch("sumsq_final_kernel");
}

extern "C" int clg_sumsq(const float* g, int64_t n, float* partials, int32_t num_partials, float* out, uint8_t* hot,
                         clgStream_t stream) {
  return clg_sumsq_sparse(g, n, partials, num_partials, out, hot, 0, 0, nullptr, stream);
}

extern "C" int clg_mark_rows(const int64_t* ids, int64_t n, int64_t row_elems, int64_t first_elem, int64_t num_rows,
                             int64_t skip_id, uint8_t* touched, clgStream_t stream) {
  using namespace clg;
  if (!ids || !touched || n < 0 || row_elems <= 0 || first_elem < 0 || num_rows < 0) return CLG_ERR_ARG;
  int st = require_sm100();
  if (st != CLG_OK) return st;
  if (n == 0) return CLG_OK;
  mark_rows_kernel<<<(unsigned)((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(ids, n, row_elems, first_elem,
                                                                                                 num_rows, skip_id, touched);
  return check_launch("mark_rows_kernel");
}

extern "C" int clg_adamw_step(const clgAdamWParams* p, clgStream_t stream) {
  using namespace clg;
  if (!p || !p->p || !p->g || !p->m || !p->v || !p->chunk_tensor || !p->tensors || p->n <= 0 || (p->n % CHUNK)) return CLG_ERR_ARG;
  if (reinterpret_cast<uintptr_t>(p->hot) & 7) return CLG_ERR_ARG;      // the kernel reads a chunk's 8 hot bytes as one word
  int st = require_sm100();
  if (st != CLG_OK) return st;
  const int64_t nchunks = p->n / CHUNK;
  int64_t grid = (int64_t)num_sms() * 16;
  if (grid > nchunks) grid = nchunks;
  adamw_kernel<true><<<(unsigned)grid, OPT_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(*p);
  return check_launch("adamw_kernel");
}

extern "C" int clg_cast_shadow(const float* params, int64_t n, const int32_t* chunk_tensor, const clgTensorMeta* tensors,
                               void* shadow, const uint32_t* mask_bits, clgStream_t stream) {
  using namespace clg;
  if (!params || !chunk_tensor || !tensors || !shadow || n <= 0 || (n % CHUNK)) return CLG_ERR_ARG;
  int st = require_sm100();
  if (st != CLG_OK) return st;
  clgAdamWParams p;
  memset(&p, 0, sizeof(p));
  p.p = const_cast<float*>(params);
  p.n = n; p.chunk_tensor = chunk_tensor; p.tensors = tensors; p.shadow = shadow; p.mask_bits = mask_bits;
  p.gscale = 1.f;
  const int64_t nchunks = n / CHUNK;
  int64_t grid = (int64_t)num_sms() * 16;
  if (grid > nchunks) grid = nchunks;
  adamw_kernel<false><<<(unsigned)grid, OPT_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(p);
  return check_launch("cast_shadow_kernel");
}
