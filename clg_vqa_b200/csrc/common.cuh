// Shared device/host helpers for the clgvqa kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <math.h>
#include "../../include/clgvqa.h"

namespace clg {

typedef __nv_bfloat16 bf16;

// ---------------------------------------------------------------------------------------------
// host-side status plumbing
// ---------------------------------------------------------------------------------------------
void set_cuda_error(cudaError_t e, const char* where);
int check_launch(const char* where);   // cudaGetLastError -> status

#define CLG_CUDA_OK(expr, where)                                   \
  do {                                                             \
    cudaError_t _e = (expr);                                       \
    if (_e != cudaSuccess) { clg::set_cuda_error(_e, where); return CLG_ERR_CUDA; } \
  } while (0)

// ---------------------------------------------------------------------------------------------
// Programmatic dependent launch: the next kernel of the stream is launched while this one drains; its prologue
// (barrier init, TMEM allocation, descriptor prefetch, index math) overlaps our tail.  Device side: pdl_trigger() at
// the top lets the dependents be scheduled, pdl_wait() blocks until every prerequisite grid has completed and its
// memory is visible - it precedes the first global access of the kernel.
// ---------------------------------------------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
#endif
bool pdl_enabled();            // CLG_PDL=0 disables the launch attribute (kernels then run fully serialised)

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

int num_sms();                 // of the current device (cached)
int require_sm100();           // CLG_OK or CLG_ERR_ARCH for the current device (cached)

// ---------------------------------------------------------------------------------------------
// dropout: counter-based hash, shared by all kernels (forward and backward regenerate the mask)
// ---------------------------------------------------------------------------------------------
// Element idx of a site draws 16 random bits: the low / high half of lowbias32((idx >> 1) ^ key),
// key = 32-bit mix of (seed, site offset).  Kept iff bits >= p * 2^16 (p is quantised to 1/65536 and
// the kept values are scaled by the exact 1/(1 - p_quantised)).  ~8 integer instructions per element.
struct Drop {
  uint32_t thresh;   // keep iff 16-bit draw >= thresh
  float scale;       // 1/(1-p_q)
  uint32_t key;
  bool on;
};

__host__ __device__ __forceinline__ uint32_t lowbias32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352dU;
  x ^= x >> 15; x *= 0x846ca68bU;
  x ^= x >> 16;
  return x;
}

__host__ __device__ inline Drop make_drop(const clgDropout& d) {
  Drop r;
  r.on = d.p > 0.f;
  uint32_t t = r.on ? (uint32_t)(d.p * 65536.0f + 0.5f) : 0u;
  if (t > 65535u) t = 65535u;
  r.thresh = t;
  r.scale = r.on ? 65536.0f / (float)(65536u - t) : 1.0f;
  uint64_t z = (d.seed + 0x9E3779B97F4A7C15ULL) * 0xBF58476D1CE4E5B9ULL;
  z ^= (d.offset + 0x94D049BB133111EBULL) * 0xD6E8FEB86659FD93ULL;
  z ^= z >> 32;
  r.key = lowbias32((uint32_t)z);
  return r;
}

// keep-mask for element idx of a site (idx < 2^32 for every site of this model)
__host__ __device__ __forceinline__ bool drop_keep(const Drop& d, uint64_t idx) {
  const uint32_t i = (uint32_t)idx;
  const uint32_t h = lowbias32((i >> 1) ^ d.key);
  const uint32_t r = (i & 1u) ? (h >> 16) : (h & 0xffffu);
  return r >= d.thresh;
}

__host__ __device__ __forceinline__ float drop_apply(const Drop& d, uint64_t idx, float v) {
  if (!d.on) return v;
  return drop_keep(d, idx) ? v * d.scale : 0.f;
}

// ---------------------------------------------------------------------------------------------
// small device helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

__device__ __forceinline__ float gelu_erf(float x) {
  return x * 0.5f * (1.0f + erff(x * 0.70710678118654752440f));
}
__device__ __forceinline__ float dgelu_erf(float x) {
  const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752440f));
  const float pdf = 0.39894228040143267794f * expf(-0.5f * x * x);
  return cdf + x * pdf;
}

// gelu(u) and gelu'(u) together from ONE exponential: erf by Abramowitz-Stegun 7.1.26 (|abs err| <= 1.5e-7,
// branch-free), whose e^{-x^2} with x = u/sqrt(2) is also the Gaussian of gelu'(u) = Phi(u) + u phi(u).
__device__ __forceinline__ float rcp_approx(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float ex2_approx(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ void gelu_pair_fast(float u, float& g, float& dg) {
  // two MUFU ops (rcp, ex2) and ~14 FMA-pipe ops per element, no slow paths / calls: the epilogue of the FFN1
  // GEMM has to fit under the tile's tensor-core time
  const float ax = fabsf(u) * 0.70710678118654752440f;
  const float t = rcp_approx(fmaf(0.3275911f, ax, 1.0f));
  const float E = ex2_approx(u * u * -0.72134752044448170368f);        // e^{-u^2/2}
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  const float q = 0.5f * poly * t * E;                                 // (1 - erf(|x|)) / 2
  const float cdf = u >= 0.f ? 1.0f - q : q;                           // Phi(u)
  g = u * cdf;
  dg = fmaf(u * E, 0.39894228040143267794f, cdf);
}

template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<bf16>(bf16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 from_f32<bf16>(float v) { return __float2bfloat16_rn(v); }

__device__ __forceinline__ float load_as_f32(const void* p, int dtype, int64_t i) {
  return dtype == CLG_BF16 ? __bfloat162float(((const bf16*)p)[i]) : ((const float*)p)[i];
}
__device__ __forceinline__ void store_from_f32(void* p, int dtype, int64_t i, float v) {
  if (dtype == CLG_BF16) ((bf16*)p)[i] = __float2bfloat16_rn(v);
  else ((float*)p)[i] = v;
}

__host__ __device__ __forceinline__ int64_t map_row(const clgRowMap& m, int64_t r) {
  if (m.group_rows <= 0) return r;
  return (r / m.group_rows) * m.group_stride + m.row_offset + (r % m.group_rows);
}

__device__ __forceinline__ bool mask_bit(const uint32_t* bits, int64_t i) {
  return (bits[i >> 5] >> (i & 31)) & 1u;
}

static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

}  // namespace clg
