// Global magnitude pruning for the IMP loop (volta/train_task_prunning.py:L45-90: prune.global_unstructured with
// L1Unstructured over the 73 attention / FFN / pooler weights, amount = fraction of the REMAINING weights).
// torch builds one 342 MB score vector and runs topk(k = 8.5 M) on it; here the k-th smallest |w| among the unpruned
// entries is found by a 4-pass radix select straight over the weight tensors (HBM-bound: 4 B per weight per pass), and
// the new {0,1} masks are written in one more pass.  Non-negative fp32 values order like their bit patterns.
#include "common.cuh"

namespace clg {

struct PruneTensor { const float* w; const float* mask; float* mask_out; int64_t n; };

// histogram of the 8-bit digit at `shift` over the unpruned entries whose higher bits equal `prefix`
__global__ void __launch_bounds__(256)
prune_hist_kernel(const PruneTensor* __restrict__ tens, int n_tensors, uint32_t prefix, uint32_t prefix_mask, int shift,
                  unsigned long long* __restrict__ hist) {
  __shared__ unsigned int sh[256];
  sh[threadIdx.x] = 0;
  __syncthreads();
  for (int t = blockIdx.y; t < n_tensors; t += gridDim.y) {
    const PruneTensor pt = tens[t];
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < pt.n; i += (int64_t)gridDim.x * blockDim.x) {
      if (pt.mask && pt.mask[i] == 0.f) continue;
      const uint32_t key = __float_as_uint(fabsf(pt.w[i]));
      if ((key & prefix_mask) == prefix) atomicAdd(&sh[(key >> shift) & 0xffu], 1u);
    }
  }
  __syncthreads();
  if (sh[threadIdx.x]) atomicAdd(&hist[threadIdx.x], (unsigned long long)sh[threadIdx.x]);
}

// mask_out = 1 for unpruned entries with |w| > T, 0 for |w| < T; of the entries equal to T exactly `ties_to_prune` are
// pruned (which ones is arbitrary, as with topk)
__global__ void __launch_bounds__(256)
prune_apply_kernel(const PruneTensor* __restrict__ tens, int n_tensors, uint32_t thr_key, unsigned long long ties_to_prune,
                   unsigned long long* __restrict__ tie_counter) {
  for (int t = blockIdx.y; t < n_tensors; t += gridDim.y) {
    const PruneTensor pt = tens[t];
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < pt.n; i += (int64_t)gridDim.x * blockDim.x) {
      float m = pt.mask ? pt.mask[i] : 1.f;
      if (m != 0.f) {
        const uint32_t key = __float_as_uint(fabsf(pt.w[i]));
        if (key < thr_key) m = 0.f;
        else if (key == thr_key && ties_to_prune > 0) {
          if (atomicAdd(tie_counter, 1ULL) < ties_to_prune) m = 0.f;
        } else m = 1.f;
      }
      pt.mask_out[i] = m;
    }
  }
}

}  // namespace clg

extern "C" int clg_prune_hist(const void* tensors, int32_t n_tensors, uint32_t prefix, uint32_t prefix_mask, int32_t shift,
                              uint64_t* hist256, clgStream_t stream) {
  using namespace clg;
  if (!tensors || !hist256 || n_tensors <= 0 || shift < 0 || shift > 24) return CLG_ERR_ARG;
  int st = require_sm100();
  if (st != CLG_OK) return st;
  dim3 grid(num_sms() * 2, n_tensors < 8 ? n_tensors : 8);
  prune_hist_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>((const PruneTensor*)tensors, n_tensors, prefix, prefix_mask, shift,
                                                                        reinterpret_cast<unsigned long long*>(hist256));
  return check_launch("prune_hist_kernel");
}

extern "C" int clg_prune_apply(const void* tensors, int32_t n_tensors, uint32_t thr_key, uint64_t ties_to_prune, uint64_t* tie_counter,
                               clgStream_t stream) {
  using namespace clg;
  if (!tensors || !tie_counter || n_tensors <= 0) return CLG_ERR_ARG;
  int st = require_sm100();
  if (st != CLG_OK) return st;
  dim3 grid(num_sms() * 2, n_tensors < 8 ? n_tensors : 8);
  prune_apply_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>((const PruneTensor*)tensors, n_tensors, thr_key,
                                                                         (unsigned long long)ties_to_prune,
                                                                         reinterpret_cast<unsigned long long*>(tie_counter));
  return check_launch("prune_apply_kernel");
}

extern "C" size_t clg_sizeof_prune_tensor(void) { return sizeof(clg::PruneTensor); }
