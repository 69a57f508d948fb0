// fp32-FFMA GEMM: the fp32 correctness mode and the small / oddly shaped GEMMs of the head
// (pooler, classifier, K = 7 location projection).  Operands may be fp32 or bf16 (converted exactly
// to fp32 on load), any major-ness, any leading dimension.  Shares the epilogue with gemm_tc.cu.
#include "gemm_epilogue.cuh"

namespace clg {

static constexpr int SBM = 64, SBN = 128, SBK = 16;

template <typename TA, typename TB>
__global__ void __launch_bounds__(256)
gemm_simt_kernel(const TA* __restrict__ A, const TB* __restrict__ B, int M, int N, int K, int64_t lda, int64_t ldb,
                 int a_k, int b_k, int k_per_split, const EpiArgs e) {
  __shared__ float As[SBK][SBM + 4];
  __shared__ float Bs[SBK][SBN + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * SBM, n0 = blockIdx.x * SBN;
  float acc[4][8];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  const int k_lo = blockIdx.z * k_per_split;
  const int k_hi = min(K, k_lo + k_per_split);
  K = k_hi;
  for (int k0 = k_lo; k0 < k_hi; k0 += SBK) {
    // A tile: 64 x 16
#pragma unroll
    for (int it = 0; it < (SBM * SBK) / 256; ++it) {
      const int i = tid + it * 256;
      int m, k;
      if (a_k) { k = i % SBK; m = i / SBK; } else { m = i % SBM; k = i / SBM; }
      const int gm = m0 + m, gk = k0 + k;
      float v = 0.f;
      if (gm < M && gk < K) v = to_f32<TA>(a_k ? A[(int64_t)gm * lda + gk] : A[(int64_t)gk * lda + gm]);
      As[k][m] = v;
    }
#pragma unroll
    for (int it = 0; it < (SBN * SBK) / 256; ++it) {
      const int i = tid + it * 256;
      int n, k;
      if (b_k) { k = i % SBK; n = i / SBK; } else { n = i % SBN; k = i / SBN; }
      const int gn = n0 + n, gk = k0 + k;
      float v = 0.f;
      if (gn < N && gk < K) v = to_f32<TB>(b_k ? B[(int64_t)gn * ldb + gk] : B[(int64_t)gk * ldb + gn]);
      Bs[k][n] = v;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < SBK; ++k) {
      const float4 a4 = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[k][tx * 8]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[k][tx * 8 + 4]);
      const float a[4] = {a4.x, a4.y, a4.z, a4.w};
      const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) epi_row_chunk<8>(e, m0 + ty * 4 + i, n0 + tx * 8, acc[i]);
}

// One element of C through the shared scalar epilogue.
__device__ __forceinline__ void epi_store_one(const EpiArgs& e, int row, int col, float acc) {
  const float r = epi_scalar(e, row, col, acc);
  if (e.colsum) atomicAdd(e.colsum + col, r);
  const int64_t ci = (int64_t)row * e.ldc + col;
  if (e.c_dtype == CLG_BF16) {
    bf16* c = (bf16*)e.C;
    c[ci] = __float2bfloat16_rn(e.accumulate ? __bfloat162float(c[ci]) + r : r);
  } else {
    float* c = (float*)e.C;
    if (e.accumulate == 2) atomicAdd(c + ci, r);
    else if (e.accumulate == 1) c[ci] += r;
    else c[ci] = r;
  }
}

// K <= 8, both operands K-major (the box-coordinate projection Linear(5 -> 768)): a thread keeps its output column's
// weights in registers and walks the CTA's rows; stores are coalesced along N.  Output-bandwidth bound.
static constexpr int SK_MAX = 8, SK_ROWS = 16;
template <typename TA, typename TB>
__global__ void __launch_bounds__(256)
smallk_fwd_kernel(const TA* __restrict__ A, const TB* __restrict__ B, int M, int N, int K, int64_t lda, int64_t ldb,
                  const EpiArgs e) {
  __shared__ float As[SK_ROWS][SK_MAX];
  const int r0 = blockIdx.x * SK_ROWS;
  if (threadIdx.x < SK_ROWS * SK_MAX) {
    const int r = threadIdx.x / SK_MAX, k = threadIdx.x % SK_MAX;
    As[r][k] = (r0 + r < M && k < K) ? to_f32<TA>(A[(int64_t)(r0 + r) * lda + k]) : 0.f;
  }
  __syncthreads();
  const int rows = min(SK_ROWS, M - r0);
  // plain form (bias only, fp32 rows of whole float4s): four columns per thread, 16 B stores
  const bool plain = e.act == CLG_ACT_NONE && !e.drop.on && !e.residual && !e.mask_bits && !e.colsum && !e.accumulate &&
                     e.alpha == 1.f && e.c_dtype == CLG_F32 && (N & 3) == 0 && (e.ldc & 3) == 0 &&
                     (reinterpret_cast<uintptr_t>(e.C) & 15) == 0;
  if (plain) {
    for (int n = threadIdx.x * 4; n < N; n += 1024) {
      float w[4][SK_MAX], bv[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        bv[j] = e.bias ? __ldg(e.bias + n + j) : 0.f;
#pragma unroll
        for (int k = 0; k < SK_MAX; ++k) w[j][k] = k < K ? to_f32<TB>(B[(int64_t)(n + j) * ldb + k]) : 0.f;
      }
      for (int r = 0; r < rows; ++r) {
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int k = 0; k < SK_MAX; ++k) {
          if (k < K) {
            const float a = As[r][k];
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[j] = fmaf(a, w[j][k], acc[j]);
          }
        }
        *reinterpret_cast<float4*>((float*)e.C + (int64_t)(r0 + r) * e.ldc + n) =
            make_float4(acc[0] + bv[0], acc[1] + bv[1], acc[2] + bv[2], acc[3] + bv[3]);
      }
    }
    return;
  }
  for (int n = threadIdx.x; n < N; n += 256) {
    float w[SK_MAX];
#pragma unroll
    for (int k = 0; k < SK_MAX; ++k) w[k] = k < K ? to_f32<TB>(B[(int64_t)n * ldb + k]) : 0.f;
    for (int r = 0; r < rows; ++r) {
      float acc = 0.f;
#pragma unroll
      for (int k = 0; k < SK_MAX; ++k)
        if (k < K) acc = fmaf(As[r][k], w[k], acc);
      epi_store_one(e, r0 + r, n, acc);
    }
  }
}

// N <= 8, both operands MN-major, long K (the weight gradient of that projection, K = region rows): a thread owns one
// output row m (coalesced reads of A along m), a CTA a slice of K; fp32 atomics combine the slices.  Reads A once.
template <typename TA, typename TB>
__global__ void __launch_bounds__(256)
smalln_wgrad_kernel(const TA* __restrict__ A, const TB* __restrict__ B, int M, int N, int K, int64_t lda, int64_t ldb,
                    int k_per_cta, const EpiArgs e) {
  __shared__ float Bs[64][SK_MAX];
  const int m = blockIdx.x * 256 + threadIdx.x;
  const int k_lo = blockIdx.y * k_per_cta, k_hi = min(K, k_lo + k_per_cta);
  float acc[SK_MAX];
#pragma unroll
  for (int n = 0; n < SK_MAX; ++n) acc[n] = 0.f;
  for (int k0 = k_lo; k0 < k_hi; k0 += 64) {
    const int kc = min(64, k_hi - k0);
    __syncthreads();
    for (int i = threadIdx.x; i < 64 * SK_MAX; i += 256) {
      const int k = i / SK_MAX, n = i % SK_MAX;
      Bs[k][n] = (k < kc && n < N) ? to_f32<TB>(B[(int64_t)(k0 + k) * ldb + n]) : 0.f;
    }
    __syncthreads();
    if (m < M) {
#pragma unroll 8
      for (int k = 0; k < kc; ++k) {
        const float a = to_f32<TA>(A[(int64_t)(k0 + k) * lda + m]);
#pragma unroll
        for (int n = 0; n < SK_MAX; ++n) acc[n] = fmaf(a, Bs[k][n], acc[n]);
      }
    }
  }
  if (m < M) {
#pragma unroll
    for (int n = 0; n < SK_MAX; ++n)
      if (n < N) epi_store_one(e, m, n, acc[n]);
  }
}

template <typename TA, typename TB>
static int launch_simt(const clgGemmParams& p, cudaStream_t s) {
  EpiArgs e = make_epi(p);
  const bool splittable = p.act == CLG_ACT_NONE && !p.bias && !p.residual && p.c_dtype == CLG_F32 && p.dropout.p <= 0.f && !p.colsum;
  if (p.K <= SK_MAX && p.a_kmajor && p.b_kmajor) {
    smallk_fwd_kernel<TA, TB><<<(p.M + SK_ROWS - 1) / SK_ROWS, 256, 0, s>>>((const TA*)p.A, (const TB*)p.B, p.M, p.N, p.K,
                                                                             p.lda, p.ldb, e);
    return check_launch("smallk_fwd_kernel");
  }
  if (p.N <= SK_MAX && !p.a_kmajor && !p.b_kmajor && splittable && p.K >= 1024) {
    const int mblocks = (p.M + 255) / 256;
    int slices = (4 * num_sms() + mblocks - 1) / mblocks;
    if (slices > p.K / 32) slices = p.K / 32;
    const int kper = (p.K + slices - 1) / slices;
    slices = (p.K + kper - 1) / kper;
    if (!p.accumulate)
      CLG_CUDA_OK(cudaMemset2DAsync(p.C, p.ldc * sizeof(float), 0, (size_t)p.N * sizeof(float), p.M, s), "memset(small-N wgrad)");
    e.accumulate = 2;
    smalln_wgrad_kernel<TA, TB><<<dim3(mblocks, slices), 256, 0, s>>>((const TA*)p.A, (const TB*)p.B, p.M, p.N, p.K, p.lda,
                                                                      p.ldb, kper, e);
    return check_launch("smalln_wgrad_kernel");
  }
  dim3 grid((p.N + SBN - 1) / SBN, (p.M + SBM - 1) / SBM);
  // split-K (atomic fp32 accumulate) for long reductions with too few output tiles (e.g. dW_loc [768 x 7])
  int splits = 1;
  const int tiles = grid.x * grid.y;
  const bool can_split = p.act == CLG_ACT_NONE && !p.bias && !p.residual && p.c_dtype == CLG_F32 && p.dropout.p <= 0.f;
  if (can_split && tiles * 2 <= num_sms() && p.K >= 1024) {
    splits = (2 * num_sms()) / tiles;
    const int max_s = p.K / 256;
    if (splits > max_s) splits = max_s;
    if (splits < 1) splits = 1;
  }
  int kper = (p.K + splits - 1) / splits;
  kper = (kper + SBK - 1) / SBK * SBK;
  splits = (p.K + kper - 1) / kper;
  if (splits > 1) {
    if (!p.accumulate)
      CLG_CUDA_OK(cudaMemset2DAsync(p.C, p.ldc * sizeof(float), 0, (size_t)p.N * sizeof(float), p.M, s), "memset(simt split-K)");
    e.accumulate = 2;
  }
  grid.z = splits;
  gemm_simt_kernel<TA, TB><<<grid, 256, 0, s>>>((const TA*)p.A, (const TB*)p.B, p.M, p.N, p.K, p.lda, p.ldb,
                                                p.a_kmajor, p.b_kmajor, kper, e);
  return check_launch("gemm_simt_kernel");
}

int gemm_simt(const clgGemmParams& p, cudaStream_t s) {
  if (p.M <= 0 || p.N <= 0 || p.K <= 0) return CLG_ERR_ARG;
  if (p.a_dtype == CLG_F32 && p.b_dtype == CLG_F32) return launch_simt<float, float>(p, s);
  if (p.a_dtype == CLG_BF16 && p.b_dtype == CLG_BF16) return launch_simt<bf16, bf16>(p, s);
  if (p.a_dtype == CLG_BF16 && p.b_dtype == CLG_F32) return launch_simt<bf16, float>(p, s);
  return launch_simt<float, bf16>(p, s);
}

}  // namespace clg
