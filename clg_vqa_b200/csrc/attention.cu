// Fused masked self-attention over the short joint text+region sequence (S <= 160, d_head = 64).
// One CTA per (head, sample): the whole K / V of the head lives in shared memory, scores never touch
// HBM.  Numerics follow the oracle's rounding points: q, k, v arrive in the operand dtype (bf16 or
// fp32), scores / softmax are fp32, P is rounded to the operand dtype before P.V, accumulation fp32.
// Backward recomputes P from the saved log-sum-exp (two passes over one S x S fp32 buffer:
// P_drop -> dV, then dS -> dQ, dK).
//
// v1: FFMA inner products out of shared memory (correct for both dtypes; the fp32 mode needs it).
#include "common.cuh"

namespace clg {

static constexpr int DH = 64;
static constexpr int ATT_THREADS = 256;
static constexpr int ATT_WARPS = ATT_THREADS / 32;
static constexpr int MAX_S = 160;
static constexpr int MAX_NJ = MAX_S / 32;

template <typename T> __device__ __forceinline__ float round_operand(float v);
template <> __device__ __forceinline__ float round_operand<float>(float v) { return v; }
template <> __device__ __forceinline__ float round_operand<bf16>(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }

// load a [S, 64] head slice (row stride ld) into smem with row stride `sld` floats
template <typename T>
__device__ __forceinline__ void load_head(float* dst, int sld, const T* src, int64_t ld, int S) {
  for (int i = threadIdx.x; i < S * (DH / 2); i += ATT_THREADS) {
    const int r = i / (DH / 2), c = (i % (DH / 2)) * 2;
    dst[r * sld + c] = to_f32<T>(src[(int64_t)r * ld + c]);
    dst[r * sld + c + 1] = to_f32<T>(src[(int64_t)r * ld + c + 1]);
  }
}

template <typename T>
__global__ void __launch_bounds__(ATT_THREADS)
attn_fwd_kernel(const clgAttentionFwdParams p, const Drop drop) {
  extern __shared__ float sm[];
  const int S = p.S, H = p.H;
  const int h = blockIdx.x, b = blockIdx.y;
  const int SP = S | 1;                       // odd row stride for the transposed K
  float* Kt = sm;                             // [64][SP]
  float* Vs = Kt + DH * SP;                   // [S][64]
  float* qs = Vs + S * DH;                    // [warps][64]
  float* ps = qs + ATT_WARPS * DH;            // [warps][SP]
  float* kb = ps + ATT_WARPS * SP;            // [S] key bias
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int Hd = H * DH;
  const T* base = (const T*)p.qkv + (int64_t)b * S * p.ld_qkv + h * DH;
  const T* Kg = base + Hd;
  const T* Vg = base + 2 * Hd;
  for (int i = threadIdx.x; i < S * DH; i += ATT_THREADS) {
    const int j = i / DH, d = i % DH;
    Kt[d * SP + j] = to_f32<T>(Kg[(int64_t)j * p.ld_qkv + d]);
    Vs[j * DH + d] = to_f32<T>(Vg[(int64_t)j * p.ld_qkv + d]);
  }
  for (int j = threadIdx.x; j < S; j += ATT_THREADS) kb[j] = p.key_bias[(int64_t)b * S + j];
  __syncthreads();

  const int NJ = (S + 31) / 32;
  for (int i = warp; i < S; i += ATT_WARPS) {
    const T* qg = base + (int64_t)i * p.ld_qkv;
    qs[warp * DH + lane] = to_f32<T>(qg[lane]);
    qs[warp * DH + lane + 32] = to_f32<T>(qg[lane + 32]);
    __syncwarp();
    float acc[MAX_NJ];
#pragma unroll
    for (int jj = 0; jj < MAX_NJ; ++jj) acc[jj] = 0.f;
#pragma unroll 8
    for (int d = 0; d < DH; ++d) {
      const float qd = qs[warp * DH + d];
#pragma unroll
      for (int jj = 0; jj < MAX_NJ; ++jj) {
        if (jj < NJ) {
          const int j = lane + jj * 32;
          acc[jj] = fmaf(qd, Kt[d * SP + (j < S ? j : 0)], acc[jj]);
        }
      }
    }
    float m = -INFINITY;
#pragma unroll
    for (int jj = 0; jj < MAX_NJ; ++jj) {
      const int j = lane + jj * 32;
      if (jj < NJ && j < S) { acc[jj] = acc[jj] * p.scale + kb[j]; m = fmaxf(m, acc[jj]); }
      else acc[jj] = -INFINITY;
    }
    m = warp_max(m);
    float sum = 0.f;
#pragma unroll
    for (int jj = 0; jj < MAX_NJ; ++jj) {
      acc[jj] = (acc[jj] == -INFINITY) ? 0.f : expf(acc[jj] - m);
      sum += acc[jj];
    }
    sum = warp_sum(sum);
    const float inv = 1.0f / sum;
    if (lane == 0 && p.lse) p.lse[((int64_t)b * H + h) * S + i] = m + logf(sum);
    const uint64_t drow = (((uint64_t)b * H + h) * S + i) * (uint64_t)S;
#pragma unroll
    for (int jj = 0; jj < MAX_NJ; ++jj) {
      const int j = lane + jj * 32;
      if (jj < NJ && j < S) ps[warp * SP + j] = round_operand<T>(drop_apply(drop, drow + j, acc[jj] * inv));
    }
    __syncwarp();
    float o0 = 0.f, o1 = 0.f;
    for (int j = 0; j < S; ++j) {
      const float pj = ps[warp * SP + j];
      o0 = fmaf(pj, Vs[j * DH + lane], o0);
      o1 = fmaf(pj, Vs[j * DH + lane + 32], o1);
    }
    T* og = (T*)p.ctx + ((int64_t)b * S + i) * p.ld_ctx + h * DH;
    og[lane] = from_f32<T>(o0);
    og[lane + 32] = from_f32<T>(o1);
    __syncwarp();
  }
}

template <typename T>
__global__ void __launch_bounds__(ATT_THREADS)
attn_bwd_kernel(const clgAttentionBwdParams p, const Drop drop) {
  extern __shared__ float sm[];
  const int S = p.S, H = p.H;
  const int h = blockIdx.x, b = blockIdx.y;
  const int SP = S | 1;
  constexpr int KP = DH + 1;                  // padded row stride for K / V (lanes index rows)
  float* Qs = sm;                             // [S][64]
  float* dOs = Qs + S * DH;                   // [S][64]
  float* Ks = dOs + S * DH;                   // [S][65]
  float* Vs = Ks + S * KP;                    // [S][65]
  float* Bf = Vs + S * KP;                    // [S][SP]  P_drop, then dS
  float* Dm = Bf + S * SP;                    // [S]  D_i = dO_i . O_i
  float* Ls = Dm + S;                         // [S]  lse
  float* kb = Ls + S;                         // [S]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int Hd = H * DH;
  const T* base = (const T*)p.qkv + (int64_t)b * S * p.ld_qkv + h * DH;
  load_head<T>(Qs, DH, base, p.ld_qkv, S);
  load_head<T>(Ks, KP, base + Hd, p.ld_qkv, S);
  load_head<T>(Vs, KP, base + 2 * Hd, p.ld_qkv, S);
  const T* dOg = (const T*)p.dctx + (int64_t)b * S * p.ld_dctx + h * DH;
  load_head<T>(dOs, DH, dOg, p.ld_dctx, S);
  for (int j = threadIdx.x; j < S; j += ATT_THREADS) {
    kb[j] = p.key_bias[(int64_t)b * S + j];
    Ls[j] = p.lse[((int64_t)b * H + h) * S + j];
  }
  __syncthreads();
  // D_i from the saved context
  const T* Og = (const T*)p.ctx + (int64_t)b * S * p.ld_ctx + h * DH;
  for (int i = warp; i < S; i += ATT_WARPS) {
    float s = dOs[i * DH + lane] * to_f32<T>(Og[(int64_t)i * p.ld_ctx + lane]) +
              dOs[i * DH + lane + 32] * to_f32<T>(Og[(int64_t)i * p.ld_ctx + lane + 32]);
    s = warp_sum(s);
    if (lane == 0) Dm[i] = s;
  }
  __syncthreads();

  const int NJ = (S + 31) / 32;
  const uint64_t dbase = (((uint64_t)b * H + h) * S) * (uint64_t)S;
  // ---- pass 1: Bf = dropout(P) rounded to the operand dtype
  for (int i = warp; i < S; i += ATT_WARPS) {
    float acc[MAX_NJ];
#pragma unroll
    for (int jj = 0; jj < MAX_NJ; ++jj) acc[jj] = 0.f;
#pragma unroll 8
    for (int d = 0; d < DH; ++d) {
      const float qd = Qs[i * DH + d];
#pragma unroll
      for (int jj = 0; jj < MAX_NJ; ++jj)
        if (jj < NJ) { const int j = lane + jj * 32; acc[jj] = fmaf(qd, Ks[(j < S ? j : 0) * KP + d], acc[jj]); }
    }
#pragma unroll
    for (int jj = 0; jj < MAX_NJ; ++jj) {
      const int j = lane + jj * 32;
      if (jj < NJ && j < S) {
        const float pr = expf(acc[jj] * p.scale + kb[j] - Ls[i]);
        Bf[i * SP + j] = round_operand<T>(drop_apply(drop, dbase + (uint64_t)i * S + j, pr));
      }
    }
  }
  __syncthreads();
  // ---- dV[j][d] = sum_i Pd[i][j] * dO[i][d]
  T* dqkv = (T*)p.dqkv + (int64_t)b * S * p.ld_dqkv + h * DH;
  for (int j = warp; j < S; j += ATT_WARPS) {
    float a0 = 0.f, a1 = 0.f;
    for (int i = 0; i < S; ++i) {
      const float pij = Bf[i * SP + j];
      a0 = fmaf(pij, dOs[i * DH + lane], a0);
      a1 = fmaf(pij, dOs[i * DH + lane + 32], a1);
    }
    T* o = dqkv + (int64_t)j * p.ld_dqkv + 2 * Hd;
    o[lane] = from_f32<T>(a0);
    o[lane + 32] = from_f32<T>(a1);
  }
  __syncthreads();
  // ---- pass 2: Bf = dS = P * (dropmask*scale * dP - D_i)
  for (int i = warp; i < S; i += ATT_WARPS) {
    float acc[MAX_NJ], dp[MAX_NJ];
#pragma unroll
    for (int jj = 0; jj < MAX_NJ; ++jj) { acc[jj] = 0.f; dp[jj] = 0.f; }
#pragma unroll 8
    for (int d = 0; d < DH; ++d) {
      const float qd = Qs[i * DH + d];
      const float od = dOs[i * DH + d];
#pragma unroll
      for (int jj = 0; jj < MAX_NJ; ++jj)
        if (jj < NJ) {
          const int j = lane + jj * 32;
          const int jc = j < S ? j : 0;
          acc[jj] = fmaf(qd, Ks[jc * KP + d], acc[jj]);
          dp[jj] = fmaf(od, Vs[jc * KP + d], dp[jj]);
        }
    }
    const float Di = Dm[i];
#pragma unroll
    for (int jj = 0; jj < MAX_NJ; ++jj) {
      const int j = lane + jj * 32;
      if (jj < NJ && j < S) {
        const float pr = expf(acc[jj] * p.scale + kb[j] - Ls[i]);
        const float dpm = drop_apply(drop, dbase + (uint64_t)i * S + j, dp[jj]);
        Bf[i * SP + j] = round_operand<T>(pr * (dpm - Di));
      }
    }
  }
  __syncthreads();
  // ---- dQ[i][d] = scale * sum_j dS[i][j] K[j][d]
  for (int i = warp; i < S; i += ATT_WARPS) {
    float a0 = 0.f, a1 = 0.f;
    for (int j = 0; j < S; ++j) {
      const float ds = Bf[i * SP + j];
      a0 = fmaf(ds, Ks[j * KP + lane], a0);
      a1 = fmaf(ds, Ks[j * KP + lane + 32], a1);
    }
    T* o = dqkv + (int64_t)i * p.ld_dqkv;
    o[lane] = from_f32<T>(a0 * p.scale);
    o[lane + 32] = from_f32<T>(a1 * p.scale);
  }
  // ---- dK[j][d] = scale * sum_i dS[i][j] Q[i][d]
  for (int j = warp; j < S; j += ATT_WARPS) {
    float a0 = 0.f, a1 = 0.f;
    for (int i = 0; i < S; ++i) {
      const float ds = Bf[i * SP + j];
      a0 = fmaf(ds, Qs[i * DH + lane], a0);
      a1 = fmaf(ds, Qs[i * DH + lane + 32], a1);
    }
    T* o = dqkv + (int64_t)j * p.ld_dqkv + Hd;
    o[lane] = from_f32<T>(a0 * p.scale);
    o[lane + 32] = from_f32<T>(a1 * p.scale);
  }
}

static size_t fwd_smem(int S) {
  const int SP = S | 1;
  return sizeof(float) * ((size_t)DH * SP + (size_t)S * DH + ATT_WARPS * DH + (size_t)ATT_WARPS * SP + S);
}
static size_t bwd_smem(int S) {
  const int SP = S | 1;
  return sizeof(float) * ((size_t)2 * S * DH + (size_t)2 * S * (DH + 1) + (size_t)S * SP + 3 * S);
}

template <typename K> static int set_smem(K kernel, size_t bytes, const char* what) {
  if (bytes > 227 * 1024) return CLG_ERR_ARG;
  CLG_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes), what);
  return CLG_OK;
}

int attention_tc_fwd(const clgAttentionFwdParams& p, cudaStream_t stream);
int attention_tc_bwd(const clgAttentionBwdParams& p, cudaStream_t stream);
int attention_pipe_fwd(const clgAttentionFwdParams& p, cudaStream_t stream);   // persistent pipelined kernels (attention_pipe.cu)
int attention_pipe_bwd(const clgAttentionBwdParams& p, cudaStream_t stream);
static int g_attn_backend = 0;   // 0 auto (pipelined tcgen05, then one-CTA-per-head tcgen05, then FFMA), 1 force FFMA, 2 skip the pipelined kernels

}  // namespace clg

extern "C" int clg_set_attention_backend(int backend) { clg::g_attn_backend = backend; return CLG_OK; }

extern "C" int clg_attention_fwd(const clgAttentionFwdParams* p, clgStream_t stream) {
  using namespace clg;
  if (!p || !p->qkv || !p->key_bias || !p->ctx || p->S <= 0 || p->S > MAX_S || p->B <= 0 || p->H <= 0) return CLG_ERR_ARG;
  int st = require_sm100();
  if (st != CLG_OK) return st;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (p->dtype == CLG_BF16 && g_attn_backend != 1) {
    if (g_attn_backend != 2) {
      st = attention_pipe_fwd(*p, s);
      if (st != CLG_ERR_UNSUPPORTED) return st;
    }
    if (p->cu_rows) return CLG_ERR_UNSUPPORTED;     // only the pipelined kernels read the packed layout
    st = attention_tc_fwd(*p, s);
    if (st != CLG_ERR_UNSUPPORTED) return st;
  }
  if (p->cu_rows) return CLG_ERR_UNSUPPORTED;
  const size_t smem = fwd_smem(p->S);
  const Drop d = make_drop(p->dropout);
  dim3 grid(p->H, p->B);
  if (p->dtype == CLG_BF16) {
    st = set_smem(attn_fwd_kernel<bf16>, smem, "attn_fwd smem");
    if (st != CLG_OK) return st;
    attn_fwd_kernel<bf16><<<grid, ATT_THREADS, smem, s>>>(*p, d);
  } else {
    st = set_smem(attn_fwd_kernel<float>, smem, "attn_fwd smem");
    if (st != CLG_OK) return st;
    attn_fwd_kernel<float><<<grid, ATT_THREADS, smem, s>>>(*p, d);
  }
  return check_launch("attn_fwd_kernel");
}

extern "C" int clg_attention_bwd(const clgAttentionBwdParams* p, clgStream_t stream) {
  using namespace clg;
  if (!p || !p->qkv || !p->key_bias || !p->ctx || !p->dctx || !p->lse || !p->dqkv || p->S <= 0 || p->S > MAX_S ||
      p->B <= 0 || p->H <= 0)
    return CLG_ERR_ARG;
  int st = require_sm100();
  if (st != CLG_OK) return st;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (p->dtype == CLG_BF16 && g_attn_backend != 1) {
    if (g_attn_backend != 2) {
      st = attention_pipe_bwd(*p, s);
      if (st != CLG_ERR_UNSUPPORTED) return st;
    }
    if (p->cu_rows) return CLG_ERR_UNSUPPORTED;
    st = attention_tc_bwd(*p, s);
    if (st != CLG_ERR_UNSUPPORTED) return st;
  }
  if (p->cu_rows) return CLG_ERR_UNSUPPORTED;
  const size_t smem = bwd_smem(p->S);
  const Drop d = make_drop(p->dropout);
  dim3 grid(p->H, p->B);
  if (p->dtype == CLG_BF16) {
    st = set_smem(attn_bwd_kernel<bf16>, smem, "attn_bwd smem");
    if (st != CLG_OK) return st;
    attn_bwd_kernel<bf16><<<grid, ATT_THREADS, smem, s>>>(*p, d);
  } else {
    st = set_smem(attn_bwd_kernel<float>, smem, "attn_bwd smem");
    if (st != CLG_OK) return st;
    attn_bwd_kernel<float><<<grid, ATT_THREADS, smem, s>>>(*p, d);
  }
  return check_launch("attn_bwd_kernel");
}
