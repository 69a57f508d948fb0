// Text-embedding gather (word + position + token-type rows, fp32 tables) and its scatter-add
// backward.  One warp per token; rows are 768 floats = 3 KB, read / written as float4 (coalesced).
// Position ids follow create_position_ids_from_input_ids (embeddings.py:L157-170):
//   pos = cumsum(ids != pad)[t] * (ids[t] != pad) + pad       (computed per sample with a warp scan)
#include "common.cuh"

namespace clg {

static constexpr int EMB_WARPS = 8;

__device__ __forceinline__ int text_pos_id(const int64_t* ids_row, int t, int T, int pad_id, int lane) {
  // inclusive count of non-pad tokens in [0, t]; T <= 64 handled with two ballots
  int cnt = 0;
  for (int base = 0; base <= t; base += 32) {
    const int idx = base + lane;
    const bool nz = (idx < T) && (idx <= t) && (ids_row[idx] != pad_id);
    cnt += __popc(__ballot_sync(0xffffffffu, nz));
  }
  const bool self = ids_row[t] != pad_id;
  return (self ? cnt : 0) + pad_id;
}

__global__ void __launch_bounds__(EMB_WARPS * 32)
embed_text_fwd_kernel(const clgEmbedTextParams p) {
  const int lane = threadIdx.x & 31;
  const int64_t tok = (int64_t)blockIdx.x * EMB_WARPS + (threadIdx.x >> 5);
  if (tok >= (int64_t)p.B * p.T) return;
  const int b = (int)(tok / p.T), t = (int)(tok % p.T);
  const int64_t id = p.ids[tok];
  int pos;
  if (p.pos_mode == 0) pos = text_pos_id(p.ids + (int64_t)b * p.T, t, p.T, p.pad_id, lane);
  else pos = p.pos_base + t;
  const float* w = p.word + id * p.N;
  const float* pe = p.pos ? p.pos + (int64_t)pos * p.N : nullptr;
  const float* te = (p.type && p.seg) ? p.type + p.seg[tok] * p.N : nullptr;
  float* o = p.out + map_row(p.out_map, tok) * p.ld_out;
  for (int c = lane * 4; c < p.N; c += 128) {
    float4 v = *reinterpret_cast<const float4*>(w + c);
    if (pe) { const float4 a = *reinterpret_cast<const float4*>(pe + c); v.x += a.x; v.y += a.y; v.z += a.z; v.w += a.w; }
    if (te) { const float4 a = *reinterpret_cast<const float4*>(te + c); v.x += a.x; v.y += a.y; v.z += a.z; v.w += a.w; }
    *reinterpret_cast<float4*>(o + c) = v;
  }
}

__global__ void __launch_bounds__(EMB_WARPS * 32)
embed_text_bwd_kernel(const clgEmbedTextBwdParams p) {
  const int lane = threadIdx.x & 31;
  const int64_t tok = (int64_t)blockIdx.x * EMB_WARPS + (threadIdx.x >> 5);
  if (tok >= (int64_t)p.B * p.T) return;
  const int b = (int)(tok / p.T), t = (int)(tok % p.T);
  const int64_t id = p.ids[tok];
  int pos;
  if (p.pos_mode == 0) pos = text_pos_id(p.ids + (int64_t)b * p.T, t, p.T, p.pad_id, lane);
  else pos = p.pos_base + t;
  const float* g = p.dout + map_row(p.dout_map, tok) * p.ld_dout;
  float* dw = (p.dword && id != p.pad_id) ? p.dword + id * p.N : nullptr;   // padding_idx row gets no gradient
  float* dp = p.dpos ? p.dpos + (int64_t)pos * p.N : nullptr;
  const int64_t sg = (p.dtype_tab && p.seg) ? p.seg[tok] : 0;
  float* dt = (p.dtype_tab && p.seg) ? p.dtype_tab + sg * p.N : nullptr;
  float* dt0 = nullptr;                          // row 0 correction when it was pre-loaded with the all-token total
  if (p.type_row0_has_total && dt) {
    if (sg == 0) dt = nullptr; else dt0 = p.dtype_tab;
  }
  for (int c = lane; c < p.N; c += 32) {
    const float v = g[c];
    if (dw) atomicAdd(dw + c, v);
    if (dp) atomicAdd(dp + c, v);
    if (dt) atomicAdd(dt + c, v);
    if (dt0) atomicAdd(dt0 + c, -v);
  }
}

// N = 32 * NR (768): a warp walks EMB_BPW samples at one token index t, keeps the row in registers and
//  - skips rows that are entirely zero (adding them changes nothing; the masked text positions - 40 % of a GQA batch -
//    all carry the padding position id and used to queue thousands of atomic adds of 0.0 on that one row);
//  - accumulates the position-table contribution in registers while the position id stays the same (at a given t it is
//    the same for every unpadded sample) and flushes it with one atomic per column: 8x fewer atomics on the ~40 hot rows.
static constexpr int EMB_BPW = 8;
template <int NR>
__global__ void __launch_bounds__(EMB_WARPS * 32)
embed_text_bwd_rows_kernel(const clgEmbedTextBwdParams p) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int N = 32 * NR;
  const int t = blockIdx.x % p.T;
  const int b0 = ((int)(blockIdx.x / p.T) * EMB_WARPS + warp) * EMB_BPW;
  float acc[NR];
#pragma unroll
  for (int i = 0; i < NR; ++i) acc[i] = 0.f;
  int acc_pos = -1;
  auto flush = [&]() {
    if (acc_pos >= 0 && p.dpos) {
      float* dp = p.dpos + (int64_t)acc_pos * N;
#pragma unroll
      for (int i = 0; i < NR; ++i) { atomicAdd(dp + lane + 32 * i, acc[i]); acc[i] = 0.f; }
    }
    acc_pos = -1;
  };
  for (int b = b0; b < min(p.B, b0 + EMB_BPW); ++b) {
    const int64_t tok = (int64_t)b * p.T + t;
    const float* g = p.dout + map_row(p.dout_map, tok) * p.ld_dout;
    float v[NR];
    bool nz = false;
#pragma unroll
    for (int i = 0; i < NR; ++i) { v[i] = g[lane + 32 * i]; nz |= v[i] != 0.f; }
    if (!__any_sync(0xffffffffu, nz)) continue;
    const int64_t id = p.ids[tok];
    int pos;
    if (p.pos_mode == 0) pos = text_pos_id(p.ids + (int64_t)b * p.T, t, p.T, p.pad_id, lane);
    else pos = p.pos_base + t;
    if (pos != acc_pos) { flush(); acc_pos = pos; }
#pragma unroll
    for (int i = 0; i < NR; ++i) acc[i] += v[i];
    if (p.dword && id != p.pad_id) {                 // padding_idx row gets no gradient
      float* dw = p.dword + id * N;
#pragma unroll
      for (int i = 0; i < NR; ++i) atomicAdd(dw + lane + 32 * i, v[i]);
    }
    if (p.dtype_tab && p.seg) {
      const int64_t sg = p.seg[tok];
      if (!p.type_row0_has_total) {
        float* dt = p.dtype_tab + sg * N;
#pragma unroll
        for (int i = 0; i < NR; ++i) atomicAdd(dt + lane + 32 * i, v[i]);
      } else if (sg != 0) {                          // row 0 was pre-loaded with the all-token total: move this row out
        float* dt = p.dtype_tab + sg * N;
#pragma unroll
        for (int i = 0; i < NR; ++i) { atomicAdd(dt + lane + 32 * i, v[i]); atomicAdd(p.dtype_tab + lane + 32 * i, -v[i]); }
      }
    }
  }
  flush();
}

// Deterministic row reduction for the data-parallel word-embedding gradient.  Input: the all-gathered (token id,
// gradient row) list sorted by id (stable).  One CTA per list position; the CTA at the head of a run of equal ids sums
// that run's rows in list order (the warps split the columns) and adds the result to the table row.  Every rank runs it
// on the same list, so the replicas' gradients - and parameters - stay bit-identical.  No host-side unique / sync.
// mode 0: table[id, :] += run sum.   mode 1 (local pre-reduction before the all-gather): the run sum replaces the run's
// FIRST row in `rows` and the other members get ids_out = pad_id, so that after the all-gather an id occurs at most once
// per rank and no run is longer than the world size.
// The 8 warps of the CTA split a run into 8 contiguous pieces (fixed by the run length only), each warp sums its piece in
// list order over all columns, and the pieces are combined in warp order: deterministic, and 8x shorter chains for the
// long runs (<s>, </s> occur once per sample).
__global__ void __launch_bounds__(256)
segment_sum_rows_kernel(float* __restrict__ rows, int64_t ld_rows, const int64_t* __restrict__ order,
                        const int64_t* __restrict__ sorted_ids, int n, float* __restrict__ table, int N, int64_t pad_id, int mode,
                        int64_t* __restrict__ ids_out) {
  extern __shared__ float part[];                 // [8][N]
  const int k0 = blockIdx.x;
  const int64_t id = sorted_ids[k0];
  if (id == pad_id || (k0 > 0 && sorted_ids[k0 - 1] == id)) return;      // not the head of a run
  int k1 = k0 + 1;
  while (k1 < n && sorted_ids[k1] == id) ++k1;
  const int len = k1 - k0;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (len == 1 && mode == 1) return;              // a single row is already its own sum
  const int per = (len + 7) / 8;
  const int a = k0 + warp * per, b = min(k1, a + per);
  for (int c = lane; c < N; c += 32) {
    float acc = 0.f;
    for (int k = a; k < b; ++k) acc += rows[order[k] * ld_rows + c];
    part[warp * N + c] = acc;
  }
  __syncthreads();
  float* dst = mode == 0 ? table + id * N : rows + order[k0] * ld_rows;
  for (int c = threadIdx.x; c < N; c += 256) {
    float acc = part[c];
#pragma unroll
    for (int w = 1; w < 8; ++w) acc += part[w * N + c];
    if (mode == 0) dst[c] += acc; else dst[c] = acc;
  }
  if (mode == 1)
    for (int k = k0 + 1 + threadIdx.x; k < k1; k += 256) ids_out[order[k]] = pad_id;
}

// mode 0 after the local pre-reduction: an id occurs at most once per rank, so runs are at most world-size long and a
// whole CTA per list position is mostly launch overhead (82 k CTAs at 8 ranks).  Here a WARP owns a list position; the
// head of a run sums its members in list order with the row spread over the lanes (float4 per lane and 128-column step).
// Same order on every rank: deterministic and replica-identical like the CTA version.
__global__ void __launch_bounds__(256)
segment_sum_rows_warp_kernel(const float* __restrict__ rows, int64_t ld_rows, const int64_t* __restrict__ order,
                             const int64_t* __restrict__ sorted_ids, int n, float* __restrict__ table, int N, int64_t pad_id) {
  const int k0 = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (k0 >= n) return;
  const int64_t id = sorted_ids[k0];
  if (id == pad_id || (k0 > 0 && sorted_ids[k0 - 1] == id)) return;      // not the head of a run
  int k1 = k0 + 1;
  while (k1 < n && sorted_ids[k1] == id) ++k1;
  float* dst = table + id * N;
  for (int c = lane * 4; c < N; c += 128) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int k = k0; k < k1; ++k) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(rows + order[k] * ld_rows + c));
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    float4 d = *reinterpret_cast<float4*>(dst + c);
    d.x += acc.x; d.y += acc.y; d.z += acc.z; d.w += acc.w;
    *reinterpret_cast<float4*>(dst + c) = d;
  }
}

}  // namespace clg

extern "C" int clg_segment_sum_rows(float* rows, int64_t ld_rows, const int64_t* order, const int64_t* sorted_ids, int32_t n,
                                    float* table, int32_t N, int64_t pad_id, int32_t mode, int64_t* ids_out, clgStream_t stream) {
  using namespace clg;
  if (!rows || !order || !sorted_ids || n < 0 || N <= 0 || N > 4096) return CLG_ERR_ARG;
  if ((mode == 0 && !table) || (mode == 1 && !ids_out) || (mode != 0 && mode != 1)) return CLG_ERR_ARG;
  int st = require_sm100();
  if (st != CLG_OK) return st;
  if (n == 0) return CLG_OK;
  // mode 0 with 16 B-granular rows: warp per list position (runs are short after the local pre-reduction; a long run is
  // still summed correctly, only serially).  The summation order differs from the CTA kernel's 8-way split, so the
  // choice depends on the arguments alone - never on the data - and every rank makes the same one.
  if (mode == 0 && (N % 4) == 0 && (ld_rows % 4) == 0 && ((reinterpret_cast<uintptr_t>(rows) | reinterpret_cast<uintptr_t>(table)) & 15) == 0) {
    segment_sum_rows_warp_kernel<<<(n + 7) / 8, 256, 0, static_cast<cudaStream_t>(stream)>>>(rows, ld_rows, order, sorted_ids, n, table, N, pad_id);
    return check_launch("segment_sum_rows_warp_kernel");
  }
  segment_sum_rows_kernel<<<n, 256, 8 * N * sizeof(float), static_cast<cudaStream_t>(stream)>>>(rows, ld_rows, order, sorted_ids, n, table,
                                                                                              N, pad_id, mode, ids_out);
  return check_launch("segment_sum_rows_kernel");
}

extern "C" int clg_embed_text_fwd(const clgEmbedTextParams* p, clgStream_t stream) {
  using namespace clg;
  if (!p || !p->ids || !p->word || !p->out || p->B <= 0 || p->T <= 0 || (p->N & 3) || p->T > 64 * 32) return CLG_ERR_ARG;
  int st = require_sm100();
  if (st != CLG_OK) return st;
  const int64_t toks = (int64_t)p->B * p->T;
  embed_text_fwd_kernel<<<(unsigned)ceil_div64(toks, EMB_WARPS), EMB_WARPS * 32, 0, static_cast<cudaStream_t>(stream)>>>(*p);
  return check_launch("embed_text_fwd_kernel");
}

extern "C" int clg_embed_text_bwd(const clgEmbedTextBwdParams* p, clgStream_t stream) {
  using namespace clg;
  if (!p || !p->ids || !p->dout || p->B <= 0 || p->T <= 0) return CLG_ERR_ARG;
  int st = require_sm100();
  if (st != CLG_OK) return st;
  const int64_t toks = (int64_t)p->B * p->T;
  if (p->N == 768) {
    const int bgroups = (p->B + EMB_WARPS * EMB_BPW - 1) / (EMB_WARPS * EMB_BPW);
    embed_text_bwd_rows_kernel<24><<<(unsigned)(bgroups * p->T), EMB_WARPS * 32, 0, static_cast<cudaStream_t>(stream)>>>(*p);
    return check_launch("embed_text_bwd_rows_kernel");
  }
  embed_text_bwd_kernel<<<(unsigned)ceil_div64(toks, EMB_WARPS), EMB_WARPS * 32, 0, static_cast<cudaStream_t>(stream)>>>(*p);
  return check_launch("embed_text_bwd_kernel");
}
