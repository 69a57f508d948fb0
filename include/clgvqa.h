/*
 * clgvqa.h - C ABI of the B200-native (sm_100a) kernels behind the VOLTA single-stream
 * vision-language encoder step that CLG-VQA fine-tunes (UC2 / M3P, GQA head, label-prior loss,
 * SFT masks, AdamW).
 *
 * The reference (nooralahzadeh/CLG-VQA) is pure Python/PyTorch and has NO FFI of its own; the
 * boundary it exposes is the nn.Module surface of volta/volta/encoders.py.
 * This header is the new layer underneath that surface: every entry point names the reference
 * code whose arithmetic it replaces.  The only callers are the ctypes wrappers in
 * clg_vqa_b200/ops.py (INTEGRATION.md shows the binding).
 *
 * Conventions
 *   - plain pointers and sizes only; all pointers are DEVICE pointers unless stated otherwise
 *   - the caller allocates everything (outputs, saved-for-backward, workspace)
 *   - kernels are enqueued on `stream` and never synchronise or allocate
 *   - return value: CLG_OK (0) or a negative clgStatus; clg_strerror() names it
 *   - sm_100 only: any other device returns CLG_ERR_ARCH (there is no fallback path)
 *   - "lowp" / operand dtype: CLG_BF16 in the bf16 mode, CLG_F32 in the fp32 correctness mode
 */
#ifndef CLGVQA_H_
#define CLGVQA_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* clgStream_t; /* cudaStream_t */

typedef enum {
  CLG_OK = 0,
  CLG_ERR_ARG = -1,         /* bad argument / unsupported shape */
  CLG_ERR_ARCH = -2,        /* device is not sm_100 */
  CLG_ERR_CUDA = -3,        /* a CUDA runtime / driver call failed (see clg_last_cuda_error) */
  CLG_ERR_UNSUPPORTED = -4, /* combination not implemented by the selected backend */
  CLG_ERR_WORKSPACE = -5    /* workspace too small */
} clgStatus;

typedef enum { CLG_F32 = 0, CLG_BF16 = 1 } clgDType;

const char* clg_strerror(int status);
const char* clg_last_cuda_error(void);
int clg_version(void);
/* CLG_OK iff `device` has compute capability 10.x */
int clg_device_check(int device);
int clg_num_sms(int device);

/* ------------------------------------------------------------------------------------------------
 * Counter-based dropout used by every kernel below: element `idx` of dropout site `offset` is
 * kept iff hash(seed, offset + idx) >= p * 2^32; kept values are scaled by 1/(1-p).  Backward
 * kernels regenerate the same mask from (seed, offset); nothing is stored.
 * Replaces nn.Dropout / F.dropout at encoders.py:L315, L411-421, L553-563, L1238 and
 * embeddings.py:L654, L667 (distributionally; torch's Philox stream cannot be matched).
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  float p;          /* 0 => disabled */
  uint32_t _pad;
  uint64_t seed;
  uint64_t offset;
} clgDropout;

/* ------------------------------------------------------------------------------------------------
 * GEMM:  C[M,N] = epilogue( alpha * sum_k A(m,k) * B(n,k) )
 *   a_kmajor: A(m,k) = A[m*lda + k]   else  A(m,k) = A[k*lda + m]
 *   b_kmajor: B(n,k) = B[n*ldb + k]   else  B(n,k) = B[k*ldb + n]
 * so forward linear y = x W^T is (a_kmajor=1, b_kmajor=1, B=W[N,K]); dgrad dx = dy W is
 * (1, 0, B=W[K',N'] read as [k][n]); wgrad dW = dy^T x is (0, 0).
 * Replaces every nn.Linear / F.linear on the path (encoders.py:L229-359, L399-425, L487-502,
 * L542-567, L603-608, L788-815; embeddings.py:L661-662; m3p_transformer.py:L178-227) and their
 * autograd.  Backend CLG_GEMM_TC = TMA + tcgen05.mma (bf16 operands, fp32 accumulate in TMEM);
 * CLG_GEMM_SIMT = fp32 FFMA kernel (fp32 correctness mode, tiny / odd shapes).
 *
 * Epilogue order:  v = alpha*acc ; +bias[n] ; [dropout if drop_before_act] ; act ;
 *                  [dropout otherwise] ; +residual[m,n] ; *mask_bit[m,n] ; (C += v | C = v)
 * ---------------------------------------------------------------------------------------------- */
typedef enum {
  CLG_ACT_NONE = 0,
  CLG_ACT_GELU = 1,   /* erf GeLU (encoders.py:L131-137); aux_out (optional) receives the pre-activation */
  CLG_ACT_RELU = 2,   /* BertTextPooler (encoders.py:L603-608) */
  CLG_ACT_TANH = 3,   /* M3P BertPooler (m3p_transformer.py:L554-560) */
  CLG_ACT_DGELU = 4,  /* v *= gelu'(aux_in)            aux_in = saved pre-activation */
  CLG_ACT_DRELU = 5,  /* v *= (aux_in > 0)             aux_in = saved activation output */
  CLG_ACT_DTANH = 6,  /* v *= 1 - aux_in^2             aux_in = saved activation output */
  CLG_ACT_GELU_GRAD = 7, /* erf GeLU; aux_out receives gelu'(pre-activation) instead of the pre-activation, so the
                            backward GEMM epilogue is a plain multiply (CLG_ACT_MUL_AUX) */
  CLG_ACT_MUL_AUX = 8 /* v *= aux_in */
} clgAct;

typedef enum { CLG_GEMM_AUTO = 0, CLG_GEMM_TC = 1, CLG_GEMM_SIMT = 2 } clgGemmBackend;

typedef struct {
  const void* A;
  const void* B;
  void* C;
  int32_t a_dtype, b_dtype, c_dtype; /* clgDType */
  int32_t M, N, K;
  int64_t lda, ldb, ldc;
  int32_t a_kmajor, b_kmajor;
  float alpha;
  int32_t act;             /* clgAct */
  const float* bias;       /* [N] or NULL */
  const float* residual;   /* fp32 [M, ld_res] or NULL */
  int64_t ld_res;
  const void* aux_in;      /* [M, ld_aux_in] of aux_in_dtype, for the D* activations */
  int64_t ld_aux_in;
  int32_t aux_in_dtype;
  int32_t aux_out_dtype;
  void* aux_out;           /* [M, ld_aux_out] pre-activation output (GELU) or NULL */
  int64_t ld_aux_out;
  const uint32_t* mask_bits; /* bit (m*N + n) of a packed {0,1} mask or NULL (SFT, train_task_sft.py:L432-461) */
  int32_t accumulate;      /* 1: C += v (atomic when split-K) */
  int32_t drop_before_act;
  clgDropout dropout;
  int32_t backend;         /* clgGemmBackend */
  int32_t split_k;         /* 0 = auto (TC backend, act==NONE && no bias only) */
  float* colsum;           /* optional [N]: colsum[n] += sum_m (epilogue result, before rounding to c_dtype); the bias
                              gradient of the layer that produced A's pre-activation, fused into the dgrad GEMM */
  /* Optional second group, row-concatenated under the first (M2 > 0): rows [M, M + M2) of the virtual output are
   * C2 = epi(A2 B2^T) with the same N, K, dtypes, major-ness and epilogue options but their own operands, output and
   * mask.  M must be a multiple of 256.  One launch then fills the machine where two small GEMMs could not: the QKV
   * (2304 x 768) and attention-output (768 x 768) weight gradients of a layer together have the tile count of one FFN
   * weight gradient. */
  const void* A2; const void* B2; void* C2;
  int64_t lda2, ldb2, ldc2;
  const uint32_t* mask_bits2;
  int32_t M2; int32_t _pad2;
} clgGemmParams;

int clg_gemm(const clgGemmParams* p, clgStream_t stream);
/* number of SMs the persistent tcgen05 GEMM grids occupy (0 = all).  Data parallelism leaves a few SMs to NCCL's
 * all-reduce kernels, which otherwise delay individual persistent CTAs and with them every GEMM they overlap. */
int clg_set_gemm_sm_limit(int n);
size_t clg_sizeof_gemm_params(void);

/* ------------------------------------------------------------------------------------------------
 * LayerNorm family.  Replaces BertLayerNorm (apex FusedLayerNorm kernel
 * volta/apex/csrc/layer_norm_cuda_kernel.cu:L279-322 fwd, L403-637 bwd, or the python fallback
 * encoders.py:L49-62) and nn.LayerNorm (m3p_transformer.py:L244, L658, L693, L708).
 *
 * Row remapping lets the embedding LNs write straight into the joint [B, S, N] sequence buffer:
 *   mapped(r) = (r / group_rows) * group_stride + row_offset + (r % group_rows)   (group_rows = 0: identity)
 *
 * forward:  y = LN(x * pre_valid[r]) * gamma + beta ; y *= post_valid[r] ; y += add_mat[amap(r)] ;
 *           y += add_vec ; y = dropout(y) ; store y (fp32, mapped row) and y_lowp (mapped row)
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  int32_t group_rows, row_offset;
  int64_t group_stride;
} clgRowMap;

typedef struct {
  const float* x;  int64_t ldx;      /* [rows, N] pre-LN input */
  const float* gamma; const float* beta;
  float* y;        int64_t ldy;      /* fp32 output (may be NULL) */
  void* y_lowp;    int64_t ldy_lowp; int32_t lowp_dtype; int32_t _pad0; /* GEMM-operand copy or NULL */
  float* mean; float* rstd;          /* [rows] saved statistics (may be NULL in eval) */
  int32_t rows, N;
  float eps; int32_t _pad1;
  clgRowMap out_map;
  const float* add_mat; int64_t ld_add; clgRowMap add_map; /* optional */
  const float* add_vec;              /* [N] optional */
  const uint8_t* row_valid;          /* [rows] optional (M3P prefix mask, m3p_transformer.py:L937, L954) */
  int32_t valid_pre, valid_post;
  clgDropout dropout;                /* applied to the stored outputs */
} clgLayerNormFwdParams;

int clg_layernorm_fwd(const clgLayerNormFwdParams* p, clgStream_t stream);
size_t clg_sizeof_layernorm_fwd_params(void);

/* backward of the above (add_mat / add_vec get dy unchanged; the caller routes it):
 *   g = dropout_out(dy[mapped r]) * post_valid ; dx = LN'(g) * pre_valid
 *   dgamma += sum_r g*xhat ; dbeta += sum_r g                 (atomic accumulate, fp32)
 *   dx (fp32) ; dx_lowp = dropout_in(dx) in lowp dtype ; dbias += sum_r dropout_in(dx)
 * where dropout_in is the dropout site of the GEMM epilogue that produced x (encoders.py:L417, L559). */
typedef struct {
  const float* dy; int64_t lddy; clgRowMap dy_map;
  const float* x;  int64_t ldx;
  const float* gamma;
  const float* mean; const float* rstd;
  float* dx;       int64_t lddx;     /* fp32 (may be NULL) */
  void* dx_lowp;   int64_t lddx_lowp; int32_t lowp_dtype; int32_t _pad0;
  float* dgamma; float* dbeta;       /* [N] accumulated (may be NULL) */
  float* dbias;                      /* [N] accumulated colsum of dropout_in(dx) (may be NULL) */
  int32_t rows, N;
  const uint8_t* row_valid; int32_t valid_pre, valid_post;
  clgDropout dropout_out;            /* site applied to y in forward */
  clgDropout dropout_in;             /* site of the producing GEMM */
} clgLayerNormBwdParams;

int clg_layernorm_bwd(const clgLayerNormBwdParams* p, clgStream_t stream);
size_t clg_sizeof_layernorm_bwd_params(void);

/* ------------------------------------------------------------------------------------------------
 * Fused joint text+region self-attention for the short concatenated sequence (S <= 160, d_head 64).
 * qkv: [B*S, 3*H*64] rows = tokens, columns = [q | k | v], head h at columns h*64 of each third.
 *   P = softmax_fp32( (q k^T) * scale + key_bias[b, :] ) ; P = dropout(P) ; ctx = P v
 * key_bias = (1-mask)*-10000 for UC2 (encoders.py:L978-995, L259-315), 0 / -inf for M3P
 * (m3p_transformer.py:L197-205).  In the bf16 mode q, k, v and P are bf16 operands with fp32
 * accumulation, exactly the rounding points of the bf16 oracle.
 * lse [B, H, S] (log-sum-exp of the scaled, biased scores) is saved for the backward, which
 * recomputes P and returns dqkv in the same layout.
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  const void* qkv; int64_t ld_qkv;
  const float* key_bias;             /* [B, S] additive, fp32 */
  void* ctx; int64_t ld_ctx;         /* [B*S, H*64] */
  float* lse;                        /* [B, H, S] or NULL (eval) */
  int32_t dtype;                     /* clgDType of qkv / ctx */
  int32_t B, S, H;
  float scale; int32_t _pad;
  clgDropout dropout;
  const int32_t* cu_rows;            /* optional [B+1] (device): packed variable-length layout - sample b owns rows
                                        [cu_rows[b], cu_rows[b+1]) of qkv / ctx (at most S of them) and key_bias is indexed by
                                        packed row.  NULL: sample b owns rows [b*S, (b+1)*S) and key_bias is [B, S].
                                        Only the pipelined tcgen05 kernels implement it (CLG_ERR_UNSUPPORTED otherwise). */
  int32_t total_rows; int32_t _pad2; /* with cu_rows: cu_rows[B] (rows beyond it are never read) */
} clgAttentionFwdParams;

typedef struct {
  const void* qkv; int64_t ld_qkv;
  const float* key_bias;
  const void* ctx; int64_t ld_ctx;
  const void* dctx; int64_t ld_dctx;
  const float* lse;
  void* dqkv; int64_t ld_dqkv;
  int32_t dtype;
  int32_t B, S, H;
  float scale; int32_t _pad;
  clgDropout dropout;
  const int32_t* cu_rows;            /* as in clgAttentionFwdParams */
  int32_t total_rows; int32_t _pad2;
} clgAttentionBwdParams;

/* 0 = auto (tcgen05 kernels for bf16, FFMA kernel for fp32 / unsupported shapes), 1 = force the FFMA kernel */
int clg_set_attention_backend(int backend);
int clg_attention_fwd(const clgAttentionFwdParams* p, clgStream_t stream);
int clg_attention_bwd(const clgAttentionBwdParams* p, clgStream_t stream);
size_t clg_sizeof_attention_fwd_params(void);
size_t clg_sizeof_attention_bwd_params(void);

/* ------------------------------------------------------------------------------------------------
 * Text embedding path.  Replaces UC2Embeddings.forward text half (embeddings.py:L640-654) and
 * create_position_ids_from_input_ids (L157-170): out[r] = word[ids[r]] + pos[posid(r)] + type[seg[r]],
 * posid = cumsum(ids != pad) * (ids != pad) + pad per sample (pos_mode 0), or the absolute position
 * pos_base + t (pos_mode 1, M3P m3p_transformer.py:L912-933; type table unused).
 * The LayerNorm that follows is clg_layernorm_fwd.  Backward scatter-adds d(out) into the three
 * tables (atomic), skipping the padding row of the word table (nn.Embedding padding_idx).
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  const int64_t* ids; const int64_t* seg; /* [B, T] */
  const float* word; const float* pos; const float* type; /* tables, row length N */
  float* out; int64_t ld_out; clgRowMap out_map;
  int32_t B, T, N, pad_id;
  int32_t pos_mode, pos_base;
  int64_t vocab;
} clgEmbedTextParams;

typedef struct {
  const int64_t* ids; const int64_t* seg;
  const float* dout; int64_t ld_dout; clgRowMap dout_map;
  float* dword; float* dpos; float* dtype_tab;    /* accumulated; any may be NULL */
  int32_t B, T, N, pad_id;
  int32_t pos_mode, pos_base;
  int64_t vocab;
  int32_t type_row0_has_total;  /* 1: row 0 of dtype_tab already received the column sums over ALL tokens (emitted by the
                                   LayerNorm backward that produced dout); segment-0 tokens then skip their type atomics
                                   (10^4-way contention on 768 addresses) and segment-1 tokens move their share from row 0
                                   to row 1 */
  int32_t _pad;
} clgEmbedTextBwdParams;

/* sorted_ids[n]: token ids sorted ascending (stable), order[n]: the row index of each sorted entry.  For every run of
 * equal ids (!= pad_id) the rows[order[k], :] of the run are summed in list order by one CTA (deterministic: 8 contiguous
 * pieces combined in order).  mode 0: table[id, :] += sum.  mode 1: the sum replaces the run's first row and
 * ids_out[order[k]] = pad_id for the other members (local pre-reduction: afterwards an id occurs once).
 * Replaces the atomic scatter for the data-parallel word-embedding gradient (apex all-reduces the dense 250002 x 768
 * table, distributed.py:L491-510; here every rank reduces the same all-gathered (token id, gradient row) list in the same
 * order, so the replicas stay bit-identical). */
int clg_segment_sum_rows(float* rows, int64_t ld_rows, const int64_t* order, const int64_t* sorted_ids, int32_t n,
                         float* table, int32_t N, int64_t pad_id, int32_t mode, int64_t* ids_out, clgStream_t stream);
int clg_embed_text_fwd(const clgEmbedTextParams* p, clgStream_t stream);
int clg_embed_text_bwd(const clgEmbedTextBwdParams* p, clgStream_t stream);
size_t clg_sizeof_embed_text_params(void);
size_t clg_sizeof_embed_text_bwd_params(void);

/* ------------------------------------------------------------------------------------------------
 * GQA answer head loss.  Replaces task_utils.py:L413-428 (+ L706-711 score, L831-841 argmax):
 *   p = softmax(z) ; K = top-k(p) ; s_b = sum_{j in K} p_j * dist[b,j] ; y_b = argmax trunc(target[b])
 *   loss = A * ( mean_b -log p[b,y_b] + lambda * mean_b s_b ) ;  score = mean_b target[b, argmax z_b]
 *   dlogits[b,i] = gscale * (A/B) * ( p_i - [i=y_b] + lambda * p_i * ( dist_i*[i in K] - s_b ) )
 * One CTA per row; out[0]=loss, out[1]=score are reduced by a second tiny deterministic pass.
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  const float* logits; int64_t ld_logits;
  const float* target; int64_t ld_target;
  const float* dist;   int64_t ld_dist;    /* NULL => CE only (validation, task_utils.py:L265-269) */
  float* dlogits;      int64_t ld_dlogits; /* NULL => forward only */
  float* row_buf;                          /* [B, 4] scratch: nll, s, score, argmax */
  float* out;                              /* [2] loss, score */
  int64_t* pred;                           /* [B] argmax z (may be NULL) */
  int32_t B, A, topk; float lambda;
  float gscale; int32_t score_mean;        /* score_mean=1: /B (train), 0: sum (val) */
} clgGqaLossParams;

int clg_gqa_prior_loss(const clgGqaLossParams* p, clgStream_t stream);
size_t clg_sizeof_gqa_loss_params(void);

/* ------------------------------------------------------------------------------------------------
 * Optimiser step over flat arenas.  Replaces torch clip_grad_norm_ + pytorch_transformers AdamW
 * (train_task.py:L326-338; algorithm SURVEY.md Appendix A.4) and the per-forward weight re-masking
 * of torch.nn.utils.prune (train_task_sft.py:L461).
 *   clg_sumsq      : out[0] += sum g^2 over the flat gradient arena (fp32 partials, deterministic 2 passes)
 *   clg_adamw_step : per 2048-element chunk c with tensor t = chunk_tensor[c]:
 *       g' = g * gscale * min(1, max_norm/(sqrt(sumsq)*gscale + 1e-6))
 *       m,v update ; p -= lr_t*corr * m/(sqrt(v)+eps) ; p -= lr_t*wd_t*p
 *       optional shadow: lowp[t] = bf16(p * mask_bit)
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  float lr, wd;
  int64_t shadow_off;   /* element offset into the bf16 shadow arena, or -1 */
  int64_t mask_off;     /* bit offset (multiple of 32) into the packed mask arena, or -1 */
  int64_t start;        /* element offset of the tensor in the arenas */
  int64_t numel;
} clgTensorMeta;

typedef struct {
  float* p; const float* g; float* m; float* v;
  int64_t n;                         /* arena length, multiple of 2048 */
  const int32_t* chunk_tensor;       /* [n/2048] tensor index, -1 = skip (padding / no grad) */
  const clgTensorMeta* tensors;      /* device array */
  const float* sumsq;                /* device scalar (may be NULL: no clipping) */
  float max_norm, gscale;
  float beta1, beta2, eps;
  float bias_c1, bias_c2;            /* 1-beta1^t, 1-beta2^t (1,1 when correct_bias is off) */
  int32_t zero_grad;                 /* 1: the step also leaves g == 0 everywhere (g is then written, despite the const):
                                        the gradient it has just read is overwritten in place, which saves the separate
                                        fill pass over the arena; blocks that `hot` marks all-zero are not touched */
  void* shadow; const uint32_t* mask_bits;
  const uint8_t* hot;                /* optional [n/256], 8 B aligned (a chunk's 8 bytes are read as one word):
                                        0 = the block's g, m, v are all zero (marked by clg_sumsq on THIS
                                        gradient): only the decoupled decay p -= lr*wd*p is applied - bit-identical to the
                                        dense update, 8 instead of 28 bytes of HBM traffic per parameter (the word-embedding
                                        rows a batch never touches); when lr*wd < 2^-25 that decay rounds to p itself and
                                        the block is not touched at all (its shadow, if any, must already hold bf16(p*mask):
                                        clg_cast_shadow after every out-of-band change of p) */
} clgAdamWParams;

/* out[0] += sum g^2 (two deterministic passes).  hot (optional, [n/256] bytes): hot[b] = 1 for every 256-element block
 * of g that holds a non-zero (never cleared here): blocks that have never seen a gradient have m = v = 0 and are
 * updated by the shortcut below. */
int clg_sumsq(const float* g, int64_t n, float* partials, int32_t num_partials, float* out, uint8_t* hot, clgStream_t stream);
/* as clg_sumsq, but the 256-element blocks [sparse_lo, sparse_hi) of the arena (the word-embedding table's gradient) are
 * read only where touched[b] != 0: every other block of that range must be exactly zero (its rows were not in the batch).
 * clg_mark_rows sets the marks from the token ids of a backward: the blocks overlapped by the gradient row of every id
 * (row `id` starts at element first_elem + id * row_elems; ids outside [0, num_rows) and id == skip_id - the padding
 * row, which receives no gradient - are ignored).  The caller clears `touched` when it clears the gradient. */
int clg_sumsq_sparse(const float* g, int64_t n, float* partials, int32_t num_partials, float* out, uint8_t* hot,
                     int64_t sparse_lo, int64_t sparse_hi, const uint8_t* touched, clgStream_t stream);
int clg_mark_rows(const int64_t* ids, int64_t n, int64_t row_elems, int64_t first_elem, int64_t num_rows, int64_t skip_id,
                  uint8_t* touched, clgStream_t stream);
int clg_adamw_step(const clgAdamWParams* p, clgStream_t stream);
size_t clg_sizeof_adamw_params(void);
size_t clg_sizeof_tensor_meta(void);

/* shadow refresh without an optimiser step: lowp = cast(p * mask) for every tensor with a shadow */
int clg_cast_shadow(const float* p, int64_t n, const int32_t* chunk_tensor, const clgTensorMeta* tensors,
                    void* shadow, const uint32_t* mask_bits, clgStream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Small utilities
 * ---------------------------------------------------------------------------------------------- */
/* dst[i] = bf16(src[i]) (region features, 2048-d fp32 -> bf16 GEMM operand) */
int clg_cast_f32_to_bf16(const float* src, void* dst, int64_t n, clgStream_t stream);
/* dst[r*ld_dst + c] = bf16(src[r*ld_src + c]) for c < cols, 0 for cols <= c < ld_dst (row-padded bf16 GEMM operand:
 * the [B, 1842] logit gradient needs a 16 B-aligned row pitch for TMA) */
int clg_cast_f32_to_bf16_2d(const float* src, int64_t ld_src, void* dst, int64_t ld_dst, int32_t rows, int32_t cols,
                            clgStream_t stream);
/* Packing of the joint [text | regions] sequence: padded positions (mask == 0; UC2 attention_mask / image_attention_mask,
 * encoders.py:L978-995) never influence a kept position (their key bias underflows exp() to exactly 0) and nothing reads
 * their outputs, so the encoder layers only need the kept rows.  keep(b, s) = mask != 0 or s == 0 (token 0 feeds the pooler,
 * encoders.py:L603-608).  cu_rows[b] = first packed row of sample b (cu_rows[B] = total), src_rows[i] = b*(T+R)+s of packed
 * row i.  One CTA; B <= 8192. */
int clg_pack_index(const int64_t* txt_mask, const int64_t* img_mask, int32_t B, int32_t T, int32_t R, int32_t* cu_rows,
                   int32_t* src_rows, clgStream_t stream);
/* dst[i, :] = src[idx[i], :] (gather) or dst[idx[i], :] = src[i, :] (scatter) for i < n; rows of `cols` elements of
 * dtype (CLG_F32 | CLG_BF16); idx_scale multiplies the index (rows of a [B, S*d] view: the first token of every sample) */
int clg_rows_gather(const void* src, int64_t ld_src, const int32_t* idx, int32_t n, void* dst, int64_t ld_dst, int32_t cols,
                    int32_t dtype, clgStream_t stream);
int clg_rows_scatter(const void* src, int64_t ld_src, const int32_t* idx, int32_t n, void* dst, int64_t ld_dst, int32_t cols,
                     int32_t dtype, clgStream_t stream);
/* ------------------------------------------------------------------------------------------------
 * Global magnitude pruning of the IMP loop (volta/train_task_prunning.py:L45-90, L728-806:
 * prune.global_unstructured(L1Unstructured, amount) over the attention / FFN / pooler weights).
 * tensors: device array of clgPruneTensor.  The k-th smallest |w| among the unpruned entries is found by a radix select:
 * four clg_prune_hist passes (8-bit digit at shift 24, 16, 8, 0 over the entries whose higher key bits equal prefix;
 * hist256 is accumulated, zero it first), the host picks the bucket between passes; clg_prune_apply then writes the new
 * {0,1} masks: pruned where |w| < T (key order = value order for non-negative floats), and ties_to_prune of the entries
 * equal to T (tie_counter: device scalar, zero it first).
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  const float* w;        /* weight values (the reference scores |weight_orig * weight_mask|) */
  const float* mask;     /* current fp32 {0,1} mask or NULL (nothing pruned yet) */
  float* mask_out;       /* new mask (may alias mask) */
  int64_t n;
} clgPruneTensor;
int clg_prune_hist(const void* tensors, int32_t n_tensors, uint32_t prefix, uint32_t prefix_mask, int32_t shift,
                   uint64_t* hist256, clgStream_t stream);
int clg_prune_apply(const void* tensors, int32_t n_tensors, uint32_t thr_key, uint64_t ties_to_prune, uint64_t* tie_counter,
                    clgStream_t stream);
size_t clg_sizeof_prune_tensor(void);

/* bits[i/32] bit (i%32) = (mask[i] != 0); n multiple of 32 not required (tail bits are 0) */
int clg_pack_mask(const float* mask, uint32_t* bits, int64_t n, clgStream_t stream);
/* w[i] *= bit(i)  (pre-zeroing of masked weights, train_task_sft.py:L432-453) */
int clg_apply_mask(float* w, const uint32_t* bits, int64_t n, clgStream_t stream);
/* out[n] += sum_m X[m*ld + n]  (bias gradients that are not fused elsewhere) */
int clg_colsum(const void* x, int32_t dtype, int64_t ld, int32_t M, int32_t N, float* out, clgStream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Launch plans: the launch sequence of an encoder layer (BertEncoder.forward, encoders.py:L848-892, is a Python loop that
 * dispatches ~1.5 k ops per step; here a layer is ~8 forward / ~13 backward launches) recorded once as filled parameter
 * structs and replayed natively.  Between two steps only the packed row count and the dropout seed change: entry i patches
 * `rows` into the int32 fields at rows_off[] and `seed` into the uint64 fields at seed_off[] of its struct (offsets relative
 * to the struct, -1 = unused), then calls the entry point of `op`.  `blob` is host memory owned by the caller and is modified
 * in place (one replay at a time per plan).  Stops at the first failing launch and returns its status.
 * ---------------------------------------------------------------------------------------------- */
enum { CLG_OP_GEMM = 0, CLG_OP_LN_FWD = 1, CLG_OP_LN_BWD = 2, CLG_OP_ATTN_FWD = 3, CLG_OP_ATTN_BWD = 4, CLG_OP_COLSUM = 5 };
typedef struct { const void* x; float* out; int64_t ld; int32_t dtype, M, N, _pad; } clgColsumParams;   /* clg_colsum's arguments */
typedef struct {
  int32_t op;            /* CLG_OP_* */
  int32_t off;           /* byte offset of the parameter struct inside blob (16 B aligned) */
  int32_t rows_off[2];
  int32_t seed_off[2];
} clgPlanEntry;
int clg_run_plan(const clgPlanEntry* entries, int32_t n, void* blob, int32_t rows, uint64_t seed, clgStream_t stream);
size_t clg_sizeof_plan_entry(void);
size_t clg_sizeof_colsum_params(void);
/* key_bias[b, s] = (mask[b, s] != 0) ? 0 : fill  from the int64 masks of the module interface.
 * UC2: two masks concatenated [text | regions], fill -10000 (encoders.py:L978-995).
 * M3P: prefix validity over [regions | text], fill -inf, also writes row_valid (m3p_transformer.py:L59-78). */
int clg_build_key_bias(const int64_t* txt_mask, const int64_t* img_mask, int32_t B, int32_t T, int32_t R,
                       int32_t m3p_prefix, float* key_bias, uint8_t* row_valid, clgStream_t stream);
/* elementwise dropout of a small [rows, N] fp32 matrix in place (pooled output, encoders.py:L1238),
 * and its backward (same call on the gradient) */
int clg_dropout_inplace(float* x, int64_t n, const clgDropout* d, clgStream_t stream);

/* out[i] = dy[i] * act'(aux[i]) for the D* activations (classifier GeLU backward, encoders.py:L788-815);
 * all fp32, out may alias dy */
int clg_act_bwd(const float* dy, const float* aux, float* out, int64_t n, int32_t act, clgStream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* CLGVQA_H_ */
