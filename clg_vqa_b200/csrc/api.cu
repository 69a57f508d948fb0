// Status plumbing, device checks, struct-size probes and the GEMM dispatcher of the C ABI.
#include <mutex>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "common.cuh"

namespace clg {

static thread_local char g_last_err[256] = "";

void set_cuda_error(cudaError_t e, const char* where) {
  snprintf(g_last_err, sizeof(g_last_err), "%s: %s", where, cudaGetErrorString(e));
}

int check_launch(const char* where) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_cuda_error(e, where); return CLG_ERR_CUDA; }
  return CLG_OK;
}

struct DevInfo { int sms = 0; int major = 0; bool ok = false; };
static DevInfo g_dev[64];
static std::mutex g_dev_mu;

static const DevInfo& dev_info() {
  int d = 0;
  cudaGetDevice(&d);
  if (d < 0 || d >= 64) d = 0;
  DevInfo& di = g_dev[d];
  if (!di.ok) {
    std::lock_guard<std::mutex> lk(g_dev_mu);
    if (!di.ok) {
      cudaDeviceGetAttribute(&di.sms, cudaDevAttrMultiProcessorCount, d);
      cudaDeviceGetAttribute(&di.major, cudaDevAttrComputeCapabilityMajor, d);
      di.ok = true;
    }
  }
  return di;
}

int num_sms() { const int s = dev_info().sms; return s > 0 ? s : 148; }
bool pdl_enabled() {
  static const bool v = [] { const char* s = getenv("CLG_PDL"); return !(s && s[0] == '0'); }();
  return v;
}
int require_sm100() { return dev_info().major == 10 ? CLG_OK : CLG_ERR_ARCH; }

int gemm_tc(const clgGemmParams& p, cudaStream_t stream);
int gemm_simt(const clgGemmParams& p, cudaStream_t stream);

}  // namespace clg

extern "C" {

const char* clg_strerror(int status) {
  switch (status) {
    case CLG_OK: return "ok";
    case CLG_ERR_ARG: return "invalid argument or unsupported shape";
    case CLG_ERR_ARCH: return "device is not sm_100 (B200); there is no fallback path";
    case CLG_ERR_CUDA: return "CUDA error (see clg_last_cuda_error)";
    case CLG_ERR_UNSUPPORTED: return "combination not implemented by the selected backend";
    case CLG_ERR_WORKSPACE: return "workspace too small";
    default: return "unknown status";
  }
}

const char* clg_last_cuda_error(void) { return clg::g_last_err; }

int clg_version(void) { return 100; }

int clg_device_check(int device) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || device < 0 || device >= n) {
    cudaGetLastError();
    return CLG_ERR_CUDA;
  }
  int major = 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device) != cudaSuccess) return CLG_ERR_CUDA;
  return major == 10 ? CLG_OK : CLG_ERR_ARCH;
}

int clg_num_sms(int device) {
  int sms = 0;
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) return CLG_ERR_CUDA;
  return sms;
}

int clg_gemm(const clgGemmParams* p, clgStream_t stream) {
  if (!p || !p->A || !p->B || !p->C) return CLG_ERR_ARG;
  int st = clg::require_sm100();
  if (st != CLG_OK) return st;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (p->M2 > 0) {
    // two row groups: one tensor-core launch when the tiling allows it, else the two GEMMs one after the other
    if (!p->A2 || !p->B2 || !p->C2) return CLG_ERR_ARG;
    if (p->backend != CLG_GEMM_SIMT && p->a_dtype == CLG_BF16 && p->b_dtype == CLG_BF16 && p->K >= 64) {
      st = clg::gemm_tc(*p, s);
      if (st != CLG_ERR_UNSUPPORTED) return st;
    }
    clgGemmParams g1 = *p, g2 = *p;
    g1.M2 = 0;
    g2.M2 = 0;
    g2.A = p->A2; g2.B = p->B2; g2.C = p->C2;
    g2.lda = p->lda2; g2.ldb = p->ldb2; g2.ldc = p->ldc2;
    g2.mask_bits = p->mask_bits2; g2.M = p->M2;
    st = clg_gemm(&g1, stream);
    return st != CLG_OK ? st : clg_gemm(&g2, stream);
  }
  if (p->backend == CLG_GEMM_SIMT) return clg::gemm_simt(*p, s);
  if (p->backend == CLG_GEMM_TC) return clg::gemm_tc(*p, s);
  // auto: tensor cores for bf16 x bf16 problems that are big enough to tile, FFMA otherwise
  const bool tc_ok = p->a_dtype == CLG_BF16 && p->b_dtype == CLG_BF16 && p->M >= 128 && p->N >= 64 && p->K >= 64;
  if (tc_ok) {
    st = clg::gemm_tc(*p, s);
    if (st != CLG_ERR_UNSUPPORTED) return st;
  }
  return clg::gemm_simt(*p, s);
}

size_t clg_sizeof_gemm_params(void) { return sizeof(clgGemmParams); }
size_t clg_sizeof_layernorm_fwd_params(void) { return sizeof(clgLayerNormFwdParams); }
size_t clg_sizeof_layernorm_bwd_params(void) { return sizeof(clgLayerNormBwdParams); }
size_t clg_sizeof_attention_fwd_params(void) { return sizeof(clgAttentionFwdParams); }
size_t clg_sizeof_attention_bwd_params(void) { return sizeof(clgAttentionBwdParams); }
size_t clg_sizeof_embed_text_params(void) { return sizeof(clgEmbedTextParams); }
size_t clg_sizeof_embed_text_bwd_params(void) { return sizeof(clgEmbedTextBwdParams); }
size_t clg_sizeof_gqa_loss_params(void) { return sizeof(clgGqaLossParams); }
size_t clg_sizeof_adamw_params(void) { return sizeof(clgAdamWParams); }
size_t clg_sizeof_tensor_meta(void) { return sizeof(clgTensorMeta); }
size_t clg_sizeof_plan_entry(void) { return sizeof(clgPlanEntry); }
size_t clg_sizeof_colsum_params(void) { return sizeof(clgColsumParams); }

int clg_run_plan(const clgPlanEntry* entries, int32_t n, void* blob, int32_t rows, uint64_t seed, clgStream_t stream) {
  if (n < 0 || (n > 0 && (!entries || !blob))) return CLG_ERR_ARG;
  uint8_t* base = static_cast<uint8_t*>(blob);
  for (int32_t i = 0; i < n; ++i) {
    const clgPlanEntry& e = entries[i];
    uint8_t* p = base + e.off;
    for (int k = 0; k < 2; ++k) {
      if (e.rows_off[k] >= 0) memcpy(p + e.rows_off[k], &rows, sizeof(int32_t));
      if (e.seed_off[k] >= 0) memcpy(p + e.seed_off[k], &seed, sizeof(uint64_t));
    }
    int st;
    switch (e.op) {
      case CLG_OP_GEMM: st = clg_gemm(reinterpret_cast<const clgGemmParams*>(p), stream); break;
      case CLG_OP_LN_FWD: st = clg_layernorm_fwd(reinterpret_cast<const clgLayerNormFwdParams*>(p), stream); break;
      case CLG_OP_LN_BWD: st = clg_layernorm_bwd(reinterpret_cast<const clgLayerNormBwdParams*>(p), stream); break;
      case CLG_OP_ATTN_FWD: st = clg_attention_fwd(reinterpret_cast<const clgAttentionFwdParams*>(p), stream); break;
      case CLG_OP_ATTN_BWD: st = clg_attention_bwd(reinterpret_cast<const clgAttentionBwdParams*>(p), stream); break;
      case CLG_OP_COLSUM: {
        const clgColsumParams* c = reinterpret_cast<const clgColsumParams*>(p);
        st = clg_colsum(c->x, c->dtype, c->ld, c->M, c->N, c->out, stream);
        break;
      }
      default: return CLG_ERR_ARG;
    }
    if (st != CLG_OK) return st;
  }
  return CLG_OK;
}

}  // extern "C"
