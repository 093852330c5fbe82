// api.cu — C-ABI entry points for the GEMV and the fused building blocks (declared in include/b200_decode.h).
#include "gemv.cuh"
#include "ops.cuh"

namespace b200 {

static int num_sms_cached() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) !=
                                                  cudaSuccess) {
      (void)cudaGetLastError();
      sms = 0;
      return 148;
    }
  }
  return sms;
}

}  // namespace b200

extern "C" {

int b200_gemv_bf16(void* y, const void* x, const void* W, const void* bias, int64_t m, int64_t n, int64_t k,
                   void* stream) {
  using namespace b200;
  B200_CHECK_ARG(y && x && W, "gemv: null pointer");
  B200_CHECK_ARG(m >= 1 && m <= 4096, "gemv: m=%lld out of range (this entry point is the decode GEMV)", (long long)m);
  int rc = b200_device_check();
  if (rc != B200_OK) return rc;
  if ((rc = gemv_setup_attributes()) != B200_OK) return rc;
  GemvPlan plan;
  if ((rc = gemv_make_plan(&plan, W, n, n, k, 1, PRO_PLAIN, EPI_PLAIN, num_sms_cached())) != B200_OK) return rc;
  plan.p.bias = (const __nv_bfloat16*)bias;
  for (int64_t r = 0; r < m; ++r) {  // W is re-streamed per row of x; the batched path is the prefill GEMM
    plan.p.x = (const __nv_bfloat16*)x + r * k;
    plan.p.y = (__nv_bfloat16*)y + r * n;
    if ((rc = gemv_launch(plan, (cudaStream_t)stream, false)) != B200_OK) return rc;
  }
  return B200_OK;
}

int b200_gemv_fused_bf16(void* y, const void* x, const void* W, int64_t n, int64_t k, int nseg, const void* norm_w,
                         float eps, const void* bias, const void* residual, int silu_mul, void* stream) {
  using namespace b200;
  B200_CHECK_ARG(y && x && W, "gemv_fused: null pointer");
  B200_CHECK_ARG(!(silu_mul && (bias || residual)), "gemv_fused: silu_mul excludes bias/residual");
  B200_CHECK_ARG(!(bias && residual), "gemv_fused: bias and residual are exclusive");
  B200_CHECK_ARG((nseg == 2) == (silu_mul != 0), "gemv_fused: nseg == 2 exactly when silu_mul");
  int rc = b200_device_check();
  if (rc != B200_OK) return rc;
  if ((rc = gemv_setup_attributes()) != B200_OK) return rc;
  GemvPlan plan;
  const int pro = norm_w ? PRO_RMSNORM : PRO_PLAIN;
  const int epi = silu_mul ? EPI_SILU_MUL : (residual ? EPI_RESIDUAL : EPI_PLAIN);
  if (pro == PRO_RMSNORM && epi == EPI_RESIDUAL) {
    set_error("gemv_fused: RMSNorm prologue with residual epilogue is not a combination of the decode path");
    return B200_ERR_UNSUPPORTED;
  }
  if ((rc = gemv_make_plan(&plan, W, n * nseg, n, k, nseg, pro, epi, num_sms_cached())) != B200_OK) return rc;
  plan.p.x = (const __nv_bfloat16*)x;
  plan.p.norm_w = (const __nv_bfloat16*)norm_w;
  plan.p.eps = eps;
  plan.p.bias = (const __nv_bfloat16*)bias;
  plan.p.residual = (const __nv_bfloat16*)residual;
  plan.p.y = (__nv_bfloat16*)y;
  return gemv_launch(plan, (cudaStream_t)stream, false);
}

int64_t b200_attn_decode_workspace_bytes(int64_t Hq, int64_t Hkv, int64_t hd, int64_t max_ctx) {
  if (Hkv <= 0 || Hq % Hkv != 0 || (hd != 64 && hd != 128) || max_ctx < 1) return -1;
  const int nsplit = b200::attn_decode_nsplit((int)hd, (int)max_ctx);
  return b200::attn_decode_ws_floats((int)Hq, (int)Hkv, (int)hd, nsplit) * 4 + ((Hq * 4 + 15) / 16) * 16;
}

int b200_attn_decode_bf16(void* out, const void* qkv, const void* q_norm, const void* k_norm, float eps,
                          const float* rope_table, const int32_t* pos, int64_t fixed_len, void* kcache, void* vcache,
                          int64_t Hq, int64_t Hkv, int64_t hd, int64_t max_ctx, void* workspace, void* stream) {
  using namespace b200;
  B200_CHECK_ARG(out && qkv && kcache && vcache && workspace, "attn_decode: null pointer");
  B200_CHECK_ARG(pos != nullptr || (fixed_len >= 1 && fixed_len <= max_ctx), "attn_decode: fixed_len out of range");
  int rc = b200_device_check();
  if (rc != B200_OK) return rc;
  if ((rc = attn_setup_attributes()) != B200_OK) return rc;
  AttnDecodeParams a{};
  a.qkv = (const __nv_bfloat16*)qkv;
  a.q_norm = (const __nv_bfloat16*)q_norm;
  a.k_norm = (const __nv_bfloat16*)k_norm;
  a.eps = eps;
  a.rope = rope_table;
  a.pos = pos;
  a.fixed_len = (int)fixed_len;
  a.kcache = (__nv_bfloat16*)kcache;
  a.vcache = (__nv_bfloat16*)vcache;
  a.out = (__nv_bfloat16*)out;
  a.tickets = (unsigned int*)workspace;
  a.ws = (float*)((uint8_t*)workspace + ((Hq * 4 + 15) / 16) * 16);
  a.Hq = (int)Hq;
  a.Hkv = (int)Hkv;
  a.nsplit = attn_decode_nsplit((int)hd, (int)max_ctx);
  a.max_ctx = (int)max_ctx;
  return launch_attn_decode(a, (int)hd, (cudaStream_t)stream, false);
}

}  // extern "C"
