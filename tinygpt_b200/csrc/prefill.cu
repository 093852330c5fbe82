// prefill.cu — the small kernels around the tcgen05 GEMM in the batched PREFILL path (S prompt tokens at once):
//   bias add (second rounding, like the reference's separate add kernel), q/k per-head RMSNorm + RoPE + in-place KV-cache
//   write for all S tokens, and causal attention of the chunk against the cache.
//   [ref: src/layer/Attention.h:71-112,156-163; TT/Operation/OpLinalg.cpp:273-275;
//    TT/Operation/OpNNLayerCuda.cuh:252-299,412-440; TFA/mma/kernel.cuh:18-203 (causal prefill)]
// Decode never comes here.  The attention kernel is a straightforward CUDA-core one (one CTA per query row and head,
// online softmax over 128-key chunks); a tensor-core flash-attention prefill is the next step.
#include "common.cuh"
#include "ops.cuh"

namespace b200 {

namespace {

// y[r, j] = bf16(y[r, j] + b[j])
__global__ void __launch_bounds__(256) bias_add_kernel(__nv_bfloat16* __restrict__ y, const __nv_bfloat16* __restrict__ b,
                                                       int64_t rows, int64_t n) {
  const int64_t total = rows * n;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x)
    y[i] = __hadd(y[i], b[i % n]);
}

// One warp per (token, head) over the q heads and the k heads of the merged qkv rows [S, qdim + 2 kvdim]:
// optional per-head RMSNorm, RoPE at position p0 + t, q written back in place, k and v written to the cache row.
template <int HD>
__global__ void __launch_bounds__(256) prefill_qk_kernel(__nv_bfloat16* __restrict__ qkv, const __nv_bfloat16* __restrict__ q_norm,
                                                         const __nv_bfloat16* __restrict__ k_norm, float eps,
                                                         const float* __restrict__ rope, __nv_bfloat16* __restrict__ kcache,
                                                         __nv_bfloat16* __restrict__ vcache, int S, int Hq, int Hkv, int p0) {
  constexpr int EPL = HD / 32;
  const int lane = threadIdx.x & 31;
  const int64_t wid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int heads = Hq + Hkv;
  if (wid >= (int64_t)S * heads) return;
  const int t = (int)(wid / heads), h = (int)(wid % heads);
  const bool is_k = h >= Hq;
  const int qdim = Hq * HD, kvdim = Hkv * HD;
  const int64_t rowoff = (int64_t)t * (qdim + 2 * kvdim);
  __nv_bfloat16* src = qkv + rowoff + (is_k ? qdim + (h - Hq) * HD : h * HD);
  const __nv_bfloat16* nw = is_k ? k_norm : q_norm;
  float x[EPL];
#pragma unroll
  for (int j = 0; j < EPL; ++j) x[j] = bf16_to_f(src[lane + 32 * j]);
  if (nw != nullptr) {
    float ss = 0.f;
#pragma unroll
    for (int j = 0; j < EPL; ++j) ss += x[j] * x[j];
    ss = warp_sum(ss);
    const float inv = rsqrtf(ss / (float)HD + eps);
#pragma unroll
    for (int j = 0; j < EPL; ++j) x[j] = round_bf16(x[j] * inv * bf16_to_f(nw[lane + 32 * j]));
  }
  const float* row = rope + (size_t)(p0 + t) * HD * 2;
#pragma unroll
  for (int j = 0; j < EPL / 2; ++j) {
    const int i = lane + 32 * j;
    const float c = row[i * 2], s = row[i * 2 + 1];
    const float x1 = x[j], x2 = x[j + EPL / 2];
    x[j] = round_bf16(x1 * c - x2 * s);
    x[j + EPL / 2] = round_bf16(x2 * c + x1 * s);
  }
  if (is_k) {
    const int kh = h - Hq;
    __nv_bfloat16* kd = kcache + ((size_t)(p0 + t) * Hkv + kh) * HD;
    __nv_bfloat16* vd = vcache + ((size_t)(p0 + t) * Hkv + kh) * HD;
    const __nv_bfloat16* vs = qkv + rowoff + qdim + kvdim + kh * HD;
#pragma unroll
    for (int j = 0; j < EPL; ++j) {
      kd[lane + 32 * j] = f_to_bf16(x[j]);
      vd[lane + 32 * j] = vs[lane + 32 * j];
    }
  } else {
#pragma unroll
    for (int j = 0; j < EPL; ++j) src[lane + 32 * j] = f_to_bf16(x[j]);
  }
}

template <int HD>
struct InvSqrtHdP;
template <>
struct InvSqrtHdP<64> {
  static constexpr float value = 0.125f;
};
template <>
struct InvSqrtHdP<128> {
  static constexpr float value = 0.08838834764831845f;
};

// Causal attention of S query rows (positions p0 … p0+S-1) against cache rows 0 … p0+row: one CTA per (row, head).
template <int HD>
__global__ void __launch_bounds__(128) attn_prefill_kernel(__nv_bfloat16* __restrict__ o, const __nv_bfloat16* __restrict__ qkv,
                                                           const __nv_bfloat16* __restrict__ kcache,
                                                           const __nv_bfloat16* __restrict__ vcache, int Hq, int Hkv,
                                                           int p0) {
  constexpr float kScale = InvSqrtHdP<HD>::value * 1.4426950408889634f;
  __shared__ float qs[HD];
  __shared__ float ps[128];
  __shared__ float red[4];
  const int row = blockIdx.x, h = blockIdx.y;
  const int kvh = h / (Hq / Hkv);
  const int tid = threadIdx.x;
  const int qdim = Hq * HD, kvdim = Hkv * HD;
  const __nv_bfloat16* qp = qkv + (size_t)row * (qdim + 2 * kvdim) + h * HD;
  for (int d = tid; d < HD; d += 128) qs[d] = bf16_to_f(qp[d]);
  __syncthreads();
  const int nkeys = p0 + row + 1;
  float m_run = -INFINITY, l_run = 0.f, acc = 0.f;
  for (int c0 = 0; c0 < nkeys; c0 += 128) {
    const int j = c0 + tid;
    float s = -INFINITY;
    if (j < nkeys) {
      const uint4* kp = reinterpret_cast<const uint4*>(kcache + ((size_t)j * Hkv + kvh) * HD);
      float d = 0.f;
#pragma unroll
      for (int piece = 0; piece < HD / 8; ++piece) {
        float qf[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) qf[e] = qs[piece * 8 + e];
        d = dot8(kp[piece], qf, d);
      }
      s = d;
    }
    float cm = warp_max(s);
    if ((tid & 31) == 0) red[tid >> 5] = cm;
    __syncthreads();
    cm = fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3]));
    __syncthreads();
    const float m_new = fmaxf(m_run, cm);
    const float corr = (m_run == -INFINITY) ? 0.f : exp2f((m_run - m_new) * kScale);
    const float e = (j < nkeys) ? exp2f(s * kScale - m_new * kScale) : 0.f;
    ps[tid] = e;
    float cs = warp_sum(e);
    if ((tid & 31) == 0) red[tid >> 5] = cs;
    __syncthreads();
    cs = red[0] + red[1] + red[2] + red[3];
    l_run = l_run * corr + cs;
    m_run = m_new;
    if (tid < HD) {
      acc *= corr;
      const int n = min(128, nkeys - c0);
      for (int jj = 0; jj < n; ++jj)
        acc = fmaf(ps[jj], bf16_to_f(vcache[((size_t)(c0 + jj) * Hkv + kvh) * HD + tid]), acc);
    }
    __syncthreads();
  }
  if (tid < HD) o[(size_t)row * qdim + h * HD + tid] = f_to_bf16(acc * (l_run > 0.f ? 1.f / l_run : 0.f));
}

}  // namespace

int launch_bias_add(void* y, const void* b, int64_t rows, int64_t n, cudaStream_t st) {
  B200_CHECK_ARG(y && b && rows > 0 && n > 0, "bias_add: bad arguments");
  int64_t g = (rows * n + 255) / 256;
  if (g > 148 * 16) g = 148 * 16;
  g_launches.fetch_add(1);
  bias_add_kernel<<<(unsigned)g, 256, 0, st>>>((__nv_bfloat16*)y, (const __nv_bfloat16*)b, rows, n);
  B200_CUDA(cudaGetLastError());
  return B200_OK;
}

int launch_prefill_qk(void* qkv, const void* q_norm, const void* k_norm, float eps, const float* rope, void* kcache,
                      void* vcache, int S, int Hq, int Hkv, int hd, int p0, cudaStream_t st) {
  B200_CHECK_ARG(hd == 64 || hd == 128, "prefill: head_dim %d not built", hd);
  const int64_t warps = (int64_t)S * (Hq + Hkv);
  const unsigned grid = (unsigned)((warps * 32 + 255) / 256);
  g_launches.fetch_add(1);
  if (hd == 64)
    prefill_qk_kernel<64><<<grid, 256, 0, st>>>((__nv_bfloat16*)qkv, (const __nv_bfloat16*)q_norm,
                                                 (const __nv_bfloat16*)k_norm, eps, rope, (__nv_bfloat16*)kcache,
                                                 (__nv_bfloat16*)vcache, S, Hq, Hkv, p0);
  else
    prefill_qk_kernel<128><<<grid, 256, 0, st>>>((__nv_bfloat16*)qkv, (const __nv_bfloat16*)q_norm,
                                                  (const __nv_bfloat16*)k_norm, eps, rope, (__nv_bfloat16*)kcache,
                                                  (__nv_bfloat16*)vcache, S, Hq, Hkv, p0);
  B200_CUDA(cudaGetLastError());
  return B200_OK;
}

int launch_attn_prefill(void* o, const void* qkv, const void* kcache, const void* vcache, int S, int Hq, int Hkv, int hd,
                        int p0, cudaStream_t st) {
  B200_CHECK_ARG(hd == 64 || hd == 128, "prefill: head_dim %d not built", hd);
  if (prefill_attn_mma_enabled()) return launch_attn_prefill_mma(o, qkv, kcache, vcache, S, Hq, Hkv, hd, p0, st);
  dim3 grid((unsigned)S, (unsigned)Hq);
  g_launches.fetch_add(1);
  if (hd == 64)
    attn_prefill_kernel<64><<<grid, 128, 0, st>>>((__nv_bfloat16*)o, (const __nv_bfloat16*)qkv,
                                                  (const __nv_bfloat16*)kcache, (const __nv_bfloat16*)vcache, Hq, Hkv, p0);
  else
    attn_prefill_kernel<128><<<grid, 128, 0, st>>>((__nv_bfloat16*)o, (const __nv_bfloat16*)qkv,
                                                   (const __nv_bfloat16*)kcache, (const __nv_bfloat16*)vcache, Hq, Hkv, p0);
  B200_CUDA(cudaGetLastError());
  return B200_OK;
}

}  // namespace b200
