// common.cuh — shared device helpers for the sm_100a decode kernels (bf16 packing, mbarrier / TMA / PDL PTX wrappers)
// and the host-side error plumbing of the C ABI (include/b200_decode.h).
#pragma once

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>

#include "../../include/b200_decode.h"

namespace b200 {

// ---------------------------------------------------------------------------------------------------- host errors
void set_error(const char* fmt, ...);

// ------------------------------------------------------------------------------- selectable code paths and their defaults
// The environment variable overrides the default BOTH ways (the named value), so a measured alternative stays one
// variable away.  Variants that were measured and lost (flag synchronisation, register-resident small-k GEMV loop,
// cross-kernel L2 prefetch, whole-token persistent kernel) are not here any more: profiles/experiments/.
struct Defaults {
  static constexpr bool kPrefillAttnMma = true;   // B200_PREFILL_ATTN   "mma" (tensor cores) | "cuda" (CUDA cores)
  // B200_GEMM "persistent" | "tile": chosen per shape when unset (gemm.cu launch_gemm_bf16)
};
bool env_flag(const char* name, bool dflt);                       // unset → dflt; else first character == '1'
bool env_choice(const char* name, char yes_initial, bool dflt);   // unset → dflt; else first character == yes_initial
int env_int(const char* name, int dflt, int lo, int hi);          // unset / unparsable → dflt; clamped to [lo, hi]
extern std::atomic<int64_t> g_launches;

#define B200_CHECK_ARG(cond, ...)     \
  do {                                \
    if (!(cond)) {                    \
      ::b200::set_error(__VA_ARGS__); \
      return B200_ERR_INVALID;        \
    }                                 \
  } while (0)

#define B200_CUDA(call)                                                                                 \
  do {                                                                                                  \
    cudaError_t e__ = (call);                                                                           \
    if (e__ != cudaSuccess) {                                                                           \
      ::b200::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__);   \
      return B200_ERR_CUDA;                                                                             \
    }                                                                                                   \
  } while (0)

// Kernel launch through cudaLaunchKernelEx with the programmatic-dependent-launch attribute (PDL): the kernel may
// start while its predecessor on the stream is still running and must call pdl_wait() before touching anything the
// predecessor writes.  Weight prefetch (TMA into shared memory) is issued BEFORE pdl_wait().
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                              bool pdl, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

#ifdef __CUDACC__
// -------------------------------------------------------------------------------------------------- device helpers
constexpr int kWarp = 32;

__device__ __forceinline__ float bf16_lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t u) { return __uint_as_float(u & 0xffff0000u); }
__device__ __forceinline__ float bf16_to_f(__nv_bfloat16 v) { return __bfloat162float(v); }
__device__ __forceinline__ __nv_bfloat16 f_to_bf16(float v) { return __float2bfloat16_rn(v); }
// round-trip through bf16 (the reference's "static_cast<T>(fp32)" rounding point)
__device__ __forceinline__ float round_bf16(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }

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

// 8 bf16 dotted with 8 fp32 values into TWO independent accumulators (even / odd elements): halves the dependent FMA
// chain of the GEMV inner loop; the caller adds the pair once per row.
__device__ __forceinline__ void dot8x2(const uint4& w, const float (&x)[8], float& a0, float& a1) {
  a0 = fmaf(bf16_lo(w.x), x[0], a0);
  a1 = fmaf(bf16_hi(w.x), x[1], a1);
  a0 = fmaf(bf16_lo(w.y), x[2], a0);
  a1 = fmaf(bf16_hi(w.y), x[3], a1);
  a0 = fmaf(bf16_lo(w.z), x[4], a0);
  a1 = fmaf(bf16_hi(w.z), x[5], a1);
  a0 = fmaf(bf16_lo(w.w), x[6], a0);
  a1 = fmaf(bf16_hi(w.w), x[7], a1);
}

// 8 bf16 (one 16-byte vector) dotted with 8 fp32 values
__device__ __forceinline__ float dot8(const uint4& w, const float (&x)[8], float acc) {
  acc = fmaf(bf16_lo(w.x), x[0], acc);
  acc = fmaf(bf16_hi(w.x), x[1], acc);
  acc = fmaf(bf16_lo(w.y), x[2], acc);
  acc = fmaf(bf16_hi(w.y), x[3], acc);
  acc = fmaf(bf16_lo(w.z), x[4], acc);
  acc = fmaf(bf16_hi(w.z), x[5], acc);
  acc = fmaf(bf16_lo(w.w), x[6], acc);
  acc = fmaf(bf16_hi(w.w), x[7], acc);
  return acc;
}
__device__ __forceinline__ void unpack8(const uint4& v, float (&x)[8]) {
  x[0] = bf16_lo(v.x); x[1] = bf16_hi(v.x);
  x[2] = bf16_lo(v.y); x[3] = bf16_hi(v.y);
  x[4] = bf16_lo(v.z); x[5] = bf16_hi(v.z);
  x[6] = bf16_lo(v.w); x[7] = bf16_hi(v.w);
}
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  __nv_bfloat162 p = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&p);
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

// ---- programmatic dependent launch
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

__device__ __forceinline__ unsigned long long global_timer_ns();

// ---- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}

// ---- TMA: 2-D tiled bulk tensor load global → shared, completion on an mbarrier (SASS: UTMALDG)
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, int32_t c0, int32_t c1,
                                            uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
// L2 eviction-priority hint for the weight stream: a decode token streams ~8 × the L2 capacity of weights, each byte
// once.  Tagging the stream evict_first keeps the lines every kernel re-reads (activations, KV cache, code) resident:
// measured −1.5 % per token on Qwen2.5-0.5B and Mistral-7B, −2.4 % on the batched Mistral step.  (evict_last on the RMSNorm
// weights measured no change.)  profiles/experiments/r02_l2_hints.txt
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void tma_load_2d_hint(void* smem_dst, const CUtensorMap* map, int32_t c0, int32_t c1,
                                                 uint64_t* bar, uint64_t pol) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "l"(pol)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// named barrier among a subset of the CTA's warps (id 1..15)
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
#endif  // __CUDACC__

// 2-D bf16 row-major [rows, cols] tensor map with box {box_cols, box_rows}, no swizzle (host).
int make_tmap_2d_bf16(CUtensorMap* out, const void* base, int64_t rows, int64_t cols, int box_rows, int box_cols);

}  // namespace b200
