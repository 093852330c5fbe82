// attn.cu — decode attention for sm_100a: one launch per layer does what the reference does with
//   split (3 memcpy) + [Qwen3 q/k RMSNorm ×2] + RoPE ×2 + KV concat (4 memcpy, O(ctx) re-copy) + TinyFA flash-attn
//   [ref: src/layer/Attention.h:71-112,156-163; src/engine/CacheManager.h:24-42;
//    TT/Operation/OpNNLayerCuda.cuh:252-299 (norm), :412-440 (rope);
//    TFA/mma/kernel.cuh:18-203, TFA/mma/softmax.cuh:67-131 (online softmax, base-2 exponent)]
//
// Why not the reference's shape: TinyFA launches one CTA per Q head with a 128-row Q tile of which one row is live
// and re-reads K/V once per Q head.  Here a CTA owns (KV head, KV split): the G = Hq/Hkv query heads that share the
// KV head are processed together so K/V are read once, the context is split across `nsplit` CTAs so the whole GPU
// works on one token, and the last CTA of a KV head (atomic ticket) merges the split partials — no second launch.
// The KV cache is a pre-allocated [max_ctx, Hkv, hd] buffer written in place at position *pos.
//
// Numerics: q·k and P·V accumulate in fp32 over bf16 inputs; rounding points of q/k (after norm, after RoPE) are the
// reference's.  Probabilities stay fp32 (the reference rounds P to bf16 before P·V, TFA/mma/layout.cuh:88-97); the
// oracle models the reference's rounding and the parity tests bound the difference.
#include "ops.cuh"

#include <algorithm>
#include <mutex>

namespace b200 {

namespace {

constexpr int kAttnThreads = 256;
constexpr int kAttnWarps = 8;
constexpr float kLog2e = 1.4426950408889634f;

template <int HD>
struct InvSqrtHd;
template <>
struct InvSqrtHd<64> {
  static constexpr float value = 0.125f;
};
template <>
struct InvSqrtHd<128> {
  static constexpr float value = 0.08838834764831845f;
};


// keys per split: fixed by the shared-memory staging buffers (K and V of one split = 64 KB for both head dims)
template <int HD>
struct SplitKeys {
  static constexpr int value = (HD == 64) ? 256 : 128;
};

// ------------------------------------------------------------------------------------------------- decode kernel
// CTA = (KV split, group of G query heads that share one KV head).  Round 1 staged K/V of the split in shared memory
// (cp.async) and ran scores / softmax / P·V as three barrier-separated phases: 2.0 µs per launch at ctx ≈ 80 for 69 KB
// of traffic (in-kernel stamps: q/k/v staging 0.54, scores 0.54, softmax 0.64, P·V 0.3).  This kernel keeps everything
// a thread needs in registers (measured against the staged kernel on B200: 409 vs 415 µs per Qwen2.5-0.5B token,
// 1 425 vs 1 445 µs for Llama-3.2-3B, 1 106 vs 1 131 µs at ctx 2 048 for Qwen3-1.7B):
//   * thread = (key group, 16-byte piece): the K and V pieces of up to 8 cached rows per thread are loaded into
//     REGISTERS before griddepcontrol.wait (the rows were written by earlier tokens), no shared-memory staging;
//   * after the wait: q/k RMSNorm + RoPE by one warp per head as before, ONE barrier, then every thread runs an online
//     softmax over its own ≤ 8 keys (scores never leave registers), the key groups of a warp merge by xor-shuffle, the
//     8 warps through 5 KB of shared memory: two more barriers per query head, none for scores / softmax / P·V.
// The new K/V row is written in place to the cache and handed to the owning thread through shared memory.
template <int HD>
__global__ void __launch_bounds__(kAttnThreads) attn_decode_reg_kernel(const AttnDecodeParams p_in, const int G) {
  constexpr int LPK = HD / 8;             // lanes per key row
  constexpr int NG = kAttnThreads / LPK;  // key rows per pass
  constexpr int KPT = 8;                  // key rows per thread
  constexpr int CHUNK = NG * KPT;         // = SplitKeys<HD>: 256 (hd 64) / 128 (hd 128)
  static_assert(CHUNK == SplitKeys<HD>::value, "split size is shared with the host-side nsplit");
  constexpr int EPL = HD / 32;
  constexpr float kScale = InvSqrtHd<HD>::value * kLog2e;
  __shared__ float q_s[8 * HD];
  __shared__ __align__(16) __nv_bfloat16 knew[HD];
  __shared__ __align__(16) __nv_bfloat16 vnew[HD];
  __shared__ float ared[kAttnWarps * 16 * 10];
  __shared__ int is_last;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int hg = blockIdx.y, split = blockIdx.x;
  AttnDecodeParams p = p_in;
  {  // batched decode: this CTA's sequence
    const long long b = blockIdx.z;
    p.qkv += b * p.qkv_bstride;
    p.out += b * p.out_bstride;
    p.kcache += b * p.cache_bstride;
    p.vcache += b * p.cache_bstride;
    p.ws += b * p.ws_bstride;
    p.tickets += b * p.tick_bstride;
  }
  const int h0 = hg * G;
  const int group = p.Hq / p.Hkv;
  const int kvh = h0 / group;
  const bool kv_leader = (h0 % group) == 0;
  if (p.trace != nullptr && hg == 0 && split == 0 && blockIdx.z == 0 && tid == 0) p.trace[0] = global_timer_ns();
  pdl_trigger();
  const bool append = (p.pos != nullptr);
  const int pos = append ? *p.pos : p.fixed_len - 1;   // stable since before the producer kernel started (engine.cu)
  const int L = pos + 1;
  const int nact = (L + CHUNK - 1) / CHUNK;
  if (split >= nact) {
    pdl_wait();
    return;
  }
  const int start = split * CHUNK;
  const int end = min(L, start + CHUNK);
  const int n_old_end = append ? min(end, pos) : end;    // rows [start, n_old_end) are already in the cache
  const bool owns_new = append && pos >= start && pos < end;
  const int grp = tid / LPK, part = tid % LPK;
  const int imax = (end - start + NG - 1) / NG;   // key slots per thread that this split can fill (CTA-uniform)

  uint4 kreg[KPT], vreg[KPT];
#pragma unroll
  for (int i = 0; i < KPT; ++i) {
    const int row = start + grp + NG * i;
    kreg[i] = make_uint4(0, 0, 0, 0);
    vreg[i] = make_uint4(0, 0, 0, 0);
    if (i < imax && row < n_old_end) {
      const size_t g = ((size_t)row * p.Hkv + kvh) * HD + part * 8;
      kreg[i] = *reinterpret_cast<const uint4*>(p.kcache + g);
      vreg[i] = *reinterpret_cast<const uint4*>(p.vcache + g);
    }
  }
  float rc[EPL / 2], rs[EPL / 2];
  if (p.rope != nullptr) {
    const float* row = p.rope + (size_t)pos * HD * 2;
#pragma unroll
    for (int j = 0; j < EPL / 2; ++j) {
      rc[j] = row[(lane + 32 * j) * 2];
      rs[j] = row[(lane + 32 * j) * 2 + 1];
    }
  }
  pdl_wait();   // qkv of this token is complete and visible from here on
  if (p.trace != nullptr && hg == 0 && split == 0 && tid == 0) p.trace[1] = global_timer_ns();
  const int qdim = p.Hq * HD, kvdim = p.Hkv * HD;
  for (int h = warp; h < G + 1; h += kAttnWarps) {
    const bool is_k = (h == G);
    if (is_k && !owns_new) break;
    const __nv_bfloat16* src = is_k ? p.qkv + qdim + kvh * HD : p.qkv + (h0 + h) * HD;
    const __nv_bfloat16* nw = is_k ? p.k_norm : p.q_norm;
    float x[EPL];
#pragma unroll
    for (int j = 0; j < EPL; ++j) x[j] = bf16_to_f(src[lane + 32 * j]);
    if (nw != nullptr) {
      float ss = 0.f;
#pragma unroll
      for (int j = 0; j < EPL; ++j) ss += x[j] * x[j];
      ss = warp_sum(ss);
      const float inv = rsqrtf(ss / (float)HD + p.eps);
#pragma unroll
      for (int j = 0; j < EPL; ++j) x[j] = round_bf16(x[j] * inv * bf16_to_f(nw[lane + 32 * j]));
    }
    if (p.rope != nullptr) {
#pragma unroll
      for (int j = 0; j < EPL / 2; ++j) {
        const float x1 = x[j], x2 = x[j + EPL / 2];
        x[j] = round_bf16(x1 * rc[j] - x2 * rs[j]);
        x[j + EPL / 2] = round_bf16(x2 * rc[j] + x1 * rs[j]);
      }
    }
    if (is_k) {
      __nv_bfloat16* kg = p.kcache + ((size_t)pos * p.Hkv + kvh) * HD;
      __nv_bfloat16* vg = p.vcache + ((size_t)pos * p.Hkv + kvh) * HD;
#pragma unroll
      for (int j = 0; j < EPL; ++j) {
        const __nv_bfloat16 kk = f_to_bf16(x[j]);
        const __nv_bfloat16 vv = p.qkv[qdim + kvdim + kvh * HD + lane + 32 * j];
        knew[lane + 32 * j] = kk;
        vnew[lane + 32 * j] = vv;
        if (kv_leader) {
          kg[lane + 32 * j] = kk;
          vg[lane + 32 * j] = vv;
        }
      }
    } else {
#pragma unroll
      for (int j = 0; j < EPL; ++j) q_s[h * HD + lane + 32 * j] = x[j];
    }
  }
  __syncthreads();
  if (p.trace != nullptr && hg == 0 && split == 0 && tid == 0) p.trace[3] = global_timer_ns();
  if (owns_new) {
#pragma unroll
    for (int i = 0; i < KPT; ++i) {
      if (start + grp + NG * i == pos) {
        kreg[i] = reinterpret_cast<const uint4*>(knew)[part];
        vreg[i] = reinterpret_cast<const uint4*>(vnew)[part];
      }
    }
  }
  for (int g = 0; g < G; ++g) {
    float qf[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) qf[e] = q_s[g * HD + part * 8 + e];
    float sc[KPT];
    float m = -INFINITY;
#pragma unroll
    for (int i = 0; i < KPT; ++i) {
      sc[i] = -INFINITY;
      if (i < imax) {   // uniform: unused slots cost one branch
        float d = dot8(kreg[i], qf, 0.f);
#pragma unroll
        for (int o = LPK / 2; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
        if ((start + grp + NG * i) < end) sc[i] = d;
        m = fmaxf(m, sc[i]);
      }
    }
    float l = 0.f, acc[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = 0.f;
    const float m_scaled = m * kScale;
#pragma unroll
    for (int i = 0; i < KPT; ++i) {
      if (i < imax) {
        const float pr = (sc[i] == -INFINITY) ? 0.f : exp2f(sc[i] * kScale - m_scaled);
        l += pr;
        float vf[8];
        unpack8(vreg[i], vf);
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] = fmaf(pr, vf[e], acc[e]);
      }
    }
#pragma unroll
    for (int o = LPK; o < 32; o <<= 1) {   // key groups that live in the same warp
      const float mo = __shfl_xor_sync(0xffffffffu, m, o);
      const float lo = __shfl_xor_sync(0xffffffffu, l, o);
      const float mn = fmaxf(m, mo);
      const float a = (m == -INFINITY) ? 0.f : exp2f((m - mn) * kScale);
      const float b = (mo == -INFINITY) ? 0.f : exp2f((mo - mn) * kScale);
      l = l * a + lo * b;
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float ao = __shfl_xor_sync(0xffffffffu, acc[e], o);
        acc[e] = acc[e] * a + ao * b;
      }
      m = mn;
    }
    if (lane < LPK) {
      float* r = ared + (warp * 16 + lane) * 10;
      r[0] = m;
      r[1] = l;
#pragma unroll
      for (int e = 0; e < 8; ++e) r[2 + e] = acc[e];
    }
    __syncthreads();
    if (tid < HD) {
      const int d = tid, pt = d >> 3, e = d & 7;
      float M = -INFINITY;
#pragma unroll
      for (int w = 0; w < kAttnWarps; ++w) M = fmaxf(M, ared[(w * 16 + pt) * 10]);
      float num = 0.f, den = 0.f;
#pragma unroll
      for (int w = 0; w < kAttnWarps; ++w) {
        const float* r = ared + (w * 16 + pt) * 10;
        const float wgt = (r[0] == -INFINITY) ? 0.f : exp2f((r[0] - M) * kScale);
        num = fmaf(wgt, r[2 + e], num);
        den = fmaf(wgt, r[1], den);
      }
      if (nact == 1) {
        p.out[(h0 + g) * HD + d] = f_to_bf16(num * (den > 0.f ? 1.f / den : 0.f));
      } else {
        float* wsr = p.ws + ((size_t)(h0 + g) * p.nsplit + split) * (HD + 2);
        wsr[d] = num;
        if (d == 0) {
          wsr[HD] = M;
          wsr[HD + 1] = den;
        }
      }
    }
    __syncthreads();
  }
  if (p.trace != nullptr && hg == 0 && split == 0 && tid == 0) p.trace[2] = global_timer_ns();
  if (nact == 1) return;
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    const unsigned int t = atomicAdd(&p.tickets[hg], 1u);
    is_last = (t == (unsigned int)nact - 1) ? 1 : 0;
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  for (int idx = tid; idx < G * HD; idx += kAttnThreads) {
    const int g = idx / HD, d = idx % HD;
    const float* base = p.ws + (size_t)(h0 + g) * p.nsplit * (HD + 2);
    float M = -INFINITY;
    for (int s = 0; s < nact; ++s) M = fmaxf(M, __ldcg(base + (size_t)s * (HD + 2) + HD));
    float num = 0.f, den = 0.f;
    for (int s = 0; s < nact; ++s) {
      const float* r = base + (size_t)s * (HD + 2);
      const float w = exp2f((__ldcg(r + HD) - M) * kScale);
      num = fmaf(w, __ldcg(r + d), num);
      den = fmaf(w, __ldcg(r + HD + 1), den);
    }
    p.out[(h0 + g) * HD + d] = f_to_bf16(num * (den > 0.f ? 1.f / den : 0.f));
  }
  if (tid == 0) p.tickets[hg] = 0;
}

// ------------------------------------------------------------------------------------------ general attention
// One CTA per (q row, head, batch); keys in chunks of 128 with online softmax.  Correctness path (prefill-shaped
// calls through b200_attn_bf16); the tensor-core prefill kernel is separate work.
template <int HD>
__global__ void __launch_bounds__(128) attn_general_kernel(__nv_bfloat16* __restrict__ o,
                                                           const __nv_bfloat16* __restrict__ q,
                                                           const __nv_bfloat16* __restrict__ k,
                                                           const __nv_bfloat16* __restrict__ v, int Sq, int Skv, int Hq,
                                                           int Hkv, int causal) {
  constexpr float kScale = InvSqrtHd<HD>::value * kLog2e;
  __shared__ float qs[HD];
  __shared__ float ps[128];
  __shared__ float red[4];
  const int row = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int kvh = h / (Hq / Hkv);
  const int tid = threadIdx.x;
  const __nv_bfloat16* qp = q + (((size_t)b * Sq + row) * Hq + h) * HD;
  for (int d = tid; d < HD; d += 128) qs[d] = bf16_to_f(qp[d]);
  __syncthreads();
  const int nkeys = causal ? min(Skv, row + 1) : Skv;
  float m_run = -INFINITY, l_run = 0.f, acc = 0.f;  // thread d < HD owns output column d
  for (int c0 = 0; c0 < nkeys; c0 += 128) {
    const int j = c0 + tid;
    float s = -INFINITY;
    if (j < nkeys) {
      const uint4* kp = reinterpret_cast<const uint4*>(k + (((size_t)b * Skv + j) * Hkv + kvh) * HD);
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
        acc = fmaf(ps[jj], bf16_to_f(v[(((size_t)b * Skv + c0 + jj) * Hkv + kvh) * HD + tid]), acc);
    }
    __syncthreads();
  }
  if (tid < HD) o[(((size_t)b * Sq + row) * Hq + h) * HD + tid] = f_to_bf16(acc * (l_run > 0.f ? 1.f / l_run : 0.f));
}

}  // namespace

int64_t attn_decode_ws_floats(int Hq, int Hkv, int hd, int nsplit) {
  (void)Hkv;
  return (int64_t)Hq * nsplit * (hd + 2);  // same size whatever the head grouping
}

// Query heads per CTA: the whole GQA group once the context is long enough for K/V re-reads to matter, else one.
// The kernel takes any G ≤ 8 that divides the GQA group (the reference accepts every Hq % Hkv == 0): the largest such
// divisor (6 → 6, 5 → 5, 16 → 8, 12 → 6).
int attn_heads_per_cta(int Hq, int Hkv, int max_ctx) {
  if (max_ctx <= 1024) return 1;
  const int group = Hq / Hkv;
  for (int g = 8; g > 1; --g)
    if (group % g == 0) return g;
  return 1;
}

int attn_setup_attributes() { return B200_OK; }   // the decode kernel uses static shared memory only

int attn_decode_nsplit(int hd, int max_ctx) {
  const int chunk = (hd == 64) ? SplitKeys<64>::value : SplitKeys<128>::value;
  return std::max(1, (max_ctx + chunk - 1) / chunk);
}

int launch_attn_decode(const AttnDecodeParams& p, int hd, cudaStream_t st, bool pdl) {
  B200_CHECK_ARG(hd == 64 || hd == 128, "attention: head_dim %d not built (64 and 128 are, like the reference)", hd);
  B200_CHECK_ARG(p.Hkv > 0 && p.Hq % p.Hkv == 0, "attention: Hq=%d must be a multiple of Hkv=%d", p.Hq, p.Hkv);
  const int G = p.heads_per_cta > 0 ? p.heads_per_cta : attn_heads_per_cta(p.Hq, p.Hkv, p.max_ctx);
  B200_CHECK_ARG((p.Hq / p.Hkv) % G == 0, "attention: heads per CTA %d must divide the GQA group %d", G, p.Hq / p.Hkv);
  B200_CHECK_ARG(p.max_ctx >= 1 && p.nsplit == attn_decode_nsplit(hd, p.max_ctx),
                 "attention: nsplit %d does not match max_ctx %d (need %d)", p.nsplit, p.max_ctx,
                 attn_decode_nsplit(hd, p.max_ctx));
  B200_CHECK_ARG(p.nsplit <= 65535, "attention: context too long");
  B200_CHECK_ARG(G >= 1 && G <= 8, "attention: at most 8 query heads per CTA (got %d)", G);
  const unsigned nb = (unsigned)std::max(1, p.batch);
  if (hd == 64)
    B200_CUDA(launch_pdl(attn_decode_reg_kernel<64>, dim3(p.nsplit, p.Hq / G, nb), dim3(kAttnThreads), 0, st, pdl, p, G));
  else
    B200_CUDA(launch_pdl(attn_decode_reg_kernel<128>, dim3(p.nsplit, p.Hq / G, nb), dim3(kAttnThreads), 0, st, pdl, p, G));
  return B200_OK;
}

int launch_attn_general(void* o, const void* q, const void* k, const void* v, int64_t B, int64_t Sq, int64_t Skv,
                        int64_t Hq, int64_t Hkv, int64_t hd, int causal, cudaStream_t st) {
  B200_CHECK_ARG(hd == 64 || hd == 128, "attention: head_dim %lld not built (64 and 128 are)", (long long)hd);
  B200_CHECK_ARG(B > 0 && Sq > 0 && Skv > 0 && Hkv > 0 && Hq % Hkv == 0 && Hq < 65536 && B < 65536,
                 "attention: bad shape");
  dim3 grid((unsigned)Sq, (unsigned)Hq, (unsigned)B);
  g_launches.fetch_add(1);
  if (hd == 64)
    attn_general_kernel<64><<<grid, 128, 0, st>>>((__nv_bfloat16*)o, (const __nv_bfloat16*)q, (const __nv_bfloat16*)k,
                                                  (const __nv_bfloat16*)v, (int)Sq, (int)Skv, (int)Hq, (int)Hkv, causal);
  else
    attn_general_kernel<128><<<grid, 128, 0, st>>>((__nv_bfloat16*)o, (const __nv_bfloat16*)q, (const __nv_bfloat16*)k,
                                                   (const __nv_bfloat16*)v, (int)Sq, (int)Skv, (int)Hq, (int)Hkv,
                                                   causal);
  B200_CUDA(cudaGetLastError());
  return B200_OK;
}

}  // namespace b200

extern "C" int b200_attn_bf16(void* o, const void* q, const void* k, const void* v, int64_t B, int64_t Sq, int64_t Skv,
                              int64_t Hq, int64_t Hkv, int64_t hd, int causal, void* stream) {
  using namespace b200;
  B200_CHECK_ARG(o && q && k && v, "attention: null pointer");
  if (causal && Sq == Skv && Sq > 1 && prefill_attn_mma_enabled())
    return launch_attn_causal_mma(o, q, k, v, B, Sq, Hq, Hkv, hd, (cudaStream_t)stream);
  return launch_attn_general(o, q, k, v, B, Sq, Skv, Hq, Hkv, hd, causal, (cudaStream_t)stream);
}
