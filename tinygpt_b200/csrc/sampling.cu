// sampling.cu — temperature / top-k / top-p / min-p sampling of one token from bf16 logits, on the device, WITHOUT
// sorting the vocabulary.  Stand-alone op behind b200_sample_bf16 and, through
// b200_engine_set_sampler, the last kernels of a token when any sampler knob is set.  On hardware: drawn index equal to
// the pinned sampler oracle over 10 configurations × 5 uniform numbers × 4 vocabularies, deterministic, and every token
// an engine draws replayable on the host (tests/test_async_sampler_gpu.py).
// Parity target: the reference's Sampler arithmetic in fp32 (its CPU path; the test-side sampler oracle is pinned against the
// reference's own Sampler.cpp).  The reference's CUDA path runs the same pipeline in the logits' dtype (bf16 divide,
// softmax and cumsum, src/engine/Sampler.cpp:36-55) and leaves equal logits at a cut to thrust's unstable sort: a
// top-p / min-p boundary can differ by an entry from it.  The uniform number is Philox(seed, tokens this ENGINE has
// generated so far): two generate calls on one engine continue the stream, like the reference's global generator.
//
// Replaces tinygpt::Sampler::sample + multinomial  [ref: src/engine/Sampler.cpp:23-78;
//   third_party/TinyTorch/src/Operation/OpSamplingCuda.cu:30-62 (inverse-CDF draw), :97-170 (topk = full thrust sort),
//   :261-330 (sort = thrust stable_sort_by_key)], which sorts all V logits up to three times per token.
//
// bf16 logits take at most 65 536 distinct values, so:
//   1. sample_hist_kernel   histogram of the order-preserving 16-bit keys (integer atomics: exact, deterministic)
//   2. sample_plan_kernel   one CTA walks the 65 536 bins from the largest value down with block prefix sums and decides
//                           how many entries of every bin survive top-k, top-p (keep while the running probability
//                           ≤ top_p, always the first) and min-p (p ≥ max p · min_p): whole bins, plus m entries of at
//                           most ONE partially kept bin
//   3. sample_draw_kernel   one CTA passes over the vocabulary in index order: an entry of the partial bin survives if
//                           its rank among equal keys is below m (lowest indices first — the oracle's tie rule),
//                           probability = e/Z, inclusive cdf, first index with cdf ≥ u·total (the reference's draw)
// Arithmetic follows the pinned CPU restatement of the sampler (tests/test_sampler_bins_model.py): fp32 value/temperature, e = expf(v − vmax), fp32 probabilities; sums of
// probabilities are accumulated in fp64 in a fixed order (the oracle adds fp32 one by one, the reference's thrust scan
// in yet another order: a top-p boundary that sits on top_p within rounding can flip by one entry in all three).
#include "common.cuh"
#include "ops.cuh"

#include <algorithm>

#include <cub/block/block_reduce.cuh>
#include <cub/block/block_scan.cuh>

namespace b200 {

namespace {

constexpr int kBins = 65536;
constexpr int kPlanThreads = 512;
constexpr int kBinsPerThread = kBins / kPlanThreads;   // 128 consecutive bins, walked downwards
constexpr int kDrawThreads = 512;
constexpr int kDrawItems = 8;                          // consecutive vocabulary entries per thread per chunk

struct SamplePlan {        // written by the plan kernel, read by the draw kernel
  double z;                // Σ kept · e over the surviving entries
  float vmax;              // largest value / temperature
  int partial_key;         // the one bin of which only `partial_keep` entries survive, or -1
  int partial_keep;
  int kmax;                // key of the largest logit
};

__device__ __forceinline__ unsigned int order_key(unsigned short bits) {
  return (bits & 0x8000u) ? (unsigned int)(unsigned short)~bits : (unsigned int)(bits | 0x8000u);
}
__device__ __forceinline__ float key_value(unsigned int key) {
  const unsigned int bits = (key & 0x8000u) ? (key & 0x7fffu) : ((~key) & 0xffffu);
  return __uint_as_float(bits << 16);
}
__device__ __forceinline__ float scaled_value(unsigned int key, float temperature) {
  const float v = key_value(key);
  return temperature > 0.f ? v / temperature : v;   // IEEE division, like the oracle
}

__global__ void __launch_bounds__(256) sample_hist_kernel(const __nv_bfloat16* __restrict__ logits, int64_t V,
                                                          unsigned int* __restrict__ hist) {
  const unsigned short* bits = reinterpret_cast<const unsigned short*>(logits);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < V; i += (int64_t)gridDim.x * blockDim.x)
    atomicAdd(&hist[order_key(bits[i])], 1u);
}

// One CTA, thread t owns bins [hi − 127, hi] with hi = 65535 − 128 t and walks them from hi downwards, so that "exclusive
// prefix over threads" + "running value inside the thread" is the prefix over all LARGER values.
__global__ void __launch_bounds__(kPlanThreads) sample_plan_kernel(const unsigned int* __restrict__ hist, int64_t V,
                                                                   float temperature, int64_t top_k, float top_p,
                                                                   float min_p, int* __restrict__ kept,
                                                                   SamplePlan* __restrict__ plan) {
  using ScanI = cub::BlockScan<long long, kPlanThreads>;
  using ScanD = cub::BlockScan<double, kPlanThreads>;
  using RedD = cub::BlockReduce<double, kPlanThreads>;
  using RedI = cub::BlockReduce<int, kPlanThreads>;
  __shared__ union {
    typename ScanI::TempStorage si;
    typename ScanD::TempStorage sd;
    typename RedD::TempStorage rd;
    typename RedI::TempStorage ri;
  } tmp;
  __shared__ int s_kmax;
  __shared__ double s_z;
  __shared__ int s_partial_key, s_partial_keep;

  const int t = threadIdx.x;
  const int hi = kBins - 1 - t * kBinsPerThread;
  if (t == 0) {
    s_partial_key = -1;
    s_partial_keep = 0;
  }

  // ---- largest present key
  int my_max = -1;
  for (int j = 0; j < kBinsPerThread; ++j)
    if (hist[hi - j] != 0u) {
      my_max = hi - j;
      break;
    }
  const int kmax_r = RedI(tmp.ri).Reduce(my_max, cub::Max());
  if (t == 0) s_kmax = kmax_r;
  __syncthreads();
  const int kmax = s_kmax;
  const float vmax = scaled_value((unsigned int)kmax, temperature);
  auto e_of = [&](int key) { return expf(scaled_value((unsigned int)key, temperature) - vmax); };

  // ---- top-k: entries kept per bin = clamp(k − #entries with a larger key, 0, count)
  const bool use_k = top_k > 0 && top_k < V;
  {
    long long mine = 0;
    for (int j = 0; j < kBinsPerThread; ++j) mine += hist[hi - j];
    long long before = 0;
    ScanI(tmp.si).ExclusiveSum(mine, before);
    __syncthreads();
    for (int j = 0; j < kBinsPerThread; ++j) {
      const long long c = hist[hi - j];
      long long k = c;
      if (use_k) k = min(max((long long)top_k - before, 0ll), c);
      kept[hi - j] = (int)k;
      before += c;
    }
  }
  __syncthreads();

  auto block_z = [&]() {   // Σ kept · e in a fixed order (thread-local walk, then cub's fixed reduction tree)
    double mine = 0.0;
    for (int j = 0; j < kBinsPerThread; ++j) {
      const int k = kept[hi - j];
      if (k != 0) mine += (double)k * (double)e_of(hi - j);
    }
    const double z = RedD(tmp.rd).Sum(mine);
    if (t == 0) s_z = z;
    __syncthreads();
    const double out = s_z;
    __syncthreads();
    return out;
  };

  // ---- top-p: keep entries while the running probability (sorted order) stays ≤ top_p; the first always survives
  if (top_p < 1.f) {
    const float zf = (float)block_z();
    double mine = 0.0;
    for (int j = 0; j < kBinsPerThread; ++j) {
      const int k = kept[hi - j];
      if (k != 0) mine += (double)k * (double)(e_of(hi - j) / zf);
    }
    double before = 0.0;
    ScanD(tmp.sd).ExclusiveSum(mine, before);
    __syncthreads();
    for (int j = 0; j < kBinsPerThread; ++j) {
      const int key = hi - j;
      const int k = kept[key];
      if (k == 0) continue;
      const float p = e_of(key) / zf;
      long long m = (long long)floor(((double)top_p - before) / (double)p + 1e-9);
      m = min(max(m, 0ll), (long long)k);
      if (key == kmax && m < 1) m = 1;
      kept[key] = (int)m;
      before += (double)k * (double)p;
    }
    __syncthreads();
  }

  // ---- min-p: drop bins whose probability is below max-probability · min_p
  if (min_p > 0.f) {
    const float zf = (float)block_z();
    const float thr = (e_of(kmax) / zf) * min_p;
    for (int j = 0; j < kBinsPerThread; ++j) {
      const int key = hi - j;
      if (kept[key] != 0 && e_of(key) / zf < thr) kept[key] = 0;
    }
    __syncthreads();
  }

  // ---- final normaliser and the (at most one) partially kept bin
  const double z = block_z();
  for (int j = 0; j < kBinsPerThread; ++j) {
    const int key = hi - j;
    const int k = kept[key];
    if (k != 0 && (unsigned int)k < hist[key]) {   // unique by construction (tests/test_sampler_bins_model.py)
      s_partial_key = key;
      s_partial_keep = k;
    }
  }
  __syncthreads();
  if (t == 0) {
    plan->z = z;
    plan->vmax = vmax;
    plan->partial_key = s_partial_key;
    plan->partial_keep = s_partial_keep;
    plan->kmax = kmax;
  }
}

// One CTA over the vocabulary in index order.  Pass 0: total probability.  Pass 1: first index whose inclusive cdf
// reaches r = u · total.  Thread t owns kDrawItems consecutive entries of each chunk, chunks are visited in order and a
// running carry (cdf so far, partial-bin entries seen so far) crosses chunks, so the cdf is the index-order cdf.
// Philox4x32-10 (Salmon et al., SC'11) written out here so that the host can reproduce the engine's uniform numbers:
// counter = {n lo, n hi, 0, 0} with n = tokens generated so far, key = seed; u = (x0 >> 8 + 0.5) · 2^-24 ∈ (0, 1).
__device__ __forceinline__ float philox_uniform(unsigned long long seed, unsigned long long n) {
  unsigned int c0 = (unsigned int)n, c1 = (unsigned int)(n >> 32), c2 = 0u, c3 = 0u;
  unsigned int k0 = (unsigned int)seed, k1 = (unsigned int)(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const unsigned int hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const unsigned int hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const unsigned int n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0;
    c1 = lo1;
    c2 = n2;
    c3 = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return ((float)(c0 >> 8) + 0.5f) * (1.0f / 16777216.0f);
}

// `pub` (engine only): publish the drawn token like the engine's argmax does (next step's input, on-device log, position,
// token mailbox); `rng_seed_valid`: draw u on the device from Philox(seed, tokens generated so far) instead of `u`.
__global__ void __launch_bounds__(kDrawThreads) sample_draw_kernel(const __nv_bfloat16* __restrict__ logits, int64_t V,
                                                                   float temperature, float u,
                                                                   const int* __restrict__ kept,
                                                                   const SamplePlan* __restrict__ plan,
                                                                   int64_t* __restrict__ out, const ArgmaxPublish pub,
                                                                   unsigned long long rng_seed, int rng_seed_valid) {
  using ScanD = cub::BlockScan<double, kDrawThreads>;
  using ScanI = cub::BlockScan<int, kDrawThreads>;
  __shared__ union {
    typename ScanD::TempStorage sd;
    typename ScanI::TempStorage si;
  } tmp;
  __shared__ double s_carry;
  __shared__ int s_rank_carry;
  __shared__ long long s_pick;

  const unsigned short* bits = reinterpret_cast<const unsigned short*>(logits);
  const double z = plan->z;
  const float zf = (float)z;
  const float vmax = plan->vmax;
  const int pkey = plan->partial_key, pkeep = plan->partial_keep;
  const int t = threadIdx.x;
  constexpr int kChunk = kDrawThreads * kDrawItems;
  double total = 0.0;
  if (rng_seed_valid) u = philox_uniform(rng_seed, *pub.gen_count);

  for (int pass = 0; pass < 2; ++pass) {
    if (t == 0) {
      s_carry = 0.0;
      s_rank_carry = 0;
      s_pick = -1;
    }
    __syncthreads();
    const double r = (double)u * total;   // pass 1 only
    for (int64_t base = 0; base < V; base += kChunk) {
      float p[kDrawItems];
      int is_partial[kDrawItems];
      int n_partial = 0;
      const int64_t i0 = base + (int64_t)t * kDrawItems;
#pragma unroll
      for (int j = 0; j < kDrawItems; ++j) {
        const int64_t i = i0 + j;
        p[j] = 0.f;
        is_partial[j] = 0;
        if (i < V) {
          const unsigned int key = order_key(bits[i]);
          const int k = kept[key];
          if (k != 0) {
            is_partial[j] = ((int)key == pkey);
            n_partial += is_partial[j];
            p[j] = expf(scaled_value(key, temperature) - vmax) / zf;
          }
        }
      }
      // rank of the partial bin's entries in index order: only the first `pkeep` survive
      int rank_before = 0;
      ScanI(tmp.si).ExclusiveSum(n_partial, rank_before);
      __syncthreads();
      int rank = s_rank_carry + rank_before;
      double mine = 0.0;
#pragma unroll
      for (int j = 0; j < kDrawItems; ++j) {
        if (is_partial[j]) {
          if (rank >= pkeep) p[j] = 0.f;
          ++rank;
        }
        mine += (double)p[j];
      }
      double before = 0.0, chunk_total = 0.0;
      ScanD(tmp.sd).ExclusiveSum(mine, before, chunk_total);
      __syncthreads();
      if (pass == 1) {   // no `s_pick < 0` guard: other warps may already be publishing into s_pick (atomicMin keeps the
                         // smallest index, and the loop leaves after the first chunk that found one)
        double c = s_carry + before;
        long long found = -1;
#pragma unroll
        for (int j = 0; j < kDrawItems; ++j) {
          c += (double)p[j];
          if (found < 0 && i0 + j < V && p[j] > 0.f && c >= r) found = i0 + j;   // only a surviving entry can be drawn
        }
        if (found >= 0) atomicMin(reinterpret_cast<unsigned long long*>(&s_pick), (unsigned long long)found);
      }
      __syncthreads();
      if (t == kDrawThreads - 1) {
        s_carry += chunk_total;
        s_rank_carry = rank;   // the last thread has seen every partial entry of the chunk
      }
      __syncthreads();
      if (pass == 1 && s_pick >= 0) break;   // uniform: s_pick is shared
    }
    if (pass == 0) {
      total = s_carry;
      __syncthreads();
    }
  }
  if (t == 0) {
    long long pick = s_pick;
    if (!(total > 0.0) || pick < 0) pick = (total > 0.0) ? (long long)(V - 1) : 0;   // the reference returns 0 when total ≤ 0
    if (out != nullptr) out[0] = pick;
    if (pub.cur_tok != nullptr) {   // same publication as argmax_kernel (ops.cu)
      *pub.cur_tok = pick;
      if (pub.pos != nullptr) *pub.pos += 1;
      const unsigned long long c = *pub.gen_count;
      pub.gen_log[c % (unsigned long long)pub.gen_cap] = pick;
      *pub.gen_count = c + 1;
      if (pub.mailbox != nullptr) {
        const unsigned long long word = (((c + 1ull) & 0xffffffffull) << 32) | (unsigned long long)(unsigned int)pick;
        asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(pub.mailbox + (c % pub.mailbox_cap)), "l"(word)
                     : "memory");
      }
    }
  }
}

}  // namespace

int64_t sample_workspace_bytes() { return (int64_t)kBins * 4 /*hist*/ + (int64_t)kBins * 4 /*kept*/ + 256 /*plan*/; }

int launch_sample(int64_t* token_out, const void* logits, int64_t V, float temperature, int64_t top_k, float top_p,
                  float min_p, float u, void* workspace, cudaStream_t st, const ArgmaxPublish* pub,
                  unsigned long long rng_seed, bool device_rng) {
  B200_CHECK_ARG((token_out || pub) && logits && workspace && V > 0 && V < (1ll << 31), "sample: bad arguments");
  B200_CHECK_ARG(!device_rng || (pub && pub->gen_count), "sample: the device RNG counts the engine's generated tokens");
  B200_CHECK_ARG(temperature >= 0.f && top_p >= 0.f && min_p >= 0.f && min_p <= 1.f && u >= 0.f && u <= 1.f,
                 "sample: temperature/top_p/min_p/u out of range");
  B200_CHECK_ARG((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "sample: workspace must be 256-byte aligned");
  unsigned int* hist = static_cast<unsigned int*>(workspace);
  int* kept = reinterpret_cast<int*>(hist + kBins);
  SamplePlan* plan = reinterpret_cast<SamplePlan*>(kept + kBins);
  B200_CUDA(cudaMemsetAsync(hist, 0, (size_t)kBins * 4, st));
  int grid = (int)std::min<int64_t>((V + 255) / 256, 148 * 4);
  g_launches.fetch_add(3);
  sample_hist_kernel<<<grid, 256, 0, st>>>((const __nv_bfloat16*)logits, V, hist);
  sample_plan_kernel<<<1, kPlanThreads, 0, st>>>(hist, V, temperature, top_k, top_p, min_p, kept, plan);
  sample_draw_kernel<<<1, kDrawThreads, 0, st>>>((const __nv_bfloat16*)logits, V, temperature, u, kept, plan, token_out,
                                                 pub ? *pub : ArgmaxPublish{}, rng_seed, device_rng ? 1 : 0);
  B200_CUDA(cudaGetLastError());
  return B200_OK;
}

}  // namespace b200

extern "C" {

int64_t b200_sample_workspace_bytes(void) { return b200::sample_workspace_bytes(); }

int b200_sample_bf16(int64_t* token_out, const void* logits, int64_t V, float temperature, int64_t top_k, float top_p,
                     float min_p, float u, void* workspace, void* stream) {
  int rc = b200_device_check();
  if (rc != B200_OK) return rc;
  return b200::launch_sample(token_out, logits, V, temperature, top_k, top_p, min_p, u, workspace, (cudaStream_t)stream,
                             nullptr, 0ull, false);
}

}  // extern "C"
