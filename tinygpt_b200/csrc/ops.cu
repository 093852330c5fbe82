// ops.cu — the small operators of the decode path as stand-alone launches (drop-in boundary B: one per TinyTorch op).
// In the fused engine (engine.cu) most of these run as GEMV prologues/epilogues or inside the attention kernel; the
// stand-alone versions exist for per-op parity against the oracle and for the registry adapter.
#include "common.cuh"
#include "ops.cuh"

namespace b200 {

// ------------------------------------------------------------------------------------------------------- RMSNorm
// [ref: TT/Operation/OpNNLayerCuda.cuh:252-357]  one CTA per row, fp32 Σx², rsqrtf, (x*inv)*w rounded once.
__global__ void __launch_bounds__(256) rmsnorm_kernel(__nv_bfloat16* __restrict__ y, const __nv_bfloat16* __restrict__ x,
                                                      const __nv_bfloat16* __restrict__ w, int dim, float eps) {
  __shared__ float red[8];
  pdl_trigger();
  pdl_wait();
  const size_t base = (size_t)blockIdx.x * dim;
  float ss = 0.f;
  for (int i = threadIdx.x; i < dim; i += blockDim.x) {
    const float v = bf16_to_f(x[base + i]);
    ss += v * v;
  }
  ss = warp_sum(ss);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
  __syncthreads();
  float tot = 0.f;
  for (int i = 0; i < (int)(blockDim.x >> 5); ++i) tot += red[i];
  const float inv = rsqrtf(tot / (float)dim + eps);
  for (int i = threadIdx.x; i < dim; i += blockDim.x) {
    float v = bf16_to_f(x[base + i]) * inv;
    if (w != nullptr) v *= bf16_to_f(w[i]);
    y[base + i] = f_to_bf16(v);
  }
}

// ---------------------------------------------------------------------------------------------------------- RoPE
// [ref: TT/Operation/OpNNLayerCuda.cuh:412-440]  one thread per rotated pair instead of one thread per (b,h,t).
__global__ void __launch_bounds__(256) rope_kernel(__nv_bfloat16* __restrict__ y, const __nv_bfloat16* __restrict__ x,
                                                   const float* __restrict__ table, int64_t B, int64_t heads, int64_t S,
                                                   int64_t hd, int64_t strideB, int64_t strideH, int64_t strideT,
                                                   int64_t offset) {
  pdl_trigger();
  pdl_wait();
  const int64_t half = hd >> 1;
  const int64_t total = B * heads * S * half;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = idx % half;
    int64_t r = idx / half;
    const int64_t t = r % S;
    r /= S;
    const int64_t h = r % heads;
    const int64_t b = r / heads;
    const int64_t base = b * strideB + h * strideH + t * strideT;
    const float* row = table + (offset + t) * hd * 2;
    const float c = row[i * 2], s = row[i * 2 + 1];
    const float x1 = bf16_to_f(x[base + i]);
    const float x2 = bf16_to_f(x[base + half + i]);
    y[base + i] = f_to_bf16(x1 * c - x2 * s);
    y[base + half + i] = f_to_bf16(x2 * c + x1 * s);
  }
}

// [ref: TT/Operation/OpNNLayerCuda.cuh:359-410]
__global__ void rope_inv_freq_kernel(float* inv_freq, int64_t half, float theta, float factor, float high, float low,
                                     float orig_ctx) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= half) return;
  float f = 1.f / powf(theta, (float)(idx << 1) / (float)(half << 1));
  if (factor != 0.f) {
    const float wave = 2.f * 3.14159265358979323846f / f;
    const float low_wave = orig_ctx / low;
    const float high_wave = orig_ctx / high;
    if (wave > low_wave) {
      f /= factor;
    } else if (wave < high_wave) {
      // unchanged
    } else {
      const float smooth = (orig_ctx / wave - low) / (high - low);
      const float scaled = f / factor;
      f = (1.f - smooth) * scaled + smooth * f;
    }
  }
  inv_freq[idx] = f;
}

__global__ void rope_table_kernel(const float* __restrict__ inv_freq, float* __restrict__ table, int64_t ctx, int64_t hd) {
  const int64_t pos = blockIdx.x;
  const int64_t half = hd >> 1;
  for (int64_t i = threadIdx.x; i < half; i += blockDim.x) {
    const float angle = (float)pos * inv_freq[i];
    const float c = cosf(angle), s = sinf(angle);
    const int64_t o1 = (pos * hd + i) * 2, o2 = (pos * hd + half + i) * 2;
    table[o1] = c;
    table[o1 + 1] = s;
    table[o2] = c;
    table[o2 + 1] = s;
  }
}

// ------------------------------------------------------------------------------------------------------- SiLU·mul
// [ref: TT/Operation/OpFusedCuda.cuh:15-29]
__global__ void __launch_bounds__(256) silu_mul_kernel(__nv_bfloat16* __restrict__ y,
                                                       const __nv_bfloat16* __restrict__ gu, int64_t I, int64_t n) {
  pdl_trigger();
  pdl_wait();
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = idx / I, j = idx % I;
    const float g = bf16_to_f(gu[r * 2 * I + j]);
    const __nv_bfloat16 sg = f_to_bf16(g / (1.f + expf(-g)));
    y[idx] = __hmul(sg, gu[r * 2 * I + I + j]);
  }
}

// ------------------------------------------------------------------------------------------------------------ add
__global__ void __launch_bounds__(256) add_kernel(__nv_bfloat16* y, const __nv_bfloat16* a, const __nv_bfloat16* b,
                                                  int64_t n) {  // y may alias a (prefill residual update)
  pdl_trigger();
  pdl_wait();
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (int64_t)gridDim.x * blockDim.x)
    y[idx] = __hadd(a[idx], b[idx]);
}

// ------------------------------------------------------------------------------------------------------ embedding
// [ref: TT/Operation/OpTransformCuda.cuh:108-120]  one CTA per token, 16-byte vector copy of the row.
__global__ void __launch_bounds__(128) embedding_kernel(__nv_bfloat16* __restrict__ y,
                                                        const __nv_bfloat16* __restrict__ table,
                                                        const int64_t* __restrict__ ids, int64_t V, int64_t H) {
  pdl_trigger();
  pdl_wait();
  int64_t id = ids[blockIdx.x];
  if (id < 0) id += V;  // the reference wraps negative indices
  if (id < 0 || id >= V) id = 0;
  const __nv_bfloat16* src = table + id * H;
  __nv_bfloat16* dst = y + (int64_t)blockIdx.x * H;
  if ((H & 7) == 0) {
    const uint4* s4 = reinterpret_cast<const uint4*>(src);
    uint4* d4 = reinterpret_cast<uint4*>(dst);
    for (int i = threadIdx.x; i < (int)(H >> 3); i += blockDim.x) d4[i] = s4[i];
  } else {
    for (int i = threadIdx.x; i < (int)H; i += blockDim.x) dst[i] = src[i];
  }
}

// --------------------------------------------------------------------------------------------------------- argmax
// [ref: TT/Operation/OpReduceCuda.cuh:145-156,188-224]  fp32 compare, ties → HIGHEST index.
// Pass 1: grid (chunks, rows) → (val, idx) per chunk; the last CTA of a row (atomic ticket) merges the chunks.
__device__ __forceinline__ void argmax_merge(float& v, int64_t& i, float ov, int64_t oi) {
  if (ov > v || (ov == v && oi > i)) {
    v = ov;
    i = oi;
  }
}

__global__ void __launch_bounds__(256) argmax_kernel(int64_t* __restrict__ out, const __nv_bfloat16* __restrict__ logits,
                                                     int64_t V, float* __restrict__ ws_val, int64_t* __restrict__ ws_idx,
                                                     unsigned int* __restrict__ ticket, const ArgmaxPublish pub) {
  __shared__ float sv[8];
  __shared__ int64_t si[8];
  __shared__ bool last;
  pdl_trigger();
  pdl_wait();
  const int row = blockIdx.y;
  const int chunks = gridDim.x;
  const __nv_bfloat16* lg = logits + (size_t)row * V;
  float v = -INFINITY;
  int64_t idx = -1;
  for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < V; j += (int64_t)chunks * blockDim.x)
    argmax_merge(v, idx, bf16_to_f(lg[j]), j);
  auto block_reduce = [&]() {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, v, o);
      const int64_t oi = __shfl_xor_sync(0xffffffffu, idx, o);
      argmax_merge(v, idx, ov, oi);
    }
    if ((threadIdx.x & 31) == 0) {
      sv[threadIdx.x >> 5] = v;
      si[threadIdx.x >> 5] = idx;
    }
    __syncthreads();
    if (threadIdx.x < 32) {
      v = threadIdx.x < (blockDim.x >> 5) ? sv[threadIdx.x] : -INFINITY;
      idx = threadIdx.x < (blockDim.x >> 5) ? si[threadIdx.x] : -1;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, v, o);
        const int64_t oi = __shfl_xor_sync(0xffffffffu, idx, o);
        argmax_merge(v, idx, ov, oi);
      }
    }
  };
  block_reduce();
  if (threadIdx.x == 0) {
    ws_val[(size_t)row * chunks + blockIdx.x] = v;
    ws_idx[(size_t)row * chunks + blockIdx.x] = idx;
    __threadfence();
    const unsigned int t = atomicAdd(&ticket[row], 1u);
    last = (t == (unsigned int)chunks - 1);
  }
  __syncthreads();
  if (!last) return;
  __threadfence();
  v = -INFINITY;
  idx = -1;
  for (int c = threadIdx.x; c < chunks; c += blockDim.x)
    argmax_merge(v, idx, __ldcg(&ws_val[(size_t)row * chunks + c]), __ldcg(&ws_idx[(size_t)row * chunks + c]));
  __syncthreads();
  block_reduce();
  if (threadIdx.x == 0) {
    out[row] = idx;
    ticket[row] = 0;  // self-reset for the next launch
    if (pub.tp_world > 1 && row == 0) {
      const unsigned int tag = (unsigned int)(*pub.tp_epoch + 1ull);
      const unsigned int gidx = (unsigned int)(idx + pub.tp_index_offset);
      for (int r = 0; r < pub.tp_world; ++r) {
        volatile uint2* c = pub.tp_cand[r];
        asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(c), "r"(__float_as_uint(v)), "r"(tag) : "memory");
        asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(c + 1), "r"(gidx), "r"(tag) : "memory");
      }
    } else if (pub.cur_tok != nullptr && pub.batch_rows > 1) {
      // batched engine: every row (sequence) publishes its own token; row 0 advances the shared position
      pub.cur_tok[row] = idx;
      const unsigned long long c = pub.gen_count[row];
      pub.gen_log[(c % (unsigned long long)pub.gen_cap) * (unsigned long long)pub.batch_rows + row] = idx;
      pub.gen_count[row] = c + 1;
      if (row == 0 && pub.pos != nullptr) *pub.pos += 1;
    } else if (pub.cur_tok != nullptr && row == 0) {
      // engine: the greedy token becomes the next step's input and is appended to the on-device log
      *pub.cur_tok = idx;
      if (pub.pos != nullptr) *pub.pos += 1;
      const unsigned long long c = *pub.gen_count;
      pub.gen_log[c % (unsigned long long)pub.gen_cap] = idx;
      *pub.gen_count = c + 1;
      if (pub.mailbox != nullptr) {  // one posted 8-byte write to host memory: {sequence tag, token} arrive together
        const unsigned long long word = (((c + 1ull) & 0xffffffffull) << 32) | (unsigned long long)(unsigned int)idx;
        asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(pub.mailbox + (c % pub.mailbox_cap)), "l"(word)
                     : "memory");
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------- host wrappers
static inline int grid_for(int64_t n, int block, int cap) {
  int64_t g = (n + block - 1) / block;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

int launch_rmsnorm(void* y, const void* x, const void* w, int64_t rows, int64_t dim, float eps, cudaStream_t st,
                   bool pdl) {
  B200_CHECK_ARG(y && x && rows > 0 && dim > 0 && rows < (1ll << 31) && dim < (1ll << 31),
                 "rmsnorm: bad arguments rows=%lld dim=%lld", (long long)rows, (long long)dim);
  B200_CUDA(launch_pdl(rmsnorm_kernel, dim3((unsigned)rows), dim3(256), 0, st, pdl, (__nv_bfloat16*)y,
                       (const __nv_bfloat16*)x, (const __nv_bfloat16*)w, (int)dim, eps));
  return B200_OK;
}

int launch_rope(void* y, const void* x, const float* table, int64_t B, int64_t S, int64_t heads, int64_t hd,
                int64_t offset, int layout, cudaStream_t st, bool pdl) {
  B200_CHECK_ARG(y && x && table, "rope: null pointer");
  B200_CHECK_ARG(B > 0 && S > 0 && heads > 0 && hd > 0 && (hd % 2) == 0, "rope: bad shape B=%lld S=%lld heads=%lld hd=%lld",
                 (long long)B, (long long)S, (long long)heads, (long long)hd);
  B200_CHECK_ARG(layout == B200_LAYOUT_BHSD || layout == B200_LAYOUT_BSHD, "rope: unknown layout %d", layout);
  int64_t sB, sH, sT;
  if (layout == B200_LAYOUT_BHSD) {
    sT = hd;
    sH = S * hd;
    sB = heads * S * hd;
  } else {
    sH = hd;
    sT = heads * hd;
    sB = S * heads * hd;
  }
  const int64_t total = B * heads * S * (hd / 2);
  B200_CUDA(launch_pdl(rope_kernel, dim3(grid_for(total, 256, 148 * 8)), dim3(256), 0, st, pdl, (__nv_bfloat16*)y,
                       (const __nv_bfloat16*)x, table, B, heads, S, hd, sB, sH, sT, offset));
  return B200_OK;
}

int launch_silu_mul(void* y, const void* gu, int64_t rows, int64_t I, cudaStream_t st, bool pdl) {
  B200_CHECK_ARG(y && gu && rows > 0 && I > 0, "silu_mul: bad arguments");
  const int64_t n = rows * I;
  B200_CUDA(launch_pdl(silu_mul_kernel, dim3(grid_for(n, 256, 148 * 8)), dim3(256), 0, st, pdl, (__nv_bfloat16*)y,
                       (const __nv_bfloat16*)gu, I, n));
  return B200_OK;
}

int launch_add(void* y, const void* a, const void* b, int64_t n, cudaStream_t st, bool pdl) {
  B200_CHECK_ARG(y && a && b && n > 0, "add: bad arguments");
  B200_CUDA(launch_pdl(add_kernel, dim3(grid_for(n, 256, 148 * 8)), dim3(256), 0, st, pdl, (__nv_bfloat16*)y,
                       (const __nv_bfloat16*)a, (const __nv_bfloat16*)b, n));
  return B200_OK;
}

int launch_embedding(void* y, const void* table, const int64_t* ids, int64_t n_ids, int64_t V, int64_t H,
                     cudaStream_t st, bool pdl) {
  B200_CHECK_ARG(y && table && ids && n_ids > 0 && V > 0 && H > 0 && n_ids < (1ll << 31), "embedding: bad arguments");
  B200_CUDA(launch_pdl(embedding_kernel, dim3((unsigned)n_ids), dim3(128), 0, st, pdl, (__nv_bfloat16*)y,
                       (const __nv_bfloat16*)table, ids, V, H));
  return B200_OK;
}

int argmax_chunks(int64_t V) {
  int64_t c = (V + 2047) / 2048;
  if (c > 148) c = 148;
  if (c < 1) c = 1;
  return (int)c;
}

int64_t argmax_workspace_bytes(int64_t rows, int64_t V) {
  const int64_t c = argmax_chunks(V);
  // [rows*c floats][rows*c int64][rows tickets], 16-byte aligned sections
  return ((rows * c * 4 + 15) / 16) * 16 + rows * c * 8 + ((rows * 4 + 15) / 16) * 16;
}

int launch_argmax(int64_t* idx, const void* logits, int64_t rows, int64_t V, void* workspace, cudaStream_t st,
                  bool pdl, const ArgmaxPublish* pub) {
  B200_CHECK_ARG(idx && logits && workspace && rows > 0 && V > 0 && rows < 65536, "argmax: bad arguments");
  const int c = argmax_chunks(V);
  uint8_t* ws = (uint8_t*)workspace;
  float* wv = (float*)ws;
  int64_t* wi = (int64_t*)(ws + ((rows * c * 4 + 15) / 16) * 16);
  unsigned int* ticket = (unsigned int*)((uint8_t*)wi + rows * c * 8);
  B200_CUDA(launch_pdl(argmax_kernel, dim3(c, (unsigned)rows), dim3(256), 0, st, pdl, idx,
                       (const __nv_bfloat16*)logits, V, wv, wi, ticket, pub ? *pub : ArgmaxPublish{}));
  return B200_OK;
}

}  // namespace b200

// --------------------------------------------------------------------------------------------------------- C ABI
extern "C" {

int b200_rmsnorm_bf16(void* y, const void* x, const void* w, int64_t rows, int64_t dim, float eps, void* stream) {
  return b200::launch_rmsnorm(y, x, w, rows, dim, eps, (cudaStream_t)stream, false);
}

int b200_rope_bf16(void* y, const void* x, const float* table, int64_t B, int64_t S, int64_t heads, int64_t hd,
                   int64_t pos_offset, int layout, void* stream) {
  return b200::launch_rope(y, x, table, B, S, heads, hd, pos_offset, layout, (cudaStream_t)stream, false);
}

int b200_rope_init_f32(float* table, int64_t hd, int64_t ctx, float theta, float factor, float high, float low,
                       int64_t orig_ctx, void* stream) {
  using namespace b200;
  B200_CHECK_ARG(table && hd > 0 && (hd % 2) == 0 && ctx > 0 && ctx < (1ll << 31), "rope_init: bad arguments");
  float* inv = nullptr;
  const int64_t half = hd / 2;
  B200_CUDA(cudaMallocAsync((void**)&inv, half * sizeof(float), (cudaStream_t)stream));
  g_launches.fetch_add(2);
  rope_inv_freq_kernel<<<(unsigned)((half + 127) / 128), 128, 0, (cudaStream_t)stream>>>(inv, half, theta, factor, high,
                                                                                        low, (float)orig_ctx);
  rope_table_kernel<<<(unsigned)ctx, 64, 0, (cudaStream_t)stream>>>(inv, table, ctx, hd);
  B200_CUDA(cudaGetLastError());
  B200_CUDA(cudaFreeAsync(inv, (cudaStream_t)stream));
  return B200_OK;
}

int b200_silu_mul_bf16(void* y, const void* gate_up, int64_t rows, int64_t I, void* stream) {
  return b200::launch_silu_mul(y, gate_up, rows, I, (cudaStream_t)stream, false);
}

int b200_add_bf16(void* y, const void* a, const void* b, int64_t n, void* stream) {
  return b200::launch_add(y, a, b, n, (cudaStream_t)stream, false);
}

int b200_embedding_bf16(void* y, const void* table, const int64_t* ids, int64_t n_ids, int64_t V, int64_t H,
                        void* stream) {
  return b200::launch_embedding(y, table, ids, n_ids, V, H, (cudaStream_t)stream, false);
}

int64_t b200_argmax_workspace_bytes(int64_t rows, int64_t V) { return b200::argmax_workspace_bytes(rows, V); }

int b200_argmax_bf16(int64_t* idx, const void* logits, int64_t rows, int64_t V, void* workspace, void* stream) {
  return b200::launch_argmax(idx, logits, rows, V, workspace, (cudaStream_t)stream, false, nullptr);
}

}  // extern "C"
