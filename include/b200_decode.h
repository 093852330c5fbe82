/*
 * b200_decode.h — C ABI of the B200-native decode engine that sits behind TinyGPT's
 * GPTEngine::generate*() loop and TinyTorch's operator registry.
 *
 * Conventions (see INTEGRATION.md for the TinyTorch-side adapter):
 *   - every pointer is a DEVICE pointer unless its name ends in _host;
 *   - bf16 is raw 16-bit storage (same bits as tinytorch::BFloat16 / __nv_bfloat16), row-major contiguous;
 *   - the caller passes its own cudaStream_t (TinyTorch: cuda::getCurrentCUDAStream(dev).stream(),
 *     third_party/TinyTorch/src/Utils/CUDAUtils.cpp:123-135) as an opaque void*;
 *   - every function returns 0 on success and a negative b200_status on failure, never throws,
 *     never synchronises the device, and (op level) never allocates; the message of the last failure on the calling
 *     thread is available from b200_last_error().  The reference's own convention is LOGE + ASSERT
 *     (third_party/TinyTorch/src/Utils/Macros.h:34-40); the adapter maps non-zero to that.
 *   - there is NO CPU fallback: on a machine without an sm_100 device every compute entry point fails with
 *     B200_ERR_NO_DEVICE.
 *
 * Reference interface each entry point replaces is cited as  [ref: file:line]  with paths relative to the
 * keith2018/TinyGPT checkout (TT/ = third_party/TinyTorch/src/).
 */
#ifndef B200_DECODE_H_
#define B200_DECODE_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default) /* the library is built with -fvisibility=hidden; these are its only exports */
#endif

/* 2: + b200_engine_set_mailbox, b200_engine_generated;  3: + b200_sample_bf16, b200_sample_workspace_bytes,
 * b200_engine_set_sampler, b200_engine_options.  All additive: callers of an older version keep working. */
#define B200_ABI_VERSION 3

typedef enum b200_status {
  B200_OK = 0,
  B200_ERR_INVALID = -1,     /* bad argument (shape, alignment, null pointer) */
  B200_ERR_UNSUPPORTED = -2, /* valid in the reference, not built here (message says what) */
  B200_ERR_CUDA = -3,        /* a CUDA runtime/driver call failed */
  B200_ERR_NO_DEVICE = -4,   /* no sm_100 device visible */
  B200_ERR_STATE = -5,       /* engine used out of order (e.g. context overflow) */
  B200_ERR_NCCL = -6
} b200_status;

/* QKV layout selector of ropeApply. [ref: TT/Operation/OpNNLayer.h:13-26 enum QKVLayout] */
typedef enum b200_qkv_layout { B200_LAYOUT_BHSD = 0, B200_LAYOUT_BSHD = 1 } b200_qkv_layout;

int b200_abi_version(void);
const char* b200_last_error(void);
/* Number of kernels this library has launched on this process since load (bench.py's gpu_launches claim). */
int64_t b200_launch_count(void);
/* 0 when a compute-capability 10.x device is current, else B200_ERR_NO_DEVICE. */
int b200_device_check(void);

/* ------------------------------------------------------------------------------------------------------------------
 * Boundary B — one entry point per TinyTorch op on the path.
 * ---------------------------------------------------------------------------------------------------------------- */

/* y[m,n] = bf16( x[m,k] · W[n,k]^T ) (fp32 accumulate, one rounding); if bias != NULL:
 * y = bf16( float(y) + float(bias[n]) )  — the reference's second rounding for 3-D Linear inputs.
 * [ref: TT/Operation/OpLinalg.cpp:244-277 matmul + addInplace; TT/Operation/OpLinalgCuda.cuh:276-293
 *  gemmStridedBatchedCudaBF16Impl; caller TT/Function/FuncNNLayer.h:14-18 FuncLinear]
 * Requirements: k % 8 == 0, W 16-byte aligned.  m is small (decode: 1); W is streamed once per 8 rows of x. */
int b200_gemv_bf16(void* y, const void* x, const void* W, const void* bias_or_null, int64_t m, int64_t n, int64_t k,
                   void* stream);

/* Batched (prefill) GEMM on the tcgen05 tensor cores: C[M,N] = bf16( A[M,K] · B[N,K]^T ), fp32 accumulation in TMEM,
 * one rounding.  The m > 1 side of op::matmul for Linear layers (bias / residual are separate ops, as in the reference).
 * [ref: TT/Operation/OpLinalg.cpp:244-277; TT/Operation/OpLinalgCuda.cuh:193-216,276-293]
 * Requirements: K % 8 == 0, N % 8 == 0, 16-byte aligned operands. */
int b200_gemm_bf16(void* C, const void* A, const void* B, int64_t M, int64_t N, int64_t K, void* stream);

/* y[r,:] = bf16( float(x[r,:]) * rsqrtf(mean(x[r,:]^2) + eps) * float(w[:]) )  (single rounding, fp32 weight multiply)
 * [ref: TT/Operation/OpNNLayerCuda.cuh:252-357 kNormSmall/kNormLarge<RMSNorm>, host :569-619] */
int b200_rmsnorm_bf16(void* y, const void* x, const void* w_or_null, int64_t rows, int64_t dim, float eps,
                      void* stream);

/* rotate-half RoPE with the reference's fp32 table [ctx, hd, 2] (cos,sin interleaved):
 * y[i] = bf16(x1*c - x2*s), y[i+hd/2] = bf16(x2*c + x1*s), position = pos_offset + t.
 * [ref: TT/Operation/OpNNLayerCuda.cuh:412-440 kRopeApply, host :658-708] */
int b200_rope_bf16(void* y, const void* x, const float* table_f32, int64_t B, int64_t S, int64_t heads, int64_t hd,
                   int64_t pos_offset, int layout /* b200_qkv_layout */, void* stream);

/* Build the fp32 cos/sin table the way the reference does (device kernels, fp32 powf/cosf/sinf).
 * scaling_factor == 0 disables llama3 scaling.  table_f32 has ctx*hd*2 floats.  Init-time only: uses a stream-ordered
 * temporary (cudaMallocAsync).
 * [ref: TT/Operation/OpNNLayerCuda.cuh:359-410 kRopeComputeInvFreq/kRopeApplyScaling/kRopePrecomputeCosSin,
 *  host :621-656 ropeInitOpCudaImpl] */
int b200_rope_init_f32(float* table_f32, int64_t hd, int64_t ctx, float theta, float scaling_factor,
                       float high_freq_factor, float low_freq_factor, int64_t original_ctx, void* stream);

/* o[B,Sq,Hq,hd] = softmax(q k^T / sqrt(hd)) v, BSHD, GQA (kv head = h / (Hq/Hkv)), causal mask top-left aligned
 * (col > row masked) exactly as TinyFA.  hd in {64, 128}.
 * [ref: TT/Operation/OpNNLayerCuda.cu:12-44 flashAttentionOpCudaImpl → TFA/flash_api.cuh:43-50 tfa::flashAttn →
 *  TFA/mma/kernel.cuh:18-203] */
int b200_attn_bf16(void* o, const void* q, const void* k, const void* v, int64_t B, int64_t Sq, int64_t Skv,
                   int64_t Hq, int64_t Hkv, int64_t hd, int causal, void* stream);

/* y[r,j] = bf16( bf16(silu(g[r,j])) * u[r,j] ), gate_up rows are [gate(I) | up(I)].
 * [ref: TT/Operation/OpFusedCuda.cuh:15-48 kSiluMul; OpElemWiseCuda.cuh:124-131 OpCudaSilu] */
int b200_silu_mul_bf16(void* y, const void* gate_up, int64_t rows, int64_t I, void* stream);

/* y = bf16(a + b) elementwise (alpha = 1). [ref: TT/Operation/OpElemWiseCuda.cuh:133-144,371-381] */
int b200_add_bf16(void* y, const void* a, const void* b, int64_t n, void* stream);

/* y[t,:] = table[ids[t],:] ; ids int64. [ref: TT/Operation/OpTransformCuda.cuh:108-120 kIndex, host :491-526] */
int b200_embedding_bf16(void* y, const void* table, const int64_t* ids, int64_t n_ids, int64_t V, int64_t H,
                        void* stream);

/* idx[r] = argmax_j logits[r,j], compared in fp32, ties resolved to the HIGHEST index (reference CUDA rule).
 * [ref: TT/Operation/OpReduceCuda.cuh:145-156 cudaWarpReduceIdx, :188-224 kReduceIdxMerge, :459-493, :581-636]
 * `workspace` must hold b200_argmax_workspace_bytes(rows, V) bytes, ZERO-INITIALISED once by the caller. */
int64_t b200_argmax_workspace_bytes(int64_t rows, int64_t V);
int b200_argmax_bf16(int64_t* idx, const void* logits, int64_t rows, int64_t V, void* workspace, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Fused building blocks of the engine, exported so that parity tests can drive the exact kernels the engine runs.
 * ---------------------------------------------------------------------------------------------------------------- */

/* One launch = [optional RMSNorm(x, norm_w, eps)] → GEMV → one of
 *   bias      y = bf16(bf16(acc) + bias)                    [ref: TT/Operation/OpLinalg.cpp:273-275]
 *   residual  y = bf16(residual + bf16(acc))                [ref: src/layer/DecoderLayer.h:40-41]
 *   silu_mul  W rows are [gate(n) | up(n)] (nseg = 2):  y = bf16(bf16(silu(bf16(acc_g))) * bf16(acc_u))
 *                                                           [ref: src/layer/GatedMLP.h:37-41]
 * m = 1.  y may alias residual. */
int b200_gemv_fused_bf16(void* y, const void* x, const void* W, int64_t n, int64_t k, int nseg,
                         const void* norm_w_or_null, float eps, const void* bias_or_null,
                         const void* residual_or_null, int silu_mul, void* stream);

/* Decode attention of one layer for one new token (B = 1, Sq = 1), one launch:
 *   q|k|v = qkv  →  [q_norm/k_norm per head]  →  RoPE(q), RoPE(k) at position *pos  →  K/V row *pos written in place
 *   into kcache/vcache [max_ctx, Hkv, hd]  →  split-KV softmax(q kᵀ/√hd) v over rows 0…*pos  →  out [Hq*hd].
 * The context is cut into splits of 256 (hd 64) / 128 (hd 128) keys, one CTA per (KV head, split).
 * *pos must have been written before the kernel that produced `qkv` was launched (the kernel reads it, and the cached
 * rows below it, ahead of its programmatic-dependency wait).
 * pos == NULL: plain attention over `fixed_len` cached rows, nothing appended, no rotation (rope_table may be NULL).
 * workspace: b200_attn_decode_workspace_bytes() bytes, ZERO-INITIALISED once by the caller (tickets self-reset).
 * [ref: src/layer/Attention.h:71-112,156-163; src/engine/CacheManager.h:24-42; TFA/mma/kernel.cuh:18-203] */
int64_t b200_attn_decode_workspace_bytes(int64_t Hq, int64_t Hkv, int64_t hd, int64_t max_ctx);
int b200_attn_decode_bf16(void* out, const void* qkv, const void* q_norm_or_null, const void* k_norm_or_null,
                          float eps, const float* rope_table, const int32_t* pos_or_null, int64_t fixed_len,
                          void* kcache, void* vcache, int64_t Hq, int64_t Hkv, int64_t hd, int64_t max_ctx,
                          void* workspace, void* stream);

/* Sampler::sample for one row of bf16 logits [V] on the device (temperature → top-k → top-p → min-p → softmax →
 * inverse-CDF draw with the caller's uniform number u in (0, 1]), without sorting the vocabulary (sampling.cu).
 * [ref: src/engine/Sampler.cpp:31-78; third_party/TinyTorch/src/Operation/OpSamplingCuda.cu:30-62,97-170,261-330]
 * Greedy decoding (temperature 0, top_k 0, top_p 1, min_p 0) is b200_argmax_bf16.  `workspace`: 256-byte aligned,
 * b200_sample_workspace_bytes() bytes. */
int64_t b200_sample_workspace_bytes(void);
int b200_sample_bf16(int64_t* token_out, const void* logits, int64_t V, float temperature, int64_t top_k, float top_p,
                     float min_p, float u, void* workspace, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Boundary A — the whole per-token forward behind GPTModel::forward / GPTEngine::genNextToken.
 * [ref: src/model/GPTModel.h:51-58 CausalLM::forward, :80-106 GPTModel; src/layer/DecoderLayer.h:38-43;
 *  src/layer/Attention.h:71-112,156-163; src/layer/GatedMLP.h:37-41; src/engine/CacheManager.h:13-55;
 *  src/engine/GPTEngine.cpp:94-99,154-174; src/engine/Sampler.cpp:23-29 greedy branch]
 * ---------------------------------------------------------------------------------------------------------------- */

typedef struct b200_model_desc {
  int32_t hidden;         /* H */
  int32_t layers;         /* L */
  int32_t q_heads;        /* Hq  (global, before tensor-parallel sharding) */
  int32_t kv_heads;       /* Hkv (global) */
  int32_t head_dim;       /* hd: 64 or 128 */
  int32_t intermediate;   /* I   (global) */
  int32_t vocab;          /* V */
  int32_t max_ctx;        /* KV-cache capacity in tokens; also the number of rows of rope_table */
  float rms_eps;
  int32_t qkv_bias;       /* Qwen2: 1 [ref: src/model/ModelQwen2.h:26-31] */
  int32_t qk_norm;        /* Qwen3: per-head RMSNorm on q and k before RoPE [ref: src/layer/Attention.h:156-163] */
  int32_t tp_rank;        /* tensor-parallel rank / world; world == 1 ⇒ single GPU */
  int32_t tp_world;
  int32_t tp_shard_attn;  /* 1: heads sharded over ranks; 0: attention replicated (Hkv % world != 0) */
} b200_model_desc;

/* Per-layer weights, all bf16, exactly the reference's state layout:
 *   qkv_w rows = [q(qDim) | k(kvDim) | v(kvDim)] × H   [ref: src/layer/Linear.h:64-79 MergedLinear views]
 *   gate_up_w rows = [gate(I) | up(I)] × H               [ref: src/layer/GatedMLP.h:19-21]
 * For tensor-parallel engines the pointers are this rank's shard (see tinygpt_b200/tp.py for the slicing rule). */
typedef struct b200_layer_weights {
  const void* input_norm; /* [H] */
  const void* qkv_w;      /* [qDim + 2 kvDim, H] */
  const void* qkv_b;      /* [qDim + 2 kvDim] or NULL */
  const void* q_norm;     /* [hd] or NULL */
  const void* k_norm;     /* [hd] or NULL */
  const void* o_w;        /* [H, qDim] */
  const void* post_norm;  /* [H] */
  const void* gate_up_w;  /* [2 I, H] */
  const void* down_w;     /* [H, I] */
} b200_layer_weights;

typedef struct b200_weight_table {
  const void* embed;        /* [V, H] */
  const void* final_norm;   /* [H] */
  const void* lm_head;      /* [V, H] (== embed when tie_word_embeddings) */
  const float* rope_table;  /* [max_ctx, hd, 2] fp32, the reference's RoPE::cache() tensor */
  const b200_layer_weights* layers_host; /* HOST array of `layers` entries holding device pointers */
} b200_weight_table;

typedef struct b200_engine b200_engine;

/* Borrow the weights (no copy), allocate the in-place KV cache [L][2][max_ctx][Hkv][hd] + workspace once,
 * encode the TMA descriptors, capture the per-token CUDA graphs.  Synchronises the device once (weights may still
 * be in flight on the loader thread's stream, see SURVEY.md §8b "Threading"). */
int b200_engine_create(const b200_model_desc* desc, const b200_weight_table* weights, b200_engine** out);
void b200_engine_destroy(b200_engine* eng);

/* New sequence: position ← 0 (the KV cache is overwritten in place). [ref: GPTModel::resetCache, GPTModel.h:91-94] */
int b200_engine_reset(b200_engine* eng, void* stream);

/* Rewind to `position` (0 <= position <= current): the cached K/V rows below it stay valid, later rows are
 * overwritten as decoding continues.  Used to re-decode from a shared prefix (and by bench.py between timed steps). */
int b200_engine_seek(b200_engine* eng, int64_t position, void* stream);

/* model()(ids): run S tokens of B sequences (ids [B, S] row-major, 1 ≤ B ≤ 8, all at the engine's current position — the
 * reference's left-padded batch, src/engine/GPTEngine.cpp:101-174) and return logits.
 *   logits_mode 0: logits_out is [B, V]      — last position only
 *   logits_mode 1: logits_out is [B, S, V]   — every position, like the reference's lm_head over all S
 * B > 1 (single-GPU, greedy engines): every sequence has its own KV cache; a decode step streams the weights ONCE for the
 * whole batch (gemv_batch.cu: tensor cores, the B activation vectors as the n = 8 MMA operand) and agrees with B independent
 * batch-1 steps to summation-order noise, like the reference's own m = B GEMM against its m = 1 path.  The first call with a new B builds
 * the batch's buffers and graphs.  b200_engine_decode / b200_engine_last_token follow the batch of the latest forward.
 * After the call the engine's "current token" is the greedy argmax of the last position (reference tie rule), so
 * b200_engine_decode can continue without a host round trip.  logits_out may be NULL. */
int b200_engine_forward(b200_engine* eng, const int64_t* ids, int64_t B, int64_t S, void* logits_out,
                        int logits_mode, void* stream);

/* The generateSync hot loop on device: n_steps × { forward(current token) → greedy argmax → becomes current }.
 * tokens_out[i] (int64) receives the token produced by step i ([n_steps][B] after a batched forward).  No host
 * synchronisation. */
int b200_engine_decode(b200_engine* eng, int64_t n_steps, int64_t* tokens_out, void* stream);

/* Greedy token(s) chosen after the most recent forward/decode step (int64 [B], device → device copy on `stream`). */
int b200_engine_last_token(b200_engine* eng, int64_t* token_out, void* stream);

/* Async token pipeline [ref: src/engine/GPTEngine.cpp:17-35 AsyncTokenPipeline / DefaultTokenPipeline::fetchTokenId —
 * a blocking Tensor::item() per token — and :180-232 generateAsync].  Here the kernel that picks the greedy token also
 * posts it into a ring in PINNED, device-mapped HOST memory, as ONE 8-byte word
 *     ((n & 0xffffffff) << 32) | (uint32)token        n = 1-based count of tokens this engine has generated
 * at ring[(n - 1) % capacity], so the host reads tokens by polling memory: no stream synchronisation, no memcpy, and
 * the engine can run several steps ahead of the consumer (EOS / abort: stop launching, b200_engine_seek back).
 * `ring_host_mapped` must be 8-byte aligned pinned memory that the device can address (cudaHostAlloc /
 * cudaHostRegister; with unified addressing the host pointer itself), zero-initialised, alive until the mailbox is
 * cleared with ring = NULL or the engine destroyed.  Re-captures the per-token graphs; single-GPU engines only. */
int b200_engine_set_mailbox(b200_engine* eng, uint64_t* ring_host_mapped, int64_t capacity, void* stream);
/* SamplerConfig of the generate loop [ref: src/engine/Sampler.h:13-22, Sampler.cpp:14-21]: with any knob set
 * (temperature > 0, top_k > 0, top_p < 1, min_p > 0) the engine picks every token with the device sampler
 * (b200_sample_bf16's kernels, uniform number = Philox4x32-10(seed, tokens generated so far)); all knobs off = greedy
 * argmax again.  Re-captures the per-token graph; single-GPU engines only. */
int b200_engine_set_sampler(b200_engine* eng, float temperature, int64_t top_k, float top_p, float min_p, uint64_t seed,
                            void* stream);
/* Host-side mirror: how many tokens this engine has generated (= n of the most recently enqueued token). */
int64_t b200_engine_generated(const b200_engine* eng);

/* Debug (engine created with B200_TRACE=1 in the environment): copies [launch][8] %globaltimer stamps (entry, after
 * the PDL wait, exit, x ready, first stage landed, first row block summed, first row block stored, unused) of the GEMV / attention launches of the most recent token into out_host; synchronises.
 * Returns the number of launches copied (0 when tracing is off). */
int64_t b200_engine_debug_trace(b200_engine* eng, uint64_t* out_host, int64_t max_entries);

/* Host-side mirrors (no device access): tokens consumed so far, and kernels per decoded token (of the batched graph
 * while the engine holds a batch: the same nodes, plus one per layer when the down projection runs in two groups). */
int64_t b200_engine_position(const b200_engine* eng);
int64_t b200_engine_launches_per_token(const b200_engine* eng);
/* Which code paths this engine was built with (environment / defaults at create time): bit 0 CUDA graph, bit 1
 * programmatic dependent launch, bit 3 batched GEMM prefill (bit 2 and bits 8-15 belonged to variants that were measured
 * and removed: profiles/experiments/). */
int64_t b200_engine_options(const b200_engine* eng);
/* Algorithmic HBM bytes one decode step reads on THIS rank at context length ctx (weights once + KV + logits). */
int64_t b200_engine_bytes_per_token(const b200_engine* eng, int64_t ctx);

/* ---- tensor parallel (one process per GPU; the caller exchanges the handles with torch.distributed) ---------- */

#define B200_IPC_HANDLE_BYTES 64
/* Size in bytes of the per-rank exchange window the engine wants mapped on every peer. */
int64_t b200_tp_window_bytes(const b200_model_desc* desc);
/* Allocate this rank's window and export its IPC handle. */
int b200_tp_window_create(int64_t bytes, void** window_out, uint8_t handle_out[B200_IPC_HANDLE_BYTES]);
/* Map a peer's window. */
int b200_tp_window_open(const uint8_t handle[B200_IPC_HANDLE_BYTES], void** peer_window_out);
int b200_tp_window_close(void* peer_window);
int b200_tp_window_destroy(void* window);
/* Like b200_engine_create, for rank desc->tp_rank of desc->tp_world: `windows_host[r]` is rank r's window as mapped
 * in this process (windows_host[tp_rank] is the local one).  The hidden-vector reductions after o_proj and
 * down_proj are done by the GEMV epilogue pushing fp32 partials into every peer's window over NVLink and the next
 * kernel's prologue summing them (no separate collective kernel). */
int b200_engine_create_tp(const b200_model_desc* desc, const b200_weight_table* weights, void* const* windows_host,
                          b200_engine** out);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* B200_DECODE_H_ */
