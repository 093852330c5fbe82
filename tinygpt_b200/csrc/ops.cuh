// ops.cuh — host launchers of the small decode-path operators (ops.cu) and of attention (attn.cu).
#pragma once
#include "common.cuh"

namespace b200 {

int launch_rmsnorm(void* y, const void* x, const void* w, int64_t rows, int64_t dim, float eps, cudaStream_t st,
                   bool pdl);
int launch_rope(void* y, const void* x, const float* table, int64_t B, int64_t S, int64_t heads, int64_t hd,
                int64_t offset, int layout, cudaStream_t st, bool pdl);
int launch_silu_mul(void* y, const void* gu, int64_t rows, int64_t I, cudaStream_t st, bool pdl);
int launch_add(void* y, const void* a, const void* b, int64_t n, cudaStream_t st, bool pdl);
int launch_embedding(void* y, const void* table, const int64_t* ids, int64_t n_ids, int64_t V, int64_t H,
                     cudaStream_t st, bool pdl);
int argmax_chunks(int64_t V);
int64_t argmax_workspace_bytes(int64_t rows, int64_t V);
// Optional tail of the engine's argmax: the last CTA publishes the greedy token as the next step's input.
struct ArgmaxPublish {
  // rows > 1 (batched decode): row r publishes into cur_tok[r], gen_log[r · gen_cap + …], gen_count[r]; row 0 alone
  // advances the (shared) position and posts to the mailbox
  int* pos = nullptr;  // position counter advanced for the next token
  int64_t* cur_tok = nullptr;
  int64_t* gen_log = nullptr;
  unsigned long long* gen_count = nullptr;
  int gen_cap = 1;
  int batch_rows = 0;   // > 1: per-row publication (see above); the log is [gen_cap][batch_rows]
  // tensor parallel (vocabulary-sharded lm_head): instead of publishing, push (max logit, GLOBAL index) into every
  // rank's candidate slot and bump its arrival counter; tp_finish_kernel (engine.cu) picks the winner.
  // async token pipeline: ring in pinned, device-mapped host memory (b200_engine_set_mailbox); null = none
  unsigned long long* mailbox = nullptr;
  unsigned long long mailbox_cap = 1;
  int tp_world = 1;
  int64_t tp_index_offset = 0;
  uint2* tp_cand[8] = {nullptr};                 // 2 words per rank: {value bits, tag}, {global index, tag}
  const unsigned long long* tp_epoch = nullptr;   // tag = *tp_epoch + 1
};
int launch_argmax(int64_t* idx, const void* logits, int64_t rows, int64_t V, void* workspace, cudaStream_t st,
                  bool pdl, const ArgmaxPublish* pub);

// ---- device sampler (sampling.cu): temperature / top-k / top-p / min-p + inverse-CDF draw without sorting
int64_t sample_workspace_bytes();
int launch_sample(int64_t* token_out, const void* logits, int64_t V, float temperature, int64_t top_k, float top_p,
                  float min_p, float u, void* workspace, cudaStream_t st, const ArgmaxPublish* pub,
                  unsigned long long rng_seed, bool device_rng);

// ---- tcgen05 prefill GEMM (gemm.cu)
int launch_gemm_bf16(void* C, const void* A, const void* B, int64_t M, int64_t N, int64_t K, cudaStream_t st);

// ---- prefill helpers (prefill.cu)
int launch_bias_add(void* y, const void* b, int64_t rows, int64_t n, cudaStream_t st);
int launch_prefill_qk(void* qkv, const void* q_norm, const void* k_norm, float eps, const float* rope, void* kcache,
                      void* vcache, int S, int Hq, int Hkv, int hd, int p0, cudaStream_t st);
int launch_attn_prefill(void* o, const void* qkv, const void* kcache, const void* vcache, int S, int Hq, int Hkv, int hd,
                        int p0, cudaStream_t st);

// ---- tensor-core prefill attention (prefill_attn.cu), opt-in: B200_PREFILL_ATTN=mma
bool prefill_attn_mma_enabled();
int launch_attn_prefill_mma(void* o, const void* qkv, const void* kcache, const void* vcache, int S, int Hq, int Hkv, int hd,
                            int p0, cudaStream_t st);
int launch_attn_causal_mma(void* o, const void* q, const void* k, const void* v, int64_t B, int64_t S, int64_t Hq,
                           int64_t Hkv, int64_t hd, cudaStream_t st);

// ---- attention (attn.cu)
// Fused decode attention of one layer for one new token (B = 1, Sq = 1):
//   q,k,v = split(qkv)  →  optional per-head RMSNorm on q,k (Qwen3)  →  RoPE(q), RoPE(k) at position *pos  →
//   K/V appended in place to the cache  →  split-KV softmax(q kᵀ/√hd) v over positions 0…*pos  →  o (bf16).
struct AttnDecodeParams {
  const __nv_bfloat16* qkv;      // [Hq*hd + 2*Hkv*hd] projections (bias already applied)
  const __nv_bfloat16* q_norm;   // [hd] or null
  const __nv_bfloat16* k_norm;   // [hd] or null
  float eps;
  const float* rope;             // [max_ctx, hd, 2] fp32 table, or null (no rotation: stand-alone attention op)
  const int* pos;                // device scalar: past length (index of the new token), stable since BEFORE the
                                 // producer kernel started (see engine.cu); null ⇒ fixed_len - 1
  int fixed_len;                 // used when pos == null: number of keys, nothing is appended
  __nv_bfloat16* kcache;         // [max_ctx, Hkv, hd]
  __nv_bfloat16* vcache;         // [max_ctx, Hkv, hd]
  __nv_bfloat16* out;            // [Hq*hd]
  float* ws;                     // split partials: [Hq / G][nsplit][G][hd + 2] floats (G = query heads per CTA)
  unsigned int* tickets;         // [Hq], zero-initialised, self-resetting
  int heads_per_cta;             // 0 = choose from max_ctx (attn_heads_per_cta)
  unsigned long long* trace;     // debug timestamps (see GemvParams::trace)
  int Hq, Hkv, nsplit, max_ctx;  // nsplit = attn_decode_nsplit(hd, max_ctx): fixed 256 (hd 64) / 128 (hd 128) keys per split
  // batched decode: grid.z = batch sequences at the SAME position (the reference's left-padded batch,
  // src/engine/GPTEngine.cpp:101-144); sequence b's buffers are b·stride elements further on
  int batch;                     // 0 / 1: one sequence
  long long qkv_bstride, out_bstride, cache_bstride, ws_bstride;
  int tick_bstride;
};
int attn_decode_nsplit(int hd, int max_ctx);
int attn_heads_per_cta(int Hq, int Hkv, int max_ctx);
int launch_attn_decode(const AttnDecodeParams& p, int hd, cudaStream_t st, bool pdl);
int64_t attn_decode_ws_floats(int Hq, int Hkv, int hd, int nsplit);
int attn_setup_attributes();

// General (any Sq, causal or not) attention — correctness path behind b200_attn_bf16 for Sq > 1.
int launch_attn_general(void* o, const void* q, const void* k, const void* v, int64_t B, int64_t Sq, int64_t Skv,
                        int64_t Hq, int64_t Hkv, int64_t hd, int causal, cudaStream_t st);

}  // namespace b200
