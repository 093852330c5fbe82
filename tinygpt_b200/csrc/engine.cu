// engine.cu — the whole per-token forward behind GPTModel::forward / GPTEngine::genNextToken as ONE CUDA graph.
//   [ref: src/model/GPTModel.h:51-58 CausalLM::forward; src/layer/DecoderLayer.h:38-43; src/layer/Attention.h:71-112;
//    src/layer/GatedMLP.h:37-41; src/engine/CacheManager.h:13-55; src/engine/GPTEngine.cpp:94-99,154-174;
//    src/engine/Sampler.cpp:23-29]
//
// Reference: ≈20 launches/memcpys per layer + host-side allocator traffic per op, O(ctx) KV re-copy per step.
// Here, per layer (5 launches, PDL-chained so weight streaming of launch i+1 overlaps the tail of launch i):
//   1. qkv   = GEMV[RMSNorm(x) prologue, +bias epilogue]
//   2. attn  = q/k-norm + RoPE + in-place KV append + split-KV attention + merge          (attn.cu)
//   3. x     = GEMV[o_proj, residual epilogue]
//   4. act   = GEMV[RMSNorm(x) prologue, merged gate|up, SiLU·mul epilogue]
//   5. x     = GEMV[down_proj, residual epilogue]
// plus embed (token gather + position bookkeeping), lm_head GEMV[final RMSNorm prologue] and argmax; the greedy token
// feeds the next step on device, so the generate loop has no host synchronisation.
//
// HBM layout (all allocated once at create): KV cache [L][2][max_ctx][Hkv][hd] bf16; activations (hidden, qkv, attn,
// act, logits) a few KB each and L2-resident; split-KV workspace; device scalars (position, token, counters).
#include <cstring>
#include <new>
#include <vector>

#include "gemv.cuh"
#include "ops.cuh"

struct b200_engine {
  b200_model_desc d{};
  int num_sms = 148;
  int qdim = 0, kvdim = 0, Hq_l = 0, Hkv_l = 0, I_l = 0, V_l = 0;  // local (per-rank) sizes
  // borrowed weights
  const __nv_bfloat16* embed = nullptr;
  const float* rope = nullptr;
  std::vector<b200_layer_weights> lw;
  // owned device memory
  uint8_t* arena = nullptr;
  size_t arena_bytes = 0;
  __nv_bfloat16 *kcache = nullptr, *vcache = nullptr;  // [L][max_ctx][Hkv_l][hd] each
  __nv_bfloat16 *x = nullptr, *qkv = nullptr, *attn = nullptr, *act = nullptr, *logits = nullptr;
  float* attn_ws = nullptr;
  unsigned int* attn_tickets = nullptr;
  void* argmax_ws = nullptr;
  int64_t* cur_tok = nullptr;   // token the next step consumes
  int64_t* gen_log = nullptr;   // ring of generated tokens
  int* pos = nullptr;           // position of the token in flight; advanced by the LAST kernel of each token, so it
                                // is stable (and readable ahead of griddepcontrol.wait) for the whole next token
  unsigned long long* gen_count = nullptr;
  int gen_cap = 0;
  int nsplit = 1;
  // plans
  std::vector<b200::GemvPlan> p_qkv, p_o, p_gu, p_down;
  b200::GemvPlan p_head{};
  // graphs
  cudaGraphExec_t g_step = nullptr;   // full token incl. lm_head + argmax
  cudaGraphExec_t g_body = nullptr;   // embed + layers only (prefill tokens whose logits nobody reads)
  // host mirrors
  int64_t h_pos = 0;
  int64_t h_gen = 0;
  int launches_per_token = 0;   // kernels of a head token (what the graph g_step launches)
  int launches_body = 0;        // kernels of a body-only token (g_body)
  // batched prefill (tcgen05 GEMM path), workspace allocated on first use
  uint8_t* pf_arena = nullptr;
  int pf_chunk = 0;  // tokens the workspace holds
  __nv_bfloat16 *pf_x = nullptr, *pf_h = nullptr, *pf_qkv = nullptr, *pf_attn = nullptr, *pf_gu = nullptr,
                *pf_act = nullptr, *pf_t = nullptr;
  bool use_prefill_gemm = true;
  const void* final_norm_w = nullptr;
  // tensor parallel
  int tp_world = 1, tp_rank = 0;
  bool shard_attn = true;
  uint8_t* win[b200::kMaxTpWorld] = {nullptr};  // exchange windows of every rank as mapped here
  unsigned long long* tp_epoch = nullptr;          // tokens completed (local)
  __nv_bfloat16* x_alt = nullptr;                  // second hidden-state buffer (TP ping-pong)
  const __nv_bfloat16* x_head = nullptr;           // residual the lm_head prologue reads (TP)
  // ---- batched decode (B ≤ 8 sequences at the same position, the reference's left-padded batch): own KV caches and
  // activations for all B sequences, GEMV plans that stream W once for the batch (gemv_batch.cu), own graphs.  Built on
  // the first forward with B > 1 (engine_batch_prepare).
  int batch = 1;                       // sequences of the most recent forward (decode / last_token follow it)
  int b_cap = 0;                       // batch the buffers / plans / graphs below were built for
  uint8_t* b_arena = nullptr;
  __nv_bfloat16 *bk = nullptr, *bv = nullptr;                 // [B][L][max_ctx][Hkv][hd]
  __nv_bfloat16 *bx = nullptr, *bqkv = nullptr, *battn = nullptr, *bact = nullptr, *blogits = nullptr;   // [B][·]
  float* b_attn_ws = nullptr;
  unsigned int* b_tickets = nullptr;
  void* b_argmax_ws = nullptr;
  int64_t* b_tok = nullptr;            // [B] token every sequence consumes next
  int64_t* b_log = nullptr;            // [gen_cap][B] generated tokens
  unsigned long long* b_cnt = nullptr; // [B] tokens generated per sequence (all equal)
  int64_t hb_gen = 0;                  // host mirror of b_cnt
  std::vector<b200::GemvPlan> bp_qkv, bp_o, bp_gu, bp_down;
  b200::GemvPlan bp_head{};
  cudaGraphExec_t gb_step = nullptr, gb_body = nullptr;
  int b_launches = 0, b_launches_body = 0;
  unsigned long long* trace = nullptr;  // B200_TRACE=1: [launch][8] globaltimer stamps of the last token
  bool use_graph = true;
  bool use_pdl = true;
  // sampler (b200_engine_set_sampler): when active the last kernels of a head token are the device sampler instead of
  // the argmax; u = Philox(seed, tokens generated so far)
  bool sampler_on = false;
  float s_temperature = 0.f, s_top_p = 1.f, s_min_p = 0.f;
  int64_t s_top_k = 0;
  unsigned long long s_seed = 0;
  void* sample_ws = nullptr;
  // async token pipeline (b200_engine_set_mailbox): ring in pinned host memory the argmax kernel posts tokens into
  unsigned long long* mailbox = nullptr;
  unsigned long long mailbox_cap = 1;
};

namespace b200 {

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

namespace {

// embed: x = E[cur_tok]   (first kernel of every token)
__global__ void __launch_bounds__(128) embed_step_kernel(__nv_bfloat16* __restrict__ x,
                                                         const __nv_bfloat16* __restrict__ table,
                                                         const int64_t* __restrict__ tok, int64_t V, int H) {
  pdl_trigger();
  pdl_wait();
  int64_t id = tok[blockIdx.x];   // one CTA per sequence of a batched step
  if (id < 0 || id >= V) id = 0;
  const uint4* s4 = reinterpret_cast<const uint4*>(table + id * H);
  uint4* d4 = reinterpret_cast<uint4*>(x + (size_t)blockIdx.x * H);
  for (int i = threadIdx.x; i < (H >> 3); i += blockDim.x) d4[i] = s4[i];
}

// Tensor parallel: last kernel of a token.  With a head: wait for every rank's (max logit, global index) candidate,
// choose with the reference tie rule (highest index among equal maxima), publish; always: advance position and epoch.
struct TpFinish {
  const uint2* cand;               // local window: world × 2 words {max-logit bits, tag}, {global index, tag}
  unsigned long long* epoch;       // tokens completed
  int* pos;
  int64_t* cur_tok;
  int64_t* gen_log;
  unsigned long long* gen_count;
  int gen_cap;
  int world;
  int with_head;
};
__global__ void tp_finish_kernel(const TpFinish f) {
  pdl_trigger();
  pdl_wait();
  if (threadIdx.x != 0) return;
  const unsigned long long ep = *f.epoch;
  if (f.with_head) {
    const unsigned int want = (unsigned int)(ep + 1ull);
    float best = -INFINITY;
    int64_t bi = -1;
    for (int r = 0; r < f.world; ++r) {
      const volatile uint2* c = f.cand + 2 * r;
      uint2 a, b;
      unsigned int spins = 0;
      unsigned long long t0 = 0;
      for (;;) {
        a.x = c[0].x; a.y = c[0].y;
        b.x = c[1].x; b.y = c[1].y;
        if (a.y == want && b.y == want) break;
        if ((++spins & 0x3fffu) == 0) {  // a peer died: fail loudly, never hang
          const unsigned long long now = global_timer_ns();
          if (t0 == 0) t0 = now;
          if (now - t0 > 4000000000ull) __trap();
        }
      }
      const float val = __uint_as_float(a.x);
      const int64_t idx = (int64_t)b.x;
      if (val > best || (val == best && idx > bi)) {
        best = val;
        bi = idx;
      }
    }
    *f.cur_tok = bi;
    const unsigned long long c = *f.gen_count;
    f.gen_log[c % (unsigned long long)f.gen_cap] = bi;
    *f.gen_count = c + 1;
  }
  *f.pos += 1;
  *f.epoch = ep + 1;
}

}  // namespace

// exchange-window layout (identical on every rank; b200_tp_window_bytes in tp.cu sizes it)
static inline size_t tp_vec_bytes(int H) { return ((size_t)H * 8 + 255) / 256 * 256; }  // {value, tag} words
static inline size_t tp_slot_off(const b200_engine* e, int point, int rank) {
  return ((size_t)point * e->tp_world + rank) * tp_vec_bytes(e->d.hidden);
}
static inline size_t tp_flags_off(const b200_engine* e) {
  return (size_t)(2 * e->d.layers + 2) * e->tp_world * tp_vec_bytes(e->d.hidden);
}
static inline size_t tp_flag_off(const b200_engine* e, int point) { return tp_flags_off(e) + (size_t)point * 256; }
static inline size_t tp_cand_off(const b200_engine* e) { return tp_flags_off(e) + (size_t)(2 * e->d.layers + 2) * 256; }

static int engine_launch_token(b200_engine* e, cudaStream_t st, bool with_head) {
  const b200_model_desc& d = e->d;
  const bool pdl = e->use_pdl;
  int rc;
  // NOTE (measured on B200, CUDA 12.9 / driver 580): when the LAST kernel node of a captured graph has a programmatic
  // (PDL) incoming edge, work enqueued after the graph launch can start before that node has finished.  The last
  // launch of each graph therefore uses a normal full dependency, and so does the first (it has no upstream).
  B200_CUDA(launch_pdl(embed_step_kernel, dim3(1), dim3(128), 0, st, false, e->x, e->embed, (const int64_t*)e->cur_tok,
                       (int64_t)d.vocab, (int)d.hidden));
  const size_t kv_layer = (size_t)d.max_ctx * e->Hkv_l * d.head_dim;
  int slot = 0;
  auto tr = [&]() -> unsigned long long* { return e->trace ? e->trace + 8 * (slot++) : nullptr; };
  for (int l = 0; l < d.layers; ++l) {
    GemvPlan q = e->p_qkv[l];
    q.p.trace = tr();
    if ((rc = gemv_launch(q, st, pdl)) != B200_OK) return rc;
    AttnDecodeParams a{};
    a.trace = tr();
    a.qkv = e->qkv;
    a.q_norm = (const __nv_bfloat16*)e->lw[l].q_norm;
    a.k_norm = (const __nv_bfloat16*)e->lw[l].k_norm;
    a.eps = d.rms_eps;
    a.rope = e->rope;
    a.pos = e->pos;
    a.fixed_len = 0;
    a.kcache = e->kcache + (size_t)l * kv_layer;
    a.vcache = e->vcache + (size_t)l * kv_layer;
    a.out = e->attn;
    a.ws = e->attn_ws;
    a.tickets = e->attn_tickets;
    a.Hq = e->Hq_l;
    a.Hkv = e->Hkv_l;
    a.nsplit = e->nsplit;
    a.max_ctx = d.max_ctx;
    if ((rc = launch_attn_decode(a, d.head_dim, st, pdl)) != B200_OK) return rc;
    GemvPlan o = e->p_o[l], gu = e->p_gu[l];
    o.p.trace = tr();
    gu.p.trace = tr();
    if ((rc = gemv_launch(o, st, pdl)) != B200_OK) return rc;
    if ((rc = gemv_launch(gu, st, pdl)) != B200_OK) return rc;
    const bool last_node = !with_head && l == d.layers - 1;
    GemvPlan dn = e->p_down[l];
    const bool tp = e->tp_world > 1;
    dn.p.pos_inc = (last_node && !tp) ? e->pos : nullptr;  // the LAST kernel of a token advances the position
    dn.p.trace = tr();
    if ((rc = gemv_launch(dn, st, pdl && !(last_node && !tp))) != B200_OK) return rc;
  }
  if (with_head) {
    GemvPlan hd = e->p_head;
    hd.p.trace = tr();
    if ((rc = gemv_launch(hd, st, pdl)) != B200_OK) return rc;
    int64_t* amax = reinterpret_cast<int64_t*>((uint8_t*)e->argmax_ws + argmax_workspace_bytes(1, e->V_l));
    ArgmaxPublish pub;
    if (e->tp_world == 1) {
      pub.pos = e->pos;
      pub.cur_tok = e->cur_tok;
      pub.gen_log = e->gen_log;
      pub.gen_count = e->gen_count;
      pub.gen_cap = e->gen_cap;
      pub.mailbox = e->mailbox;
      pub.mailbox_cap = e->mailbox_cap;
      if (e->sampler_on) {
        if ((rc = launch_sample(nullptr, e->logits, e->V_l, e->s_temperature, e->s_top_k, e->s_top_p, e->s_min_p, 0.f,
                                e->sample_ws, st, &pub, e->s_seed, true)) != B200_OK)
          return rc;
      } else if ((rc = launch_argmax(amax, e->logits, 1, e->V_l, e->argmax_ws, st, false, &pub)) != B200_OK) {
        return rc;
      }
    } else {
      pub.tp_world = e->tp_world;
      pub.tp_index_offset = (int64_t)e->tp_rank * e->V_l;
      pub.tp_epoch = e->tp_epoch;
      for (int r = 0; r < e->tp_world; ++r)
        pub.tp_cand[r] = reinterpret_cast<uint2*>(e->win[r] + tp_cand_off(e) + (size_t)e->tp_rank * 16);
      if ((rc = launch_argmax(amax, e->logits, 1, e->V_l, e->argmax_ws, st, pdl, &pub)) != B200_OK) return rc;
    }
  }
  if (e->tp_world > 1) {
    TpFinish f{};
    f.cand = reinterpret_cast<const uint2*>(e->win[e->tp_rank] + tp_cand_off(e));
    f.epoch = e->tp_epoch;
    f.pos = e->pos;
    f.cur_tok = e->cur_tok;
    f.gen_log = e->gen_log;
    f.gen_count = e->gen_count;
    f.gen_cap = e->gen_cap;
    f.world = e->tp_world;
    f.with_head = with_head ? 1 : 0;
    B200_CUDA(launch_pdl(tp_finish_kernel, dim3(1), dim3(32), 0, st, false, f));  // last node: full dependency
  }
  return B200_OK;
}

static int engine_launch_token_batch(b200_engine* e, cudaStream_t st, bool with_head);

static int engine_capture(b200_engine* e, bool with_head, cudaGraphExec_t* out, bool batched = false) {
  cudaStream_t st;
  B200_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  cudaGraph_t g = nullptr;
  cudaError_t err = cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal);
  if (err != cudaSuccess) {
    cudaStreamDestroy(st);
    set_error("cudaStreamBeginCapture failed: %s", cudaGetErrorString(err));
    return B200_ERR_CUDA;
  }
  const int64_t before = g_launches.load();
  int rc = batched ? engine_launch_token_batch(e, st, with_head) : engine_launch_token(e, st, with_head);
  const int64_t n = g_launches.load() - before;
  g_launches.fetch_sub(n);  // capture does not execute anything
  err = cudaStreamEndCapture(st, &g);
  if (rc != B200_OK || err != cudaSuccess) {
    if (g) cudaGraphDestroy(g);
    cudaStreamDestroy(st);
    if (rc == B200_OK) {
      set_error("cudaStreamEndCapture failed: %s", cudaGetErrorString(err));
      rc = B200_ERR_CUDA;
    }
    return rc;
  }
  if (batched) (with_head ? e->b_launches : e->b_launches_body) = (int)n;
  else if (with_head) e->launches_per_token = (int)n;
  else e->launches_body = (int)n;
  err = cudaGraphInstantiate(out, g, 0);
  cudaGraphDestroy(g);
  cudaStreamDestroy(st);
  if (err != cudaSuccess) {
    set_error("cudaGraphInstantiate failed: %s", cudaGetErrorString(err));
    return B200_ERR_CUDA;
  }
  return B200_OK;
}


// ---------------------------------------------------------------------------------------------------- batched decode
// B sequences at the same position [ref: GPTEngine::generateSync left-pads the prompts to one length and runs the model
// on [B, S] ids, src/engine/GPTEngine.cpp:101-174; every Linear then sees m = B rows].  One graph per token: embed (B
// CTAs) + L × {qkv, attention (grid.z = B), o_proj, gate|up, down} with the batched GEMV (W streamed once for all B) +
// lm_head + argmax (B rows).
static int engine_launch_token_batch(b200_engine* e, cudaStream_t st, bool with_head) {
  const b200_model_desc& d = e->d;
  const bool pdl = e->use_pdl;
  const int B = e->b_cap;
  int rc;
  B200_CUDA(launch_pdl(embed_step_kernel, dim3(B), dim3(128), 0, st, false, e->bx, e->embed, (const int64_t*)e->b_tok,
                       (int64_t)d.vocab, (int)d.hidden));
  const size_t kv_layer = (size_t)d.max_ctx * e->Hkv_l * d.head_dim;
  const int nqkv = e->qdim + 2 * e->kvdim;
  int slot = 0;
  auto tr = [&]() -> unsigned long long* { return e->trace ? e->trace + 8 * (slot++) : nullptr; };
  for (int l = 0; l < d.layers; ++l) {
    GemvPlan q = e->bp_qkv[l];
    q.p.trace = tr();
    if ((rc = gemv_launch(q, st, pdl)) != B200_OK) return rc;
    AttnDecodeParams a{};
    a.trace = tr();
    a.qkv = e->bqkv;
    a.q_norm = (const __nv_bfloat16*)e->lw[l].q_norm;
    a.k_norm = (const __nv_bfloat16*)e->lw[l].k_norm;
    a.eps = d.rms_eps;
    a.rope = e->rope;
    a.pos = e->pos;
    a.kcache = e->bk + (size_t)l * kv_layer;
    a.vcache = e->bv + (size_t)l * kv_layer;
    a.out = e->battn;
    a.ws = e->b_attn_ws;
    a.tickets = e->b_tickets;
    a.Hq = e->Hq_l;
    a.Hkv = e->Hkv_l;
    a.nsplit = e->nsplit;
    a.max_ctx = d.max_ctx;
    a.batch = B;
    a.qkv_bstride = nqkv;
    a.out_bstride = e->qdim;
    a.cache_bstride = (long long)d.layers * (long long)kv_layer;
    a.ws_bstride = attn_decode_ws_floats(e->Hq_l, e->Hkv_l, d.head_dim, e->nsplit);
    a.tick_bstride = e->Hq_l;
    if ((rc = launch_attn_decode(a, d.head_dim, st, pdl)) != B200_OK) return rc;
    GemvPlan o = e->bp_o[l], gu = e->bp_gu[l];
    o.p.trace = tr();
    gu.p.trace = tr();
    if ((rc = gemv_launch(o, st, pdl)) != B200_OK) return rc;
    if ((rc = gemv_launch(gu, st, pdl)) != B200_OK) return rc;
    const bool last_node = !with_head && l == d.layers - 1;
    GemvPlan dn = e->bp_down[l];
    dn.p.pos_inc = last_node ? e->pos : nullptr;   // the LAST kernel of a token advances the position
    dn.p.trace = tr();
    if ((rc = gemv_launch(dn, st, pdl && !last_node)) != B200_OK) return rc;
  }
  if (with_head) {
    GemvPlan hd = e->bp_head;
    hd.p.trace = tr();
    if ((rc = gemv_launch(hd, st, pdl)) != B200_OK) return rc;
    int64_t* amax = reinterpret_cast<int64_t*>((uint8_t*)e->b_argmax_ws + argmax_workspace_bytes(B, e->V_l));
    ArgmaxPublish pub;
    pub.pos = e->pos;
    pub.cur_tok = e->b_tok;
    pub.gen_log = e->b_log;
    pub.gen_count = e->b_cnt;
    pub.gen_cap = e->gen_cap;
    pub.batch_rows = B;
    if ((rc = launch_argmax(amax, e->blogits, B, e->V_l, e->b_argmax_ws, st, false, &pub)) != B200_OK) return rc;
  }
  return B200_OK;
}

static void engine_batch_free(b200_engine* e) {
  if (e->gb_step) cudaGraphExecDestroy(e->gb_step);
  if (e->gb_body) cudaGraphExecDestroy(e->gb_body);
  if (e->b_arena) cudaFree(e->b_arena);
  e->gb_step = e->gb_body = nullptr;
  e->b_arena = nullptr;
  e->b_cap = 0;
}

// Buffers, plans and graphs for a batch of B sequences (kept until a forward asks for another B).
static int engine_batch_prepare(b200_engine* e, int B) {
  if (e->b_cap == B) return B200_OK;
  const b200_model_desc& d = e->d;
  B200_CHECK_ARG(B >= 2 && B <= kMaxBatch, "engine_forward: batch %d not built (1 … %d sequences per step)", B, kMaxBatch);
  if (e->tp_world > 1 || e->sampler_on || e->mailbox != nullptr) {
    set_error("engine_forward: batched decode is single-GPU, greedy, without a token mailbox");
    return B200_ERR_UNSUPPORTED;
  }
  int rc;
  if ((rc = gemv_batch_setup_attributes()) != B200_OK) return rc;
  B200_CUDA(cudaDeviceSynchronize());
  engine_batch_free(e);
  const size_t kv_layer = (size_t)d.max_ctx * e->Hkv_l * d.head_dim;
  const size_t kv_bytes = (size_t)B * d.layers * kv_layer * 2;
  const int nqkv = e->qdim + 2 * e->kvdim;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    const size_t o = off;
    off = align_up(off + bytes, 256);
    return o;
  };
  const size_t o_k = take(kv_bytes), o_v = take(kv_bytes);
  const size_t o_x = take((size_t)B * d.hidden * 2), o_qkv = take((size_t)B * nqkv * 2);
  const size_t o_attn = take((size_t)B * e->qdim * 2), o_act = take((size_t)B * e->I_l * 2);
  const size_t o_logits = take((size_t)B * e->V_l * 2);
  const size_t ws_floats = (size_t)attn_decode_ws_floats(e->Hq_l, e->Hkv_l, d.head_dim, e->nsplit);
  const size_t o_ws = take((size_t)B * ws_floats * 4), o_tick = take((size_t)B * e->Hq_l * 4);
  const size_t o_amax = take((size_t)argmax_workspace_bytes(B, e->V_l) + (size_t)B * 8 + 16);
  const size_t o_tok = take((size_t)B * 8), o_log = take((size_t)e->gen_cap * B * 8), o_cnt = take((size_t)B * 8);
  B200_CUDA(cudaMalloc((void**)&e->b_arena, off));
  B200_CUDA(cudaMemset(e->b_arena + o_x, 0, off - o_x));   // control + activations (KV rows are written before read)
  e->bk = (__nv_bfloat16*)(e->b_arena + o_k);
  e->bv = (__nv_bfloat16*)(e->b_arena + o_v);
  e->bx = (__nv_bfloat16*)(e->b_arena + o_x);
  e->bqkv = (__nv_bfloat16*)(e->b_arena + o_qkv);
  e->battn = (__nv_bfloat16*)(e->b_arena + o_attn);
  e->bact = (__nv_bfloat16*)(e->b_arena + o_act);
  e->blogits = (__nv_bfloat16*)(e->b_arena + o_logits);
  e->b_attn_ws = (float*)(e->b_arena + o_ws);
  e->b_tickets = (unsigned int*)(e->b_arena + o_tick);
  e->b_argmax_ws = (void*)(e->b_arena + o_amax);
  e->b_tok = (int64_t*)(e->b_arena + o_tok);
  e->b_log = (int64_t*)(e->b_arena + o_log);
  e->b_cnt = (unsigned long long*)(e->b_arena + o_cnt);
  e->hb_gen = 0;
  e->b_cap = B;
  // plans: the single-sequence plans (same tensor maps, shapes, ring rule) pointed at the batch buffers
  e->bp_qkv = e->p_qkv;
  e->bp_o = e->p_o;
  e->bp_gu = e->p_gu;
  e->bp_down = e->p_down;
  e->bp_head = e->p_head;
  for (int l = 0; l < d.layers; ++l) {
    GemvPlan &q = e->bp_qkv[l], &o = e->bp_o[l], &g = e->bp_gu[l], &dn = e->bp_down[l];
    q.p.x = e->bx;
    q.p.y = e->bqkv;
    o.p.x = e->battn;
    o.p.residual = e->bx;
    o.p.y = e->bx;
    g.p.x = e->bx;
    g.p.y = e->bact;
    dn.p.x = e->bact;
    dn.p.residual = e->bx;
    dn.p.y = e->bx;
    if ((rc = gemv_plan_set_batch(&q, B)) != B200_OK || (rc = gemv_plan_set_batch(&o, B)) != B200_OK ||
        (rc = gemv_plan_set_batch(&g, B)) != B200_OK || (rc = gemv_plan_set_batch(&dn, B)) != B200_OK)
      return rc;
  }
  e->bp_head.p.x = e->bx;
  e->bp_head.p.y = e->blogits;
  if ((rc = gemv_plan_set_batch(&e->bp_head, B)) != B200_OK) return rc;
  if (e->use_graph) {
    if ((rc = engine_capture(e, true, &e->gb_step, true)) != B200_OK) return rc;
    if ((rc = engine_capture(e, false, &e->gb_body, true)) != B200_OK) return rc;
  }
  return B200_OK;
}

static int engine_run_token_batch(b200_engine* e, cudaStream_t st, bool with_head) {
  if (e->use_graph) {
    B200_CUDA(cudaGraphLaunch(with_head ? e->gb_step : e->gb_body, st));
    g_launches.fetch_add(with_head ? e->b_launches : e->b_launches_body);
    return B200_OK;
  }
  return engine_launch_token_batch(e, st, with_head);
}

static int engine_run_token(b200_engine* e, cudaStream_t st, bool with_head) {
  if (e->use_graph) {
    cudaGraphExec_t g = with_head ? e->g_step : e->g_body;
    B200_CUDA(cudaGraphLaunch(g, st));
    g_launches.fetch_add(with_head ? e->launches_per_token : e->launches_body);
    return B200_OK;
  }
  return engine_launch_token(e, st, with_head);
}


static int engine_build(const b200_model_desc* desc, const b200_weight_table* w, void* const* windows,
                        b200_engine** out) {
  B200_CHECK_ARG(desc && w && out, "engine_create: null argument");
  *out = nullptr;
  int rc = b200_device_check();
  if (rc != B200_OK) return rc;
  const b200_model_desc& d = *desc;
  B200_CHECK_ARG(d.hidden > 0 && d.layers > 0 && d.q_heads > 0 && d.kv_heads > 0 && d.intermediate > 0 && d.vocab > 0 &&
                     d.max_ctx > 0,
                 "engine_create: non-positive model dimension");
  B200_CHECK_ARG(d.head_dim == 64 || d.head_dim == 128, "engine_create: head_dim %d not built (64, 128)", d.head_dim);
  B200_CHECK_ARG(d.q_heads % d.kv_heads == 0, "engine_create: q_heads %% kv_heads != 0");
  B200_CHECK_ARG(d.hidden % 8 == 0 && d.intermediate % 8 == 0, "engine_create: H and I must be multiples of 8");
  const int world = windows ? d.tp_world : 1;
  B200_CHECK_ARG(windows != nullptr || d.tp_world <= 1, "engine_create: use b200_engine_create_tp for tensor-parallel engines");
  B200_CHECK_ARG(world >= 1 && world <= kMaxTpWorld && d.tp_rank >= 0 && d.tp_rank < std::max(world, 1),
                 "engine_create_tp: bad rank %d / world %d", d.tp_rank, d.tp_world);
  const bool shard_attn = world > 1 && d.tp_shard_attn != 0;
  B200_CHECK_ARG(d.intermediate % world == 0 && d.vocab % world == 0 && (d.intermediate / world) % 8 == 0,
                 "engine_create_tp: intermediate/vocab not divisible by world %d", world);
  B200_CHECK_ARG(!shard_attn || (d.q_heads % world == 0 && d.kv_heads % world == 0),
                 "engine_create_tp: heads (%d/%d) not divisible by world %d — pass tp_shard_attn = 0", d.q_heads,
                 d.kv_heads, world);
  B200_CHECK_ARG(w->embed && w->final_norm && w->lm_head && w->rope_table && w->layers_host,
                 "engine_create: null weight pointer");

  b200_engine* e = new (std::nothrow) b200_engine();
  if (!e) {
    set_error("engine_create: out of host memory");
    return B200_ERR_INVALID;
  }
  struct Guard {
    b200_engine* e;
    ~Guard() {
      if (e) b200_engine_destroy(e);
    }
  } guard{e};

  e->d = d;
  e->tp_world = world;
  e->tp_rank = world > 1 ? d.tp_rank : 0;
  e->shard_attn = shard_attn;
  e->d.tp_world = world;
  e->d.tp_rank = e->tp_rank;
  e->Hq_l = shard_attn ? d.q_heads / world : d.q_heads;
  e->Hkv_l = shard_attn ? d.kv_heads / world : d.kv_heads;
  e->I_l = d.intermediate / world;
  e->V_l = d.vocab / world;
  for (int r = 0; r < world && windows; ++r) {
    B200_CHECK_ARG(windows[r] != nullptr, "engine_create_tp: window %d is null", r);
    e->win[r] = (uint8_t*)windows[r];
  }
  e->qdim = e->Hq_l * d.head_dim;
  e->kvdim = e->Hkv_l * d.head_dim;
  e->embed = (const __nv_bfloat16*)w->embed;
  e->rope = w->rope_table;
  e->lw.assign(w->layers_host, w->layers_host + d.layers);
  for (int l = 0; l < d.layers; ++l) {
    const b200_layer_weights& lw = e->lw[l];
    B200_CHECK_ARG(lw.input_norm && lw.qkv_w && lw.o_w && lw.post_norm && lw.gate_up_w && lw.down_w,
                   "engine_create: layer %d has a null weight", l);
    B200_CHECK_ARG(!d.qkv_bias || lw.qkv_b, "engine_create: layer %d: qkv_bias set but qkv_b is null", l);
    B200_CHECK_ARG(!d.qk_norm || (lw.q_norm && lw.k_norm), "engine_create: layer %d: qk_norm set but norms null", l);
  }
  const char* env = std::getenv("B200_NO_GRAPH");
  e->use_graph = !(env && env[0] == '1');
  env = std::getenv("B200_NO_PDL");
  e->use_pdl = !(env && env[0] == '1');
  env = std::getenv("B200_NO_PREFILL_GEMM");
  e->use_prefill_gemm = !(env && env[0] == '1');
  e->final_norm_w = w->final_norm;

  int dev = 0;
  B200_CUDA(cudaGetDevice(&dev));
  B200_CUDA(cudaDeviceGetAttribute(&e->num_sms, cudaDevAttrMultiProcessorCount, dev));
  B200_CUDA(cudaDeviceSynchronize());  // weights may still be in flight on the loader's stream
  if ((rc = gemv_setup_attributes()) != B200_OK) return rc;
  if ((rc = attn_setup_attributes()) != B200_OK) return rc;

  e->nsplit = attn_decode_nsplit(d.head_dim, d.max_ctx);
  env = std::getenv("B200_TRACE");
  if (env && env[0] == '1') {
    B200_CUDA(cudaMalloc((void**)&e->trace, (size_t)(5 * d.layers + 8) * 8 * 8));
    B200_CUDA(cudaMemset(e->trace, 0, (size_t)(5 * d.layers + 8) * 8 * 8));
  }
  e->gen_cap = 1 << 16;

  // ---- one arena for everything the engine owns
  const size_t kv_bytes = (size_t)d.layers * d.max_ctx * e->kvdim * 2;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    const size_t o = off;
    off = align_up(off + bytes, 256);
    return o;
  };
  const size_t o_k = take(kv_bytes), o_v = take(kv_bytes);
  const size_t o_x = take((size_t)d.hidden * 2);
  const size_t o_x2 = take((size_t)d.hidden * 2);
  const size_t o_qkv = take((size_t)(e->qdim + 2 * e->kvdim) * 2);
  const size_t o_attn = take((size_t)e->qdim * 2);
  const size_t o_act = take((size_t)e->I_l * 2);
  const size_t o_logits = take((size_t)e->V_l * 2);
  const size_t o_ws = take((size_t)attn_decode_ws_floats(e->Hq_l, e->Hkv_l, d.head_dim, e->nsplit) * 4);
  const size_t o_tick = take((size_t)e->Hq_l * 4);
  const size_t o_amax = take((size_t)argmax_workspace_bytes(1, e->V_l) + 16);
  const size_t o_tok = take(8);
  const size_t o_log = take((size_t)e->gen_cap * 8);
  const size_t o_pos = take(16);
  const size_t o_cnt = take(8);
  const size_t o_epoch = take(16);
  const size_t o_sample = take((size_t)sample_workspace_bytes());
  e->arena_bytes = off;
  B200_CUDA(cudaMalloc((void**)&e->arena, e->arena_bytes));
  // zero only the small control region + activations (the KV cache is always written before it is read)
  B200_CUDA(cudaMemset(e->arena + o_x, 0, e->arena_bytes - o_x));
  e->kcache = (__nv_bfloat16*)(e->arena + o_k);
  e->vcache = (__nv_bfloat16*)(e->arena + o_v);
  e->x = (__nv_bfloat16*)(e->arena + o_x);
  e->x_alt = (__nv_bfloat16*)(e->arena + o_x2);
  e->tp_epoch = (unsigned long long*)(e->arena + o_epoch);
  e->qkv = (__nv_bfloat16*)(e->arena + o_qkv);
  e->attn = (__nv_bfloat16*)(e->arena + o_attn);
  e->act = (__nv_bfloat16*)(e->arena + o_act);
  e->logits = (__nv_bfloat16*)(e->arena + o_logits);
  e->attn_ws = (float*)(e->arena + o_ws);
  e->attn_tickets = (unsigned int*)(e->arena + o_tick);
  e->argmax_ws = (void*)(e->arena + o_amax);
  e->cur_tok = (int64_t*)(e->arena + o_tok);
  e->gen_log = (int64_t*)(e->arena + o_log);
  e->pos = (int*)(e->arena + o_pos);
  e->gen_count = (unsigned long long*)(e->arena + o_cnt);
  e->sample_ws = (void*)(e->arena + o_sample);

  // ---- shared-memory budgets = ring depth per GEMV.  Measured (B200, profiles/r02_ring_depth_sweep.txt): what decides
  // the HBM-bound models is that EVERY kernel keeps ≈ 8 stages (128 KB per SM, 19 MB chip-wide) in flight — enough to
  // cover the HBM latency at full bandwidth.  Round 1 split 220 KB between consecutive kernels so that PDL could keep
  // both resident; that left down_proj / o_proj with 2–4 stages and they streamed at ≈ 4 TB/s (Mistral-7B 0.60–0.79 of
  // the roofline depending on the box; 0.90 with 8 stages everywhere, Llama-3.2-3B 0.66 → 0.80, Qwen3-1.7B 0.61 →
  // 0.65).  So: every GEMV gets what it wants up to 8 stages; kernels that want less (the small models' matrices fit
  // whole) still co-reside and prefetch under PDL; the lm_head, alone at the end of the token, gets the maximum.
  const int nqkv_rows = e->qdim + 2 * e->kvdim;
  const int kStages = env_int("B200_GEMV_RING", 8, 2, 13);
  auto budget = [&](int64_t n, int64_t k, int nseg) {
    const int64_t k_pad = (k + 1023) / 1024 * 1024;
    const int cap = kStages * (16 * 1024 + 16) + (int)k_pad * 2 + 64;
    return std::min(gemv_smem_wanted(n, k, nseg, e->num_sms), std::min(cap, kGemvMaxSmem));
  };
  int b_qkv = budget(nqkv_rows, d.hidden, 1);
  int b_o = budget(d.hidden, e->qdim, 1);
  int b_gu = budget(e->I_l, d.hidden, 2);
  int b_dn = budget(d.hidden, e->I_l, 1);
  const int b_head = kGemvMaxSmem;
  if (std::getenv("B200_UNIFORM_SMEM")) b_qkv = b_o = b_gu = b_dn = kGemvDefaultSmem;

  // ---- GEMV plans (TMA descriptors are encoded once, here).
  // Tensor parallel: the row-sharded o_proj / down_proj push fp32 partials of the hidden vector into every rank's
  // window (EPI_TP_PUSH) and the NEXT GEMV's prologue reduces them, adds the residual and normalises
  // (PRO_TP_RMSNORM); the hidden state ping-pongs between two buffers because CTA 0 stores the new state while other
  // CTAs still read the old one.
  e->p_qkv.resize(d.layers);
  e->p_o.resize(d.layers);
  e->p_gu.resize(d.layers);
  e->p_down.resize(d.layers);
  const int nqkv = e->qdim + 2 * e->kvdim;
  const bool tp = world > 1;
  __nv_bfloat16* cur = e->x;          // buffer holding the residual stream (embed writes here)
  __nv_bfloat16* other = e->x_alt;
  int pending_point = -1;             // exchange point whose partials still have to be reduced into `cur`
  int next_point = 0;
  auto make_reduce_prologue = [&](GemvPlan& pl) {  // consume the pending exchange in this plan's prologue
    pl.p.tp_world = world;
    pl.p.tp_partials = reinterpret_cast<const uint2*>(e->win[e->tp_rank] + tp_slot_off(e, pending_point, 0));
    pl.p.tp_stride = (int)(tp_vec_bytes(d.hidden) / 8);
    pl.p.tp_epoch = e->tp_epoch;
    pl.p.tp_residual = cur;
    pl.p.tp_h_out = other;
    std::swap(cur, other);
    pending_point = -1;
  };
  auto make_push_epilogue = [&](GemvPlan& pl) {
    pl.p.tp_world = world;
    pl.p.tp_epoch = e->tp_epoch;
    for (int r = 0; r < world; ++r)
      pl.p.tp_push[r] = reinterpret_cast<uint2*>(e->win[r] + tp_slot_off(e, next_point, e->tp_rank));
    pending_point = next_point++;
  };
  for (int l = 0; l < d.layers; ++l) {
    const b200_layer_weights& lw = e->lw[l];
    GemvPlan& q = e->p_qkv[l];
    const int q_pro = pending_point >= 0 ? PRO_TP_RMSNORM : PRO_RMSNORM;
    if ((rc = gemv_make_plan(&q, lw.qkv_w, nqkv, nqkv, d.hidden, 1, q_pro, EPI_PLAIN, e->num_sms, b_qkv)) != B200_OK)
      return rc;
    q.p.x = cur;
    if (q_pro == PRO_TP_RMSNORM) make_reduce_prologue(q);
    q.p.norm_w = (const __nv_bfloat16*)lw.input_norm;
    q.p.eps = d.rms_eps;
    q.p.bias = d.qkv_bias ? (const __nv_bfloat16*)lw.qkv_b : nullptr;
    q.p.y = e->qkv;

    GemvPlan& o = e->p_o[l];
    const int o_epi = shard_attn ? EPI_TP_PUSH : EPI_RESIDUAL;
    if ((rc = gemv_make_plan(&o, lw.o_w, d.hidden, d.hidden, e->qdim, 1, PRO_PLAIN, o_epi, e->num_sms, b_o)) != B200_OK)
      return rc;
    o.p.x = e->attn;
    o.p.residual = cur;
    o.p.y = cur;
    if (o_epi == EPI_TP_PUSH) make_push_epilogue(o);

    GemvPlan& g = e->p_gu[l];
    const int g_pro = pending_point >= 0 ? PRO_TP_RMSNORM : PRO_RMSNORM;
    if ((rc = gemv_make_plan(&g, lw.gate_up_w, 2 * (int64_t)e->I_l, e->I_l, d.hidden, 2, g_pro, EPI_SILU_MUL,
                             e->num_sms, b_gu)) != B200_OK)
      return rc;
    g.p.x = cur;
    if (g_pro == PRO_TP_RMSNORM) make_reduce_prologue(g);
    g.p.norm_w = (const __nv_bfloat16*)lw.post_norm;
    g.p.eps = d.rms_eps;
    g.p.y = e->act;

    GemvPlan& dn = e->p_down[l];
    const int d_epi = tp ? EPI_TP_PUSH : EPI_RESIDUAL;
    if ((rc = gemv_make_plan(&dn, lw.down_w, d.hidden, d.hidden, e->I_l, 1, PRO_PLAIN, d_epi, e->num_sms, b_dn)) !=
        B200_OK)
      return rc;
    dn.p.x = e->act;
    dn.p.residual = cur;
    dn.p.y = cur;
    if (d_epi == EPI_TP_PUSH) make_push_epilogue(dn);
  }
  const int h_pro = pending_point >= 0 ? PRO_TP_RMSNORM : PRO_RMSNORM;
  if ((rc = gemv_make_plan(&e->p_head, w->lm_head, e->V_l, e->V_l, d.hidden, 1, h_pro, EPI_PLAIN, e->num_sms,
                           b_head)) != B200_OK)
    return rc;
  e->p_head.p.x = cur;
  if (h_pro == PRO_TP_RMSNORM) make_reduce_prologue(e->p_head);
  e->p_head.p.norm_w = (const __nv_bfloat16*)w->final_norm;
  e->p_head.p.eps = d.rms_eps;
  e->p_head.p.y = e->logits;
  // with two reductions per layer (or one per layer and an even layer count) the stream is back in e->x when the next
  // token's embedding is written; otherwise the embedding must go to wherever layer 0 expects it — layer 0 always
  // reads e->x, and nothing of a previous token is read from `cur`, so no fix-up is needed.

  e->launches_per_token = 1 + 5 * d.layers + 2 + (tp ? 1 : 0);
  if (e->use_graph) {
    if ((rc = engine_capture(e, true, &e->g_step)) != B200_OK) return rc;
    if ((rc = engine_capture(e, false, &e->g_body)) != B200_OK) return rc;
  }
  guard.e = nullptr;
  *out = e;
  return B200_OK;
}

// ------------------------------------------------------------------------------------------------ batched prefill
// S prompt tokens at once (single GPU): activations [S, ·], every Linear on the tcgen05 GEMM, the reference's rounding
// points kept by running bias / residual / SiLU·mul / norms as the separate ops they are in the reference
// [ref: src/model/GPTModel.h:51-58; src/layer/Attention.h:71-112; src/layer/GatedMLP.h:37-41].  K/V rows of all S
// tokens are written in place; the lm_head runs for the LAST position only (GPTEngine::genNextToken narrows to it).
static constexpr int kPrefillChunk = 2048;
static constexpr int kPrefillMin = 8;

static int prefill_workspace(b200_engine* e, int chunk) {
  if (e->pf_arena != nullptr && e->pf_chunk >= chunk) return B200_OK;
  if (e->pf_arena) {
    B200_CUDA(cudaDeviceSynchronize());
    cudaFree(e->pf_arena);
    e->pf_arena = nullptr;
  }
  const b200_model_desc& d = e->d;
  const size_t nqkv = (size_t)e->qdim + 2 * e->kvdim;
  size_t off = 0;
  auto take = [&](size_t elems) {
    const size_t o = off;
    off = align_up(off + elems * 2, 256);
    return o;
  };
  const size_t o_x = take((size_t)chunk * d.hidden), o_h = take((size_t)chunk * d.hidden);
  const size_t o_qkv = take((size_t)chunk * nqkv), o_attn = take((size_t)chunk * e->qdim);
  const size_t o_gu = take((size_t)chunk * 2 * e->I_l), o_act = take((size_t)chunk * e->I_l);
  const size_t o_t = take((size_t)chunk * d.hidden);
  B200_CUDA(cudaMalloc((void**)&e->pf_arena, off));
  e->pf_chunk = chunk;
  e->pf_x = (__nv_bfloat16*)(e->pf_arena + o_x);
  e->pf_h = (__nv_bfloat16*)(e->pf_arena + o_h);
  e->pf_qkv = (__nv_bfloat16*)(e->pf_arena + o_qkv);
  e->pf_attn = (__nv_bfloat16*)(e->pf_arena + o_attn);
  e->pf_gu = (__nv_bfloat16*)(e->pf_arena + o_gu);
  e->pf_act = (__nv_bfloat16*)(e->pf_arena + o_act);
  e->pf_t = (__nv_bfloat16*)(e->pf_arena + o_t);
  return B200_OK;
}

// tokens ids[0..S) at positions p0 … p0+S-1; leaves the hidden state of the last token in e->x when `last`
static int prefill_chunk(b200_engine* e, const int64_t* ids, int S, int p0, bool last, cudaStream_t st) {
  const b200_model_desc& d = e->d;
  const int H = d.hidden, nqkv = e->qdim + 2 * e->kvdim;
  const size_t kv_layer = (size_t)d.max_ctx * e->Hkv_l * d.head_dim;
  int rc;
  if ((rc = launch_embedding(e->pf_x, e->embed, ids, S, d.vocab, H, st, false)) != B200_OK) return rc;
  for (int l = 0; l < d.layers; ++l) {
    const b200_layer_weights& lw = e->lw[l];
    if ((rc = launch_rmsnorm(e->pf_h, e->pf_x, lw.input_norm, S, H, d.rms_eps, st, false)) != B200_OK) return rc;
    if ((rc = launch_gemm_bf16(e->pf_qkv, e->pf_h, lw.qkv_w, S, nqkv, H, st)) != B200_OK) return rc;
    if (d.qkv_bias && (rc = launch_bias_add(e->pf_qkv, lw.qkv_b, S, nqkv, st)) != B200_OK) return rc;
    __nv_bfloat16* kc = e->kcache + (size_t)l * kv_layer;
    __nv_bfloat16* vc = e->vcache + (size_t)l * kv_layer;
    if ((rc = launch_prefill_qk(e->pf_qkv, lw.q_norm, lw.k_norm, d.rms_eps, e->rope, kc, vc, S, e->Hq_l, e->Hkv_l,
                                d.head_dim, p0, st)) != B200_OK)
      return rc;
    if ((rc = launch_attn_prefill(e->pf_attn, e->pf_qkv, kc, vc, S, e->Hq_l, e->Hkv_l, d.head_dim, p0, st)) != B200_OK)
      return rc;
    if ((rc = launch_gemm_bf16(e->pf_t, e->pf_attn, lw.o_w, S, H, e->qdim, st)) != B200_OK) return rc;
    if ((rc = launch_add(e->pf_x, e->pf_x, e->pf_t, (int64_t)S * H, st, false)) != B200_OK) return rc;
    if ((rc = launch_rmsnorm(e->pf_h, e->pf_x, lw.post_norm, S, H, d.rms_eps, st, false)) != B200_OK) return rc;
    if ((rc = launch_gemm_bf16(e->pf_gu, e->pf_h, lw.gate_up_w, S, 2 * (int64_t)e->I_l, H, st)) != B200_OK) return rc;
    if ((rc = launch_silu_mul(e->pf_act, e->pf_gu, S, e->I_l, st, false)) != B200_OK) return rc;
    if ((rc = launch_gemm_bf16(e->pf_t, e->pf_act, lw.down_w, S, H, e->I_l, st)) != B200_OK) return rc;
    if ((rc = launch_add(e->pf_x, e->pf_x, e->pf_t, (int64_t)S * H, st, false)) != B200_OK) return rc;
  }
  if (last)
    B200_CUDA(cudaMemcpyAsync(e->x, e->pf_x + (size_t)(S - 1) * H, (size_t)H * 2, cudaMemcpyDeviceToDevice, st));
  return B200_OK;
}

// whole prompt through the GEMM path, then lm_head + argmax for the last position
static int engine_prefill(b200_engine* e, const int64_t* ids, int64_t S, cudaStream_t st) {
  int rc;
  // tokens per pass through the layers (default 2048: measured 22.1 → 14.3 ms for config 4 against chunks of 512):
  // B200_PREFILL_CHUNK (multiple of 128, 128 … 8192) overrides — bigger
  // chunks give the o_proj / down_proj GEMMs more than 64 tiles for 148 SMs at the price of a bigger workspace
  int chunk = kPrefillChunk;
  if (const char* env = std::getenv("B200_PREFILL_CHUNK")) {
    const int v = std::atoi(env);
    if (v >= 128 && v <= 8192 && v % 128 == 0) chunk = v;
  }
  if ((rc = prefill_workspace(e, (int)std::min<int64_t>(S, chunk))) != B200_OK) return rc;
  const int p_start = (int)e->h_pos;
  for (int64_t t0 = 0; t0 < S; t0 += chunk) {
    const int Sc = (int)std::min<int64_t>(chunk, S - t0);
    if ((rc = prefill_chunk(e, ids + t0, Sc, p_start + (int)t0, t0 + Sc == S, st)) != B200_OK) return rc;
  }
  // the head's argmax advances pos by one: park it on the last prompt position first
  const int v[2] = {p_start + (int)S - 1, 0};
  B200_CUDA(cudaMemcpyAsync(e->pos, v, 4, cudaMemcpyHostToDevice, st));
  if ((rc = gemv_launch(e->p_head, st, false)) != B200_OK) return rc;
  int64_t* amax = reinterpret_cast<int64_t*>((uint8_t*)e->argmax_ws + argmax_workspace_bytes(1, e->V_l));
  ArgmaxPublish pub;
  pub.pos = e->pos;
  pub.cur_tok = e->cur_tok;
  pub.gen_log = e->gen_log;
  pub.gen_count = e->gen_count;
  pub.gen_cap = e->gen_cap;
  pub.mailbox = e->mailbox;
  pub.mailbox_cap = e->mailbox_cap;
  if (e->sampler_on)
    return launch_sample(nullptr, e->logits, e->V_l, e->s_temperature, e->s_top_k, e->s_top_p, e->s_min_p, 0.f,
                         e->sample_ws, st, &pub, e->s_seed, true);
  return launch_argmax(amax, e->logits, 1, e->V_l, e->argmax_ws, st, false, &pub);
}

}  // namespace b200

// --------------------------------------------------------------------------------------------------------- C ABI
extern "C" {

int b200_engine_create(const b200_model_desc* desc, const b200_weight_table* weights, b200_engine** out) {
  return b200::engine_build(desc, weights, nullptr, out);
}

int b200_engine_create_tp(const b200_model_desc* desc, const b200_weight_table* weights, void* const* windows_host,
                          b200_engine** out) {
  using namespace b200;
  B200_CHECK_ARG(windows_host != nullptr, "engine_create_tp: windows_host is null");
  return engine_build(desc, weights, windows_host, out);
}

void b200_engine_destroy(b200_engine* e) {
  if (!e) return;
  if (e->g_step) cudaGraphExecDestroy(e->g_step);
  if (e->g_body) cudaGraphExecDestroy(e->g_body);
  if (e->arena) cudaFree(e->arena);
  if (e->pf_arena) cudaFree(e->pf_arena);
  if (e->trace) cudaFree(e->trace);
  b200::engine_batch_free(e);
  delete e;
}

int b200_engine_reset(b200_engine* e, void* stream) {
  using namespace b200;
  B200_CHECK_ARG(e, "engine_reset: null engine");
  B200_CUDA(cudaMemsetAsync(e->pos, 0, 16, (cudaStream_t)stream));
  e->h_pos = 0;
  return B200_OK;
}

int b200_engine_seek(b200_engine* e, int64_t position, void* stream) {
  using namespace b200;
  B200_CHECK_ARG(e, "engine_seek: null engine");
  if (position < 0 || position > e->h_pos) {
    set_error("engine_seek: position %lld is not inside the cached prefix [0, %lld]", (long long)position,
              (long long)e->h_pos);
    return B200_ERR_STATE;
  }
  const int v[2] = {(int)position, (int)position};
  // small enough to travel in the command stream; stream-ordered with the graphs that read it
  B200_CUDA(cudaMemcpyAsync(e->pos, v, 8, cudaMemcpyHostToDevice, (cudaStream_t)stream));
  e->h_pos = position;
  return B200_OK;
}

int b200_engine_forward(b200_engine* e, const int64_t* ids, int64_t B, int64_t S, void* logits_out, int logits_mode,
                        void* stream) {
  using namespace b200;
  B200_CHECK_ARG(e && ids, "engine_forward: null argument");
  B200_CHECK_ARG(B >= 1 && B <= kMaxBatch, "engine_forward: batch %lld not built (1 … %d sequences per step)", (long long)B,
                 kMaxBatch);
  B200_CHECK_ARG(S >= 1, "engine_forward: empty sequence");
  B200_CHECK_ARG(logits_mode == 0 || logits_mode == 1, "engine_forward: logits_mode must be 0 or 1");
  if (e->h_pos + S > e->d.max_ctx) {
    set_error("engine_forward: position %lld + %lld tokens exceeds max_ctx %d", (long long)e->h_pos, (long long)S,
              e->d.max_ctx);
    return B200_ERR_STATE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const size_t vbytes = (size_t)e->V_l * 2;
  if (B > 1) {
    // ---- batch of B sequences at the same position (ids [B, S] row-major, logits [B, 1 or S, V])
    int rc = engine_batch_prepare(e, (int)B);
    if (rc != B200_OK) return rc;
    e->batch = (int)B;
    const size_t kv_seq = (size_t)e->d.layers * e->d.max_ctx * e->Hkv_l * e->d.head_dim;
    if (e->use_prefill_gemm && logits_mode == 0 && S >= kPrefillMin) {
      // prompts: the batch-1 tensor-core prefill, sequence by sequence, into that sequence's KV cache (prefill is
      // compute-bound: nothing to share between sequences); it leaves the last-position logits and the greedy token
      __nv_bfloat16 *k0 = e->kcache, *v0 = e->vcache;
      for (int64_t b = 0; b < B; ++b) {
        e->kcache = e->bk + (size_t)b * kv_seq;
        e->vcache = e->bv + (size_t)b * kv_seq;
        rc = engine_prefill(e, ids + b * S, S, st);
        e->kcache = k0;
        e->vcache = v0;
        if (rc != B200_OK) return rc;
        B200_CUDA(cudaMemcpyAsync(e->blogits + (size_t)b * e->V_l, e->logits, vbytes, cudaMemcpyDeviceToDevice, st));
        B200_CUDA(cudaMemcpyAsync(e->b_tok + b, e->cur_tok, 8, cudaMemcpyDeviceToDevice, st));
        e->h_gen += 1;   // the prefill's argmax appended to the single-sequence log
      }
      if (logits_out != nullptr)
        B200_CUDA(cudaMemcpyAsync(logits_out, e->blogits, (size_t)B * vbytes, cudaMemcpyDeviceToDevice, st));
      e->h_pos += S;
      return B200_OK;
    }
    for (int64_t t = 0; t < S; ++t) {
      // column t of ids → the B current tokens
      B200_CUDA(cudaMemcpy2DAsync(e->b_tok, 8, ids + t, (size_t)S * 8, 8, (size_t)B, cudaMemcpyDeviceToDevice, st));
      const bool head = (t == S - 1) || (logits_mode == 1 && logits_out != nullptr);
      if ((rc = engine_run_token_batch(e, st, head)) != B200_OK) return rc;
      if (head) e->hb_gen += 1;
      if (head && logits_out != nullptr) {
        if (logits_mode == 1)   // [B][S][V]: row t of every sequence
          B200_CUDA(cudaMemcpy2DAsync((uint8_t*)logits_out + (size_t)t * vbytes, (size_t)S * vbytes, e->blogits, vbytes,
                                      vbytes, (size_t)B, cudaMemcpyDeviceToDevice, st));
        else
          B200_CUDA(cudaMemcpyAsync(logits_out, e->blogits, (size_t)B * vbytes, cudaMemcpyDeviceToDevice, st));
      }
    }
    e->h_pos += S;
    return B200_OK;
  }
  e->batch = 1;
  if (e->use_prefill_gemm && e->tp_world == 1 && logits_mode == 0 && S >= kPrefillMin) {
    // batched prefill: every Linear of the prompt on the tcgen05 GEMM, lm_head for the last position only
    int rc = engine_prefill(e, ids, S, st);
    if (rc != B200_OK) return rc;
    if (logits_out != nullptr) B200_CUDA(cudaMemcpyAsync(logits_out, e->logits, vbytes, cudaMemcpyDeviceToDevice, st));
    e->h_gen += 1;
    e->h_pos += S;
    return B200_OK;
  }
  for (int64_t t = 0; t < S; ++t) {
    B200_CUDA(cudaMemcpyAsync(e->cur_tok, ids + t, 8, cudaMemcpyDeviceToDevice, st));
    const bool head = (t == S - 1) || (logits_mode == 1 && logits_out != nullptr);
    int rc = engine_run_token(e, st, head);
    if (rc != B200_OK) return rc;
    if (head && logits_out != nullptr) {
      uint8_t* dst = (uint8_t*)logits_out + (logits_mode == 1 ? (size_t)t * vbytes : 0);
      B200_CUDA(cudaMemcpyAsync(dst, e->logits, vbytes, cudaMemcpyDeviceToDevice, st));
    }
    if (head) e->h_gen += 1;
  }
  e->h_pos += S;
  return B200_OK;
}

int b200_engine_decode(b200_engine* e, int64_t n_steps, int64_t* tokens_out, void* stream) {
  using namespace b200;
  B200_CHECK_ARG(e && n_steps >= 0, "engine_decode: bad argument");
  B200_CHECK_ARG(n_steps <= e->gen_cap, "engine_decode: at most %d steps per call", e->gen_cap);
  if (e->h_pos + n_steps > e->d.max_ctx) {
    set_error("engine_decode: position %lld + %lld steps exceeds max_ctx %d", (long long)e->h_pos, (long long)n_steps,
              e->d.max_ctx);
    return B200_ERR_STATE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (e->batch > 1) {   // tokens_out [n_steps][B]
    const int64_t B = e->batch, first_b = e->hb_gen;
    for (int64_t i = 0; i < n_steps; ++i) {
      int rc = engine_run_token_batch(e, st, true);
      if (rc != B200_OK) return rc;
    }
    e->hb_gen += n_steps;
    e->h_pos += n_steps;
    if (tokens_out != nullptr && n_steps > 0) {
      const int64_t a = first_b % e->gen_cap;
      const int64_t n1 = std::min<int64_t>(n_steps, e->gen_cap - a);
      B200_CUDA(cudaMemcpyAsync(tokens_out, e->b_log + a * B, (size_t)n1 * B * 8, cudaMemcpyDeviceToDevice, st));
      if (n1 < n_steps)
        B200_CUDA(cudaMemcpyAsync(tokens_out + n1 * B, e->b_log, (size_t)(n_steps - n1) * B * 8, cudaMemcpyDeviceToDevice, st));
    }
    return B200_OK;
  }
  const int64_t first = e->h_gen;
  for (int64_t i = 0; i < n_steps; ++i) {
    int rc = engine_run_token(e, st, true);
    if (rc != B200_OK) return rc;
  }
  e->h_gen += n_steps;
  e->h_pos += n_steps;
  if (tokens_out != nullptr && n_steps > 0) {
    const int64_t a = first % e->gen_cap;
    const int64_t n1 = std::min<int64_t>(n_steps, e->gen_cap - a);
    B200_CUDA(cudaMemcpyAsync(tokens_out, e->gen_log + a, (size_t)n1 * 8, cudaMemcpyDeviceToDevice, st));
    if (n1 < n_steps)
      B200_CUDA(cudaMemcpyAsync(tokens_out + n1, e->gen_log, (size_t)(n_steps - n1) * 8, cudaMemcpyDeviceToDevice, st));
  }
  return B200_OK;
}

int b200_engine_last_token(b200_engine* e, int64_t* token_out, void* stream) {
  using namespace b200;
  B200_CHECK_ARG(e && token_out, "engine_last_token: null argument");
  if (e->batch > 1) {   // token_out [B]
    B200_CUDA(cudaMemcpyAsync(token_out, e->b_tok, (size_t)e->batch * 8, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return B200_OK;
  }
  B200_CUDA(cudaMemcpyAsync(token_out, e->cur_tok, 8, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return B200_OK;
}

int64_t b200_engine_position(const b200_engine* e) { return e ? e->h_pos : -1; }
int64_t b200_engine_generated(const b200_engine* e) { return e ? e->h_gen : -1; }

int b200_engine_set_sampler(b200_engine* e, float temperature, int64_t top_k, float top_p, float min_p, uint64_t seed,
                            void* stream) {
  using namespace b200;
  B200_CHECK_ARG(e, "engine_set_sampler: null engine");
  B200_CHECK_ARG(temperature >= 0.f && top_p >= 0.f && min_p >= 0.f && min_p <= 1.f && top_k >= 0,
                 "engine_set_sampler: parameter out of range");
  if (e->tp_world > 1) {
    set_error("engine_set_sampler: tensor-parallel engines pick the token from vocabulary shards; sampler not built");
    return B200_ERR_UNSUPPORTED;
  }
  // Sampler::Sampler (src/engine/Sampler.cpp:14-21): sampling is on as soon as any of the four knobs is set
  const bool on = temperature > 0.f || top_k > 0 || top_p < 1.f || min_p > 0.f;
  B200_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  e->sampler_on = on;
  e->s_temperature = temperature;
  e->s_top_k = top_k;
  e->s_top_p = top_p;
  e->s_min_p = min_p;
  e->s_seed = seed;
  if (e->use_graph) {
    cudaGraphExec_t fresh = nullptr;
    int rc = engine_capture(e, true, &fresh);
    if (rc != B200_OK) return rc;
    if (e->g_step) cudaGraphExecDestroy(e->g_step);
    e->g_step = fresh;
  }
  return B200_OK;
}

int b200_engine_set_mailbox(b200_engine* e, uint64_t* ring, int64_t capacity, void* stream) {
  using namespace b200;
  B200_CHECK_ARG(e, "engine_set_mailbox: null engine");
  B200_CHECK_ARG(ring == nullptr || (capacity >= 2 && (reinterpret_cast<uintptr_t>(ring) & 7) == 0),
                 "engine_set_mailbox: ring must be 8-byte aligned with capacity >= 2");
  if (e->tp_world > 1) {
    set_error("engine_set_mailbox: tensor-parallel engines publish through tp_finish_kernel; mailbox not built");
    return B200_ERR_UNSUPPORTED;
  }
  if (ring != nullptr) {  // the pointer must be addressable from the device
    cudaPointerAttributes at{};
    cudaError_t err = cudaPointerGetAttributes(&at, ring);
    if (err != cudaSuccess || at.type != cudaMemoryTypeHost || at.devicePointer == nullptr) {
      (void)cudaGetLastError();
      set_error("engine_set_mailbox: ring is not pinned, device-mapped host memory");
      return B200_ERR_INVALID;
    }
    ring = static_cast<uint64_t*>(at.devicePointer);
  }
  // graphs in flight keep the old parameters: drain the stream, then re-capture with the new ring baked in
  B200_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  e->mailbox = reinterpret_cast<unsigned long long*>(ring);
  e->mailbox_cap = ring ? (unsigned long long)capacity : 1ull;
  if (e->use_graph) {
    cudaGraphExec_t fresh = nullptr;
    int rc = engine_capture(e, true, &fresh);
    if (rc != B200_OK) return rc;
    if (e->g_step) cudaGraphExecDestroy(e->g_step);
    e->g_step = fresh;
  }
  return B200_OK;
}

int64_t b200_engine_debug_trace(b200_engine* e, uint64_t* out_host, int64_t max_entries) {
  if (!e || !e->trace || !out_host) return 0;
  const int64_t n = std::min<int64_t>(max_entries, 5 * e->d.layers + 1);
  if (cudaMemcpy(out_host, e->trace, (size_t)n * 8 * 8, cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
  return n;
}
int64_t b200_engine_launches_per_token(const b200_engine* e) {
  if (!e) return -1;
  return e->batch > 1 && e->b_cap == e->batch ? e->b_launches : e->launches_per_token;   // the graph the next step runs
}
int64_t b200_engine_options(const b200_engine* e) {
  if (!e) return -1;
  return (e->use_graph ? 1 : 0) | (e->use_pdl ? 2 : 0) | (e->use_prefill_gemm ? 8 : 0);
}

int64_t b200_engine_bytes_per_token(const b200_engine* e, int64_t ctx) {
  if (!e) return -1;
  const b200_model_desc& d = e->d;
  const int64_t per_layer = (int64_t)(e->qdim + 2 * e->kvdim) * d.hidden + (int64_t)e->qdim * d.hidden +
                            3ll * e->I_l * d.hidden;
  const int64_t weights = 2 * ((int64_t)d.layers * per_layer + (int64_t)e->V_l * d.hidden);
  const int64_t kv = 4ll * d.layers * e->kvdim * ctx + 4ll * d.layers * e->kvdim;
  return weights + kv;
}

}  // extern "C"
