// mega.cuh — the persistent whole-token kernel (mega.cu): host-side description of a token as a list of ops.
#pragma once
#include "gemv.cuh"

namespace b200 {

// A token = embed (in-kernel gather) + L × {qkv GEMV, attention, o_proj GEMV, gate|up GEMV, down GEMV} + lm_head GEMV.
// Every CTA of the one persistent kernel walks the same list; a CTA takes part in an op only when the op's row blocks
// (or attention work items), dealt round-robin from `cta_off`, give it work.
enum MegaKind : int { MK_GEMV = 0, MK_ATTN = 1 };
// GEMV epilogue that only exists inside the whole-token kernel: bf16 logits to HBM + running greedy argmax
constexpr int EPI_LOGITS = 4;

// Activation vectors travel between CTAs as 4-byte words {tag : 16 | bf16 value : 16} ("LL" words: data and flag in
// one store, so a consumer that sees the tag has the value — no counter, no fence, one L2 round trip per dependency).
struct MegaOp {
  int kind;
  // ---- GEMV
  int rpw, nseg, pro, epi;
  int n, k, k_pad, seg_rows, rowblocks, ksteps, cta_off, tmap;
  int x_ev;        // event (op index within the token) that produced x_ll's current content; -1: x is the embedding row
  int res_ev;      // likewise for res_ll; -1: residual is the embedding row
  float eps;
  const __nv_bfloat16* norm_w;
  const __nv_bfloat16* bias;
  const uint32_t* x_ll;
  const uint32_t* res_ll;
  uint32_t* y_ll;            // null for EPI_LOGITS
  // Every LL vector exists in `rep` copies `rstride` words apart: writers store all copies, CTA c reads copy c % rep,
  // so that at most grid / rep CTAs poll the same L2 lines (measured: tools/micro/ll_bench.cu)
  int x_rep, x_rstride, res_rep, res_rstride, y_rep, y_rstride;
  // Long vectors (k > 2048: HBM-bound models) are not polled word by word by every thread of every CTA — that floods L2
  // under the weight stream.  Their producers count themselves done on a per-op arrival counter (fence + red.add once
  // per CTA) and ONE thread per consumer CTA polls that counter with back-off before the vector is read.
  int x_ctr;        // op index whose counter gates x_ll, or -1: poll the words
  int x_ctr_count;  // CTAs that arrive on it per token
  int y_ctr;        // 1: this op's CTAs arrive on counter[op index] after their last store
  __nv_bfloat16* y_plain;    // EPI_LOGITS: bf16 logits [n]
  // ---- attention
  int layer;
  int heads_per_item;        // query heads (of one KV head) a work item handles
  const __nv_bfloat16* q_norm;
  const __nv_bfloat16* k_norm;
};

struct MegaParams {
  const MegaOp* ops;
  const CUtensorMap* tmaps;
  int n_ops;                 // ops per token in THIS launch (without the lm_head when with_head == 0)
  int events_per_token;      // tag stride per token (constant for the engine, whatever with_head says)
  int n_tokens;
  int with_head;
  unsigned long long tok_seq0;   // tokens this engine's whole-token kernel has processed before this launch
  const __nv_bfloat16* embed;
  int V, H;
  const float* rope;
  int hd, Hq, Hkv, max_ctx;
  float eps;
  __nv_bfloat16* kcache;
  __nv_bfloat16* vcache;
  size_t kv_layer_stride;        // elements per layer
  const uint32_t* qkv_ll;
  uint32_t* attn_ll;
  int qkv_rep, qkv_rstride, attn_rep, attn_rstride;
  float* attn_ws;                // split partials [Hq][nsplit][hd + 2]
  unsigned int* attn_tickets;    // [Hq / heads_per_item], zero-initialised, self-resetting
  int nsplit;
  int* pos;
  int64_t* cur_tok;
  int64_t* gen_log;
  unsigned long long* gen_count;
  int gen_cap;
  unsigned long long* mailbox;
  unsigned long long mailbox_cap;
  uint2* cand;                   // [grid][2] {value bits, tag}, {index, tag}
  int stages;
  int xs_elems;                  // capacity of the activation staging vector (bf16 elements, multiple of 1024)
  unsigned long long* ctr;       // [ops per token][16] arrival counters (one 128-byte line each), monotone
  unsigned long long* trace;     // debug: [cta][op + 1][4] %globaltimer stamps for the LAST token of the launch
  int trace_ops;                 // row length of `trace` in ops (5 L + 2)
};

struct MegaPlan {
  MegaParams p{};
  int grid = 0;
  int smem = 0;
  int n_ops_body = 0;            // ops without the lm_head
  void* dev_blob = nullptr;      // ops + tensor maps (one cudaMalloc)
};

int mega_setup_attributes();
size_t mega_smem_fixed(int xs_elems, int H);   // shared memory the kernel needs besides the ring (ring stage: 16 KB + 16 B)
constexpr int kMegaStageBytes = 16 * 1024;
constexpr int kMegaMaxSmem = 227 * 1024;
int mega_launch(const MegaPlan& plan, int n_tokens, bool with_head, unsigned long long tok_seq0, cudaStream_t st);

// shape choice shared with the per-op GEMV (gemv.cu)
struct GemvShape {
  int rpw, box_r, kb, k_pad, stage_bytes;
  int64_t rbs;
};
GemvShape gemv_pick_shape(int64_t n, int64_t k, int nseg, int num_sms, int max_rpw);

}  // namespace b200
