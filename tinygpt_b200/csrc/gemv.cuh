// gemv.cuh — host-side plan/launch interface of the weight-streaming bf16 GEMV (see gemv.cu).
#pragma once
#include "common.cuh"

namespace b200 {

// Prologue: how the activation vector x[k] gets into shared memory.
//   PRO_PLAIN       x is read as is
//   PRO_RMSNORM     x ← bf16(x * rsqrt(mean(x²)+eps) * w)              (fused input/post-attention/final RMSNorm)
//   PRO_TP_RMSNORM  h ← bf16(residual + bf16(Σ_r partial_r)); x ← rmsnorm(h); CTA 0 stores h  (tensor parallel)
enum GemvPro : int { PRO_PLAIN = 0, PRO_RMSNORM = 1, PRO_TP_RMSNORM = 2 };
// Epilogue: what happens to the fp32 row sums.
//   EPI_PLAIN     y = bf16(acc); with bias: y = bf16(y + bias)         (reference: second rounding, 3-D Linear input)
//   EPI_RESIDUAL  y = bf16(residual + bf16(acc))                        (o_proj / down_proj + DecoderLayer add)
//   EPI_SILU_MUL  y = bf16(bf16(silu(bf16(acc_gate))) * bf16(acc_up))   (merged gate|up projection + SiLUMul)
//   EPI_TP_PUSH   fp32 acc stored into every tensor-parallel peer's exchange window + arrival flag
enum GemvEpi : int { EPI_PLAIN = 0, EPI_RESIDUAL = 1, EPI_SILU_MUL = 2, EPI_TP_PUSH = 3 };

constexpr int kMaxTpWorld = 8;
constexpr int kMaxBatch = 8;   // sequences one batched decode step serves (gemv_batch.cu: W streamed once for all of them)

// kernel geometry shared by gemv.cu and gemv_batch.cu
namespace gemvk {
constexpr int kNW = 8;                      // consumer warps
constexpr int kThreads = (kNW + 1) * 32;    // + 1 producer warp
constexpr int kBoxK = 256;                  // columns per 512-byte row piece
constexpr int kRowBytes = kBoxK * 2;        // 512
constexpr int kConsumers = kNW * 32;
constexpr int kMaxStages = 32;
// 256-column pieces per pipeline stage: every stage carries 16 KB (32 KB for RPW 4 × 2 segments)
__host__ __device__ constexpr int kboxes(int rpw, int nseg) { return (4 / (rpw * nseg)) > 0 ? 4 / (rpw * nseg) : 1; }
}  // namespace gemvk

struct GemvParams {
  const __nv_bfloat16* x;         // [k] activation
  const __nv_bfloat16* norm_w;    // [k] RMSNorm weight or null
  float eps;
  const __nv_bfloat16* bias;      // [n] or null
  const __nv_bfloat16* residual;  // [n]; may alias y
  __nv_bfloat16* y;               // [n]
  // batched decode (gemv_batch.cu): `batch` sequences, x / residual / y of sequence b at b·stride elements
  int batch;                      // 0 or 1: the single-sequence kernels
  int x_stride, y_stride;         // = k, n (residual uses y_stride)
  int n;                          // output rows (per segment)
  int k;                          // reduction length
  int k_pad;                      // k rounded up to the 256-element TMA box
  int seg_rows;                   // row offset of segment 1 inside W (NSEG == 2: gate rows | up rows)
  int rowblocks;                  // ceil(n / (8 * RPW))
  int stages;                     // depth of the shared-memory ring
  unsigned long long* trace;      // debug (B200_TRACE=1): CTA 0 stores globaltimer at entry / after the PDL wait / at exit
  int* pos_inc;                   // engine: when set, CTA 0 advances the token position after its last row block
  // ---- tensor-parallel exchange (peer-mapped windows over NVLink).  Low-latency protocol: every fp32 partial travels
  // as ONE 8-byte store {value bits, tag} with tag = token epoch + 1, so data and "flag" arrive atomically together —
  // no system fence, no remote atomic, no separate flag poll (the idea of NCCL's LL protocol).
  uint2* tp_push[kMaxTpWorld];            // EPI_TP_PUSH: slot for MY partial inside rank r's window ([n] × 8 bytes)
  const uint2* tp_partials;               // PRO_TP_RMSNORM: local window, tp_world vectors, tp_stride words apart
  int tp_stride;
  const unsigned long long* tp_epoch;     // device counter: tokens completed (stable during a token)
  const __nv_bfloat16* tp_residual;       // PRO_TP_RMSNORM: hidden state before the add
  __nv_bfloat16* tp_h_out;                // PRO_TP_RMSNORM: CTA 0 stores the new hidden state (≠ tp_residual)
  int tp_world;
};

struct GemvPlan {
  CUtensorMap tmap;
  GemvParams p;
  int rpw;   // rows per consumer warp: 1, 2 or 4
  int nseg;  // 1, or 2 for merged gate|up
  int pro;
  int epi;
  int grid;
  int smem;
  int batch;   // > 1: gemv_batch_kernel (tensor cores, gemv_batch.cu; gemv_plan_set_batch)
  int sub;     // sequences per launch when `batch` activation vectors do not fit beside the ring (k = 14336 at B = 8:
               // two launches of 4, W streamed twice); 0 = all of them in one launch
};

// Plan a GEMV over W[rows_total, k] (row-major bf16).  `n` = rows produced (per segment).
// `smem_budget` bounds the dynamic shared memory of the launch (ring depth): consecutive PDL-chained kernels co-reside
// on an SM, so the engine hands out budgets whose pairwise sums fit in 227 KB (gemv_smem_wanted tells how much a
// kernel could use to hold ALL of its busiest CTA's weights).
constexpr int kGemvDefaultSmem = 144 * 1024;  // 8 stages + the activation vector: the ring depth that saturates HBM
constexpr int kGemvMaxSmem = 200 * 1024;
int gemv_make_plan(GemvPlan* plan, const void* W, int64_t rows_total, int64_t n, int64_t k, int nseg, int pro, int epi,
                   int num_sms, int smem_budget = kGemvDefaultSmem);
int gemv_smem_wanted(int64_t n, int64_t k, int nseg, int num_sms);
int gemv_launch(const GemvPlan& plan, cudaStream_t stream, bool pdl);
// Turn a single-sequence plan into a batched one: B ≤ kMaxBatch activation vectors staged side by side (shared memory
// taken from the ring; `sub` < B sequences per launch when they do not fit), x / y / residual strides set to k / n.
// Single-GPU prologues / epilogues only.
int gemv_plan_set_batch(GemvPlan* plan, int B);
int gemv_batch_launch(const GemvPlan& plan, cudaStream_t stream, bool pdl);   // gemv_batch.cu
int gemv_batch_setup_attributes();
int gemv_batch_fixed_smem(const GemvPlan& plan, int nb);   // shared memory beside the ring for nb sequences
int gemv_setup_attributes();  // cudaFuncSetAttribute(max dynamic smem) for every instantiation, once per process

}  // namespace b200
