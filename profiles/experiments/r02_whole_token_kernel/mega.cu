// mega.cu — the whole decode token (and, looped, a whole run of tokens) as ONE persistent kernel for sm_100a.
//
// Why: the per-op engine (engine.cu) runs a token as 5·L + 3 PDL-chained kernels.  For the small models every one of
// those grid-wide dependencies costs ≈ 1 µs of launch/drain/flush latency on top of a 1–4 µs body against ≈ 1 µs of HBM
// time (profiles/r02_*: Qwen2.5-0.5B 0.37 of the HBM roofline), and for the large ones HBM idles across every
// boundary.  Here 148 CTAs (one per SM) stay resident for the whole token:
//   * warp 8 of every CTA is a TMA producer that streams THIS CTA's share of every weight matrix of the token, in op
//     order, through one shared-memory ring (cp.async.bulk.tensor.2d → SASS UTMALDG, full/empty mbarriers).  It never
//     waits for an activation, so HBM keeps streaming across op, layer and token boundaries; the ring (≈ 190 KB per
//     SM, 28 MB chip-wide) is what decouples it from the consumers.
//   * warps 0–7 consume: for every op they wait for the activation vector, run the (fused) prologue, dot their rows
//     out of the ring, run the (fused) epilogue and publish their rows.
//   * Activations cross CTAs as 4-byte words {tag:16 | bf16:16} in L2 ("LL" protocol: data and flag in one store).  A
//     consumer polls the vector itself until every word carries this op's tag: one L2 round trip per dependency, no
//     counter, no fence, no kernel boundary.  The tag is (token sequence · events per token + op index) mod 65535 + 1;
//     every word of a buffer is rewritten by every event on it, so a stale tag can never match.
//   * Attention runs inside the same kernel on the CTAs its (head, KV split) items are dealt to; K/V rows of earlier
//     tokens are prefetched into registers before the q/k/v vector of this token is polled.
//   * lm_head keeps a running greedy argmax (last index wins, like the reference); the 148 CTA candidates are
//     exchanged as LL words and every CTA picks the winner itself, so the next token starts without another sync.
// Numerics: the GEMV inner loops, prologues and epilogues are the per-op kernels' (gemv.cu) — same FMA order per row,
// same rounding points — so logits are bit-identical to the per-op engine whenever attention is (attention merges
// its per-thread online-softmax partials in a different order: within the summation-order floor, see tests).
//   [ref: src/model/GPTModel.h:51-58; src/layer/DecoderLayer.h:38-43; src/layer/Attention.h:71-112,156-163;
//    src/layer/GatedMLP.h:37-41; src/engine/GPTEngine.cpp:94-99,154-174; src/engine/Sampler.cpp:23-29]
#include "mega.cuh"

#include <algorithm>
#include <mutex>

namespace b200 {

namespace {

constexpr int kNW = 8;                       // consumer warps
constexpr int kConsumers = kNW * 32;
constexpr int kThreads = kConsumers + 32;    // + producer warp
constexpr int kBoxK = 256;
constexpr int kRowBytes = kBoxK * 2;
constexpr int kStageBytes = 16 * 1024;
constexpr float kLog2e = 1.4426950408889634f;
constexpr int kMaxItemHeads = 8;             // query heads per attention work item
constexpr unsigned long long kTrapNs = 4000000000ull;

__device__ __forceinline__ uint4 ld_volatile_u4(const uint4* p) {
  uint4 v;
  asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint2 ld_volatile_u2(const uint2* p) {
  uint2 v;
  asm volatile("ld.volatile.global.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t ld_volatile_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_volatile_u32(uint32_t* p, uint32_t v) {
  asm volatile("st.volatile.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void st_volatile_u2(uint2* p, uint2 v) {
  asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(v.x), "r"(v.y) : "memory");
}

// A spin that sees no progress for ~4 s is a bug (or a CTA that is not resident): fail loudly, never hang the GPU.
struct SpinGuard {
  unsigned int n = 0;
  unsigned long long t0 = 0;
  __device__ __forceinline__ void tick() {
    if ((++n & 0x3fffu) == 0) {
      const unsigned long long now = global_timer_ns();
      if (t0 == 0) t0 = now;
      if (now - t0 > kTrapNs) __trap();
    }
  }
};

__device__ __forceinline__ uint32_t ll_tag(unsigned long long tok_seq, int events, int ev) {
  return (uint32_t)((tok_seq * (unsigned long long)events + (unsigned long long)ev) % 65535ull) + 1u;
}
__device__ __forceinline__ uint32_t ll_word(uint32_t tag, __nv_bfloat16 v) {
  return (tag << 16) | (uint32_t)__bfloat16_as_ushort(v);
}
// 8 consecutive LL words → 8 packed bf16 (spins until all carry `tag`; both 16-byte loads are in flight together; a
// failed round backs off `backoff_ns` so that idle CTAs do not flood L2 while the weight stream runs through it)
__device__ __forceinline__ uint4 ll_poll8(const uint32_t* p, uint32_t tag, unsigned int backoff_ns) {
  SpinGuard g;
  for (;;) {
    const uint4 a = ld_volatile_u4(reinterpret_cast<const uint4*>(p));
    const uint4 b = ld_volatile_u4(reinterpret_cast<const uint4*>(p) + 1);
    const uint32_t all = ((a.x >> 16) ^ tag) | ((a.y >> 16) ^ tag) | ((a.z >> 16) ^ tag) | ((a.w >> 16) ^ tag) |
                         ((b.x >> 16) ^ tag) | ((b.y >> 16) ^ tag) | ((b.z >> 16) ^ tag) | ((b.w >> 16) ^ tag);
    if (all == 0u)
      return make_uint4((a.x & 0xffffu) | (a.y << 16), (a.z & 0xffffu) | (a.w << 16), (b.x & 0xffffu) | (b.y << 16),
                        (b.z & 0xffffu) | (b.w << 16));
    if (backoff_ns) __nanosleep(backoff_ns);
    g.tick();
  }
}
__device__ __forceinline__ __nv_bfloat16 ll_poll1(const uint32_t* p, uint32_t tag) {
  SpinGuard g;
  for (;;) {
    const uint32_t v = ld_volatile_u32(p);
    if ((v >> 16) == tag) return __ushort_as_bfloat16((unsigned short)(v & 0xffffu));
    g.tick();
  }
}

__device__ __forceinline__ void mbar_wait_guarded(uint64_t* bar, uint32_t parity) {
  SpinGuard g;
  while (!mbar_try_wait(bar, parity)) g.tick();
}

struct Ring {
  int s;
  uint32_t ph;
};
__device__ __forceinline__ void ring_next(Ring& r, int stages) {
  if (++r.s == stages) {
    r.s = 0;
    r.ph ^= 1u;
  }
}

// greedy argmax candidate; the reference's CUDA rule: fp32 compare, the HIGHEST index wins a tie
// [ref: third_party/TinyTorch/src/Operation/OpReduceCuda.cuh:145-156]
struct Best {
  float v;
  int i;
};
__device__ __forceinline__ void best_merge(Best& b, float v, int i) {
  if (v > b.v || (v == b.v && i > b.i)) {
    b.v = v;
    b.i = i;
  }
}

// shared-memory carve-up (dynamic; no static __shared__ in this kernel)
struct Smem {
  uint8_t* ring;
  __nv_bfloat16* xs;     // staged activation vector of the current op (normalised where the op has a norm prologue)
  __nv_bfloat16* emb;    // embedding row of the current token (layer 0's x and residual)
  uint64_t* full;
  uint64_t* empty;
  float* red;            // [16]
  float* q_s;            // [kMaxItemHeads][128] fp32 rotated queries
  __nv_bfloat16* knew;   // [128] this token's K row (after norm + RoPE)
  __nv_bfloat16* vnew;   // [128]
  float* ared;           // [kNW][16 parts][10] per-warp attention partials {m, l, acc[8]}
  float* cbv;            // [kNW] argmax candidates of the warps
  int* cbi;              // [kNW]
  int* flag;             // [4] is_last / winner
};
__host__ __device__ inline size_t smem_fixed_bytes(int xs_elems, int H) {
  return (size_t)xs_elems * 2 + (size_t)((H * 2 + 15) / 16 * 16) + 16 * 4 + kMaxItemHeads * 128 * 4 + 2 * 128 * 2 +
         kNW * 16 * 10 * 4 + kNW * 8 + 16 + 64;
}
__device__ __forceinline__ Smem carve(uint8_t* base, const MegaParams& P) {
  Smem s;
  uint8_t* p = base;
  s.ring = p;
  p += (size_t)P.stages * kStageBytes;
  s.xs = reinterpret_cast<__nv_bfloat16*>(p);
  p += (size_t)P.xs_elems * 2;
  s.emb = reinterpret_cast<__nv_bfloat16*>(p);
  p += (size_t)((P.H * 2 + 15) / 16 * 16);
  s.full = reinterpret_cast<uint64_t*>(p);
  p += (size_t)P.stages * 8;
  s.empty = reinterpret_cast<uint64_t*>(p);
  p += (size_t)P.stages * 8;
  s.red = reinterpret_cast<float*>(p);
  p += 16 * 4;
  s.q_s = reinterpret_cast<float*>(p);
  p += kMaxItemHeads * 128 * 4;
  s.knew = reinterpret_cast<__nv_bfloat16*>(p);
  p += 128 * 2;
  s.vnew = reinterpret_cast<__nv_bfloat16*>(p);
  p += 128 * 2;
  s.ared = reinterpret_cast<float*>(p);
  p += kNW * 16 * 10 * 4;
  s.cbv = reinterpret_cast<float*>(p);
  p += kNW * 4;
  s.cbi = reinterpret_cast<int*>(p);
  p += kNW * 4;
  s.flag = reinterpret_cast<int*>(p);
  return s;
}

__device__ __forceinline__ void cbar() { named_bar_sync(1, kConsumers); }
__device__ __forceinline__ unsigned long long* trace_slot(const MegaParams& P, int op_index, int what) {
  return P.trace + ((size_t)blockIdx.x * P.trace_ops + op_index) * 4 + what;
}
// arrival counter of an op: one thread per producer CTA, after a CTA barrier that follows the CTA's last store
__device__ __forceinline__ void ctr_arrive(const MegaParams& P, int op_index) {
  __threadfence();
  red_release_gpu_add_u64(P.ctr + (size_t)op_index * 16, 1ull);
}
__device__ __forceinline__ void ctr_wait(const MegaParams& P, int op_index, unsigned long long target) {
  SpinGuard g;
  while (ld_acquire_gpu_u64(P.ctr + (size_t)op_index * 16) < target) {
    __nanosleep(100);
    g.tick();
  }
}

// ------------------------------------------------------------------------------------------------ GEMV rows
// The k loop, reduction and epilogue of gemv_stream_kernel (gemv.cu), unchanged in arithmetic: warp w owns rows
// w·RPW … of every 8·RPW-row block, two fp32 accumulators per row, xor-shuffle tree, lane r finishes row r.
// Everything the row loop needs from the op descriptor, read ONCE into registers: the descriptor lives in global memory
// and the mbarrier asm statements are memory clobbers, so a reference would be re-loaded in every k step.
struct RowArgs {
  int n, rowblocks, ksteps, epi, res_ev, y_rep, y_rstride, stages;
  const __nv_bfloat16* bias;
  const uint32_t* res_ll;   // already offset to this CTA's copy
  uint32_t* y_ll;
  __nv_bfloat16* y_plain;
};
template <int RPW, int NSEG>
__device__ __forceinline__ void gemv_rows(const RowArgs op, const Smem& sm, Ring& ring, int first_rb,
                                          int grid, uint32_t tag_y, uint32_t tag_res, Best& best, int warp, int lane) {
  constexpr int KB = (4 / (RPW * NSEG)) > 0 ? 4 / (RPW * NSEG) : 1;
  constexpr int kBoxR = kNW * RPW;
  constexpr int kBoxBytes = kBoxR * kRowBytes;
  static_assert(KB * NSEG * kBoxBytes == kStageBytes, "every stage of the unified ring is 16 KB");
  const uint8_t* const my_rows = sm.ring + (size_t)(warp * RPW) * kRowBytes + lane * 16;
  for (int rb = first_rb; rb < op.rowblocks; rb += grid) {
    const int row_base = rb * kBoxR + warp * RPW;
    const int row = row_base + lane;
    const bool mine = lane < RPW && row < op.n;
    // epilogue operands are requested now so that their latency hides behind the k loop
    __nv_bfloat16 res_v = f_to_bf16(0.f), bias_v = f_to_bf16(0.f);
    if (mine) {
      if (op.epi == EPI_RESIDUAL) res_v = (op.res_ev < 0) ? sm.emb[row] : ll_poll1(op.res_ll + row, tag_res);
      if (op.epi == EPI_PLAIN && op.bias != nullptr) bias_v = op.bias[row];
    }
    float acc[NSEG][RPW], acc_b[NSEG][RPW];
#pragma unroll
    for (int seg = 0; seg < NSEG; ++seg)
#pragma unroll
      for (int r = 0; r < RPW; ++r) acc[seg][r] = acc_b[seg][r] = 0.f;
    for (int ks = 0; ks < op.ksteps; ++ks) {
      uint4 xq[KB];
#pragma unroll
      for (int kb = 0; kb < KB; ++kb)
        xq[kb] = *reinterpret_cast<const uint4*>(sm.xs + (ks * KB + kb) * kBoxK + lane * 8);
      mbar_wait_guarded(&sm.full[ring.s], ring.ph);
      const uint8_t* st = my_rows + (size_t)ring.s * kStageBytes;
      uint4 wv[KB][NSEG][RPW];
#pragma unroll
      for (int kb = 0; kb < KB; ++kb)
#pragma unroll
        for (int seg = 0; seg < NSEG; ++seg)
#pragma unroll
          for (int r = 0; r < RPW; ++r)
            wv[kb][seg][r] = *reinterpret_cast<const uint4*>(st + (kb * NSEG + seg) * kBoxBytes + r * kRowBytes);
#pragma unroll
      for (int kb = 0; kb < KB; ++kb) {
        float xf[8];
        unpack8(xq[kb], xf);
#pragma unroll
        for (int seg = 0; seg < NSEG; ++seg)
#pragma unroll
          for (int r = 0; r < RPW; ++r) dot8x2(wv[kb][seg][r], xf, acc[seg][r], acc_b[seg][r]);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&sm.empty[ring.s]);
      ring_next(ring, op.stages);
    }
#pragma unroll
    for (int seg = 0; seg < NSEG; ++seg)
#pragma unroll
      for (int r = 0; r < RPW; ++r) acc[seg][r] = warp_sum(acc[seg][r] + acc_b[seg][r]);
    float a0 = acc[0][0], a1 = acc[NSEG - 1][0];
#pragma unroll
    for (int r = 1; r < RPW; ++r) {
      if (lane == r) {
        a0 = acc[0][r];
        a1 = acc[NSEG - 1][r];
      }
    }
    if (mine) {
      __nv_bfloat16 v = f_to_bf16(a0);
      if (op.epi == EPI_PLAIN) {
        if (op.bias != nullptr) v = __hadd(v, bias_v);   // the reference's second rounding (separate add kernel)
      } else if (op.epi == EPI_RESIDUAL) {
        v = __hadd(res_v, v);
      } else if (op.epi == EPI_SILU_MUL) {
        const float g = round_bf16(a0);
        const __nv_bfloat16 sg = f_to_bf16(g / (1.f + expf(-g)));
        v = __hmul(sg, f_to_bf16(a1));
      }
      if (op.epi == EPI_LOGITS) {
        op.y_plain[row] = v;
        best_merge(best, bf16_to_f(v), row);
      } else {
        const uint32_t word = ll_word(tag_y, v);
        for (int c = 0; c < op.y_rep; ++c) st_volatile_u32(op.y_ll + (size_t)c * op.y_rstride + row, word);
      }
    }
  }
}

// One GEMV op of the token for this CTA's consumers: stage x (poll + fused RMSNorm), then the rows.
__device__ __forceinline__ void gemv_op(const MegaOp* __restrict__ opp, const MegaParams& P, const Smem& sm, Ring& ring,
                                        unsigned long long tok_seq, int op_index, Best& best, int cta, int grid,
                                        int ctid, int warp, int lane, bool trace) {
  const int rowblocks = opp->rowblocks;
  int cprime = cta - opp->cta_off;
  if (cprime < 0) cprime += grid;
  if (cprime >= rowblocks) return;   // no rows of this op for this CTA (uniform over the CTA)
  const int k = opp->k, k_pad = opp->k_pad, pro = opp->pro, x_ev = opp->x_ev;
  RowArgs ra;
  ra.n = opp->n;
  ra.rowblocks = rowblocks;
  ra.ksteps = opp->ksteps;
  ra.epi = opp->epi;
  ra.res_ev = opp->res_ev;
  ra.y_rep = opp->y_rep;
  ra.y_rstride = opp->y_rstride;
  ra.stages = P.stages;
  ra.bias = opp->bias;
  ra.res_ll = opp->res_ll != nullptr ? opp->res_ll + (size_t)(cta % opp->res_rep) * opp->res_rstride : nullptr;
  ra.y_ll = opp->y_ll;
  ra.y_plain = opp->y_plain;
  const uint32_t tag_x = x_ev >= 0 ? ll_tag(tok_seq, P.events_per_token, x_ev) : 0u;
  const uint32_t tag_res = ra.res_ev >= 0 ? ll_tag(tok_seq, P.events_per_token, ra.res_ev) : 0u;
  const uint32_t tag_y = ll_tag(tok_seq, P.events_per_token, op_index);
  const uint32_t* x_ll = x_ev >= 0 ? opp->x_ll + (size_t)(cta % opp->x_rep) * opp->x_rstride : nullptr;
  const unsigned int backoff = k > 2048 ? 200u : 0u;   // long vectors belong to HBM-bound models: the ring hides the wait
  const int nvec = k >> 3, nvec_pad = k_pad >> 3;
  uint4* const xv = reinterpret_cast<uint4*>(sm.xs);
  if (opp->x_ctr >= 0) {   // gated by the producers' arrival counter: one poller per CTA, then one pass over the words
    if (ctid == 0) ctr_wait(P, opp->x_ctr, (tok_seq + 1ull) * (unsigned long long)opp->x_ctr_count);
    cbar();
  }
  // ---- activation vector: 8 elements per thread and step, the per-op kernels' partition (same sum-of-squares order)
  float ss = 0.f;
  for (int i = ctid; i < nvec_pad; i += kConsumers) {
    uint4 q = make_uint4(0, 0, 0, 0);
    if (i < nvec) {
      q = (x_ev < 0) ? reinterpret_cast<const uint4*>(sm.emb)[i] : ll_poll8(x_ll + 8 * i, tag_x, backoff);
      if (pro == PRO_RMSNORM) {
        float xf[8];
        unpack8(q, xf);
#pragma unroll
        for (int e = 0; e < 8; ++e) ss += xf[e] * xf[e];
      }
    }
    xv[i] = q;
  }
  if (pro == PRO_RMSNORM) {
    ss = warp_sum(ss);
    if (lane == 0) sm.red[warp] = ss;
    cbar();
    float tot = 0.f;
#pragma unroll
    for (int w = 0; w < kNW; ++w) tot += sm.red[w];
    const float inv = rsqrtf(tot / (float)k + opp->eps);
    const uint4* wg = reinterpret_cast<const uint4*>(opp->norm_w);
    for (int i = ctid; i < nvec; i += kConsumers) {   // each thread rescales the vectors it staged itself
      float xf[8], wf[8];
      unpack8(xv[i], xf);
      unpack8(wg[i], wf);
      uint4 o;  // reference order: normed = x * inv; normed *= w; one rounding
      o.x = pack2(xf[0] * inv * wf[0], xf[1] * inv * wf[1]);
      o.y = pack2(xf[2] * inv * wf[2], xf[3] * inv * wf[3]);
      o.z = pack2(xf[4] * inv * wf[4], xf[5] * inv * wf[5]);
      o.w = pack2(xf[6] * inv * wf[6], xf[7] * inv * wf[7]);
      xv[i] = o;
    }
  }
  cbar();
  if (trace && ctid == 0) *trace_slot(P, op_index, 1) = global_timer_ns();
  // ---- rows
  const int key = opp->rpw * 4 + opp->nseg;
  switch (key) {
    case 1 * 4 + 1: gemv_rows<1, 1>(ra, sm, ring, cprime, grid, tag_y, tag_res, best, warp, lane); break;
    case 2 * 4 + 1: gemv_rows<2, 1>(ra, sm, ring, cprime, grid, tag_y, tag_res, best, warp, lane); break;
    case 4 * 4 + 1: gemv_rows<4, 1>(ra, sm, ring, cprime, grid, tag_y, tag_res, best, warp, lane); break;
    case 1 * 4 + 2: gemv_rows<1, 2>(ra, sm, ring, cprime, grid, tag_y, tag_res, best, warp, lane); break;
    case 2 * 4 + 2: gemv_rows<2, 2>(ra, sm, ring, cprime, grid, tag_y, tag_res, best, warp, lane); break;
    default: __trap();
  }
  // the next op overwrites xs / red: every warp must be done reading them (and every row of this CTA is stored)
  cbar();
  if (opp->y_ctr && ctid == 0) ctr_arrive(P, op_index);
}

// ------------------------------------------------------------------------------------------------ attention
// One work item = (group of `heads_per_item` query heads of one KV head, KV split).  256 consumer threads =
// NG key groups × LPK lanes per key row; a thread keeps up to 8 K and 8 V pieces (16 bytes each) in registers.
template <int HD>
__device__ __forceinline__ void attn_op(const MegaOp* __restrict__ opp, const MegaParams& P, const Smem& sm,
                                        unsigned long long tok_seq, int op_index, int pos, int cta, int grid, int ctid,
                                        int warp, int lane, bool trace) {
  struct {
    int heads_per_item, cta_off, x_ev, layer;
    const __nv_bfloat16 *q_norm, *k_norm;
  } op;
  op.heads_per_item = opp->heads_per_item;
  op.cta_off = opp->cta_off;
  op.x_ev = opp->x_ev;
  op.layer = opp->layer;
  op.q_norm = opp->q_norm;
  op.k_norm = opp->k_norm;
  const uint32_t* const qkv_ll = P.qkv_ll + (size_t)(cta % P.qkv_rep) * P.qkv_rstride;
  constexpr int LPK = HD / 8;            // lanes per key row
  constexpr int NG = kConsumers / LPK;   // key rows per pass
  constexpr int KPT = 8;                 // key rows per thread
  constexpr int CHUNK = NG * KPT;        // keys per split: 256 (hd 64) / 128 (hd 128)
  constexpr int EPL = HD / 32;
  const float kScale = (HD == 64 ? 0.125f : 0.08838834764831845f) * kLog2e;

  const int Gh = op.heads_per_item;
  const int groups = P.Hq / Gh;
  const int L = pos + 1;
  const int nact = (L + CHUNK - 1) / CHUNK;
  const int items = groups * nact;
  int cprime = cta - op.cta_off;
  if (cprime < 0) cprime += grid;
  const uint32_t tag_q = ll_tag(tok_seq, P.events_per_token, op.x_ev);
  const uint32_t tag_y = ll_tag(tok_seq, P.events_per_token, op_index);
  const int qdim = P.Hq * HD, kvdim = P.Hkv * HD;
  const int gqa = P.Hq / P.Hkv;
  __nv_bfloat16* const kc = P.kcache + (size_t)op.layer * P.kv_layer_stride;
  __nv_bfloat16* const vc = P.vcache + (size_t)op.layer * P.kv_layer_stride;
  const int grp = ctid / LPK, part = ctid % LPK;

  for (int item = cprime; item < items; item += grid) {
    const int hg = item / nact, split = item % nact;
    const int h0 = hg * Gh;
    const int kvh = h0 / gqa;
    const bool kv_leader = (h0 % gqa) == 0 ;
    const int start = split * CHUNK;
    const int end = min(L, start + CHUNK);
    const bool owns_new = pos >= start && pos < end;

    // ---- 1. K/V rows of earlier tokens → registers (issued before this token's q/k/v is polled; L2 / HBM latency
    //         overlaps the wait).  Rows were written by other CTAs, possibly earlier in this launch: read through L2.
    uint4 kreg[KPT], vreg[KPT];
#pragma unroll
    for (int i = 0; i < KPT; ++i) {
      const int row = start + grp + NG * i;
      kreg[i] = make_uint4(0, 0, 0, 0);
      vreg[i] = make_uint4(0, 0, 0, 0);
      if (row < end && row < pos) {
        const size_t g = ((size_t)row * P.Hkv + kvh) * HD + part * 8;
        kreg[i] = __ldcg(reinterpret_cast<const uint4*>(kc + g));
        vreg[i] = __ldcg(reinterpret_cast<const uint4*>(vc + g));
      }
    }
    float rc[EPL / 2], rs[EPL / 2];
    {
      const float* rrow = P.rope + (size_t)pos * HD * 2;
#pragma unroll
      for (int j = 0; j < EPL / 2; ++j) {
        rc[j] = rrow[(lane + 32 * j) * 2];
        rs[j] = rrow[(lane + 32 * j) * 2 + 1];
      }
    }
    if (opp->x_ctr >= 0) {
      if (ctid == 0) ctr_wait(P, opp->x_ctr, (tok_seq + 1ull) * (unsigned long long)opp->x_ctr_count);
      cbar();
    }
    // ---- 2. this token's q heads (and k, v when the split holds the new row): optional per-head RMSNorm, RoPE, with
    //         the reference's roundings [ref: src/layer/Attention.h:156-163; TT/Operation/OpNNLayerCuda.cuh:412-440]
    for (int h = warp; h < Gh + 1; h += kNW) {
      const bool is_k = (h == Gh);
      if (is_k && !owns_new) break;
      const uint32_t* src = qkv_ll + (is_k ? qdim + kvh * HD : (h0 + h) * HD);
      const __nv_bfloat16* nw = is_k ? op.k_norm : op.q_norm;
      float x[EPL];
#pragma unroll
      for (int j = 0; j < EPL; ++j) x[j] = bf16_to_f(ll_poll1(src + lane + 32 * j, tag_q));
      if (nw != nullptr) {
        float ss = 0.f;
#pragma unroll
        for (int j = 0; j < EPL; ++j) ss += x[j] * x[j];
        ss = warp_sum(ss);
        const float inv = rsqrtf(ss / (float)HD + P.eps);
#pragma unroll
        for (int j = 0; j < EPL; ++j) x[j] = round_bf16(x[j] * inv * bf16_to_f(nw[lane + 32 * j]));
      }
#pragma unroll
      for (int j = 0; j < EPL / 2; ++j) {
        const float x1 = x[j], x2 = x[j + EPL / 2];
        x[j] = round_bf16(x1 * rc[j] - x2 * rs[j]);
        x[j + EPL / 2] = round_bf16(x2 * rc[j] + x1 * rs[j]);
      }
      if (is_k) {
        __nv_bfloat16* kg = kc + ((size_t)pos * P.Hkv + kvh) * HD;
        __nv_bfloat16* vg = vc + ((size_t)pos * P.Hkv + kvh) * HD;
#pragma unroll
        for (int j = 0; j < EPL; ++j) {
          const __nv_bfloat16 kk = f_to_bf16(x[j]);
          const __nv_bfloat16 vv = ll_poll1(qkv_ll + qdim + kvdim + kvh * HD + lane + 32 * j, tag_q);
          sm.knew[lane + 32 * j] = kk;
          sm.vnew[lane + 32 * j] = vv;
          if (kv_leader) {   // in place into the cache [ref: src/engine/CacheManager.h:24-42 appends by re-copying]
            kg[lane + 32 * j] = kk;
            vg[lane + 32 * j] = vv;
          }
        }
      } else {
#pragma unroll
        for (int j = 0; j < EPL; ++j) sm.q_s[h * HD + lane + 32 * j] = x[j];
      }
    }
    cbar();
    if (owns_new) {
#pragma unroll
      for (int i = 0; i < KPT; ++i) {
        if (start + grp + NG * i == pos) {
          kreg[i] = reinterpret_cast<const uint4*>(sm.knew)[part];
          vreg[i] = reinterpret_cast<const uint4*>(sm.vnew)[part];
        }
      }
    }
    // ---- 3. per query head: scores, online softmax over this thread's keys, P·V, merge across the CTA
    for (int g = 0; g < Gh; ++g) {
      float qf[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) qf[e] = sm.q_s[g * HD + part * 8 + e];
      float sc[KPT];
      float m = -INFINITY;
#pragma unroll
      for (int i = 0; i < KPT; ++i) {
        float d = dot8(kreg[i], qf, 0.f);
#pragma unroll
        for (int o = LPK / 2; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
        const bool valid = (start + grp + NG * i) < end;
        sc[i] = valid ? d : -INFINITY;
        m = fmaxf(m, sc[i]);
      }
      float l = 0.f, acc[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] = 0.f;
      const float m_scaled = m * kScale;
#pragma unroll
      for (int i = 0; i < KPT; ++i) {
        const float p = (sc[i] == -INFINITY) ? 0.f : exp2f(sc[i] * kScale - m_scaled);
        l += p;
        float vf[8];
        unpack8(vreg[i], vf);
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] = fmaf(p, vf[e], acc[e]);
      }
      // merge the key groups that live in the same warp (lanes with the same `part`)
#pragma unroll
      for (int o = LPK; o < 32; o <<= 1) {
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
        float* r = sm.ared + (warp * 16 + lane) * 10;
        r[0] = m;
        r[1] = l;
#pragma unroll
        for (int e = 0; e < 8; ++e) r[2 + e] = acc[e];
      }
      cbar();
      if (ctid < HD) {
        const int d = ctid, pt = d >> 3, e = d & 7;
        float M = -INFINITY;
#pragma unroll
        for (int w = 0; w < kNW; ++w) M = fmaxf(M, sm.ared[(w * 16 + pt) * 10]);
        float num = 0.f, den = 0.f;
#pragma unroll
        for (int w = 0; w < kNW; ++w) {
          const float* r = sm.ared + (w * 16 + pt) * 10;
          const float wgt = (r[0] == -INFINITY) ? 0.f : exp2f((r[0] - M) * kScale);
          num = fmaf(wgt, r[2 + e], num);
          den = fmaf(wgt, r[1], den);
        }
        if (nact == 1) {
          const float o = num * (den > 0.f ? 1.f / den : 0.f);
          const uint32_t word = ll_word(tag_y, f_to_bf16(o));
          for (int c = 0; c < P.attn_rep; ++c) st_volatile_u32(P.attn_ll + (size_t)c * P.attn_rstride + (h0 + g) * HD + d, word);
        } else {
          float* wsr = P.attn_ws + ((size_t)(h0 + g) * P.nsplit + split) * (HD + 2);
          wsr[d] = num;
          if (d == 0) {
            wsr[HD] = M;
            wsr[HD + 1] = den;
          }
        }
      }
      cbar();
    }
    // ---- 4. several splits: the last CTA of the head group (atomic ticket) merges the partials and publishes
    if (nact > 1) {
      __threadfence();
      cbar();
      if (ctid == 0) {
        const unsigned int t = atomicAdd(&P.attn_tickets[hg], 1u);
        sm.flag[0] = (t == (unsigned int)nact - 1) ? 1 : 0;
      }
      cbar();
      if (sm.flag[0]) {
        __threadfence();
        for (int idx = ctid; idx < Gh * HD; idx += kConsumers) {
          const int g = idx / HD, d = idx % HD;
          const float* base = P.attn_ws + (size_t)(h0 + g) * P.nsplit * (HD + 2);
          float M = -INFINITY;
          for (int s = 0; s < nact; ++s) M = fmaxf(M, __ldcg(base + (size_t)s * (HD + 2) + HD));
          float num = 0.f, den = 0.f;
          for (int s = 0; s < nact; ++s) {
            const float* r = base + (size_t)s * (HD + 2);
            const float w = exp2f((__ldcg(r + HD) - M) * kScale);
            num = fmaf(w, __ldcg(r + d), num);
            den = fmaf(w, __ldcg(r + HD + 1), den);
          }
          const uint32_t word = ll_word(tag_y, f_to_bf16(num * (den > 0.f ? 1.f / den : 0.f)));
          for (int c = 0; c < P.attn_rep; ++c) st_volatile_u32(P.attn_ll + (size_t)c * P.attn_rstride + (h0 + g) * HD + d, word);
        }
        if (ctid == 0) P.attn_tickets[hg] = 0;   // self-reset
      }
      cbar();
      if (opp->y_ctr && sm.flag[0] && ctid == 0) ctr_arrive(P, op_index);
    } else if (opp->y_ctr && ctid == 0) {
      ctr_arrive(P, op_index);   // the head loop ended with a CTA barrier after the last store
    }
    if (trace && ctid == 0) *trace_slot(P, op_index, 1) = global_timer_ns();
  }
}

// ------------------------------------------------------------------------------------------------ the kernel
__global__ void __launch_bounds__(kThreads, 1) token_kernel(const __grid_constant__ MegaParams P) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const Smem sm = carve(smem_raw, P);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int grid = (int)gridDim.x, cta = (int)blockIdx.x;

  if (threadIdx.x == 0) {
    for (int s = 0; s < P.stages; ++s) {
      mbar_init(&sm.full[s], 1);
      mbar_init(&sm.empty[s], kNW);
    }
    fence_mbar_init();
  }
  __syncthreads();

  if (warp == kNW) {
    // ---------------------------------------------------------------- producer: the token's weights, in op order
    if (lane != 0) return;
    Ring ring{0, 1u};   // first pass over the ring: slots are free
    for (int t = 0; t < P.n_tokens; ++t) {
      for (int j = 0; j < P.n_ops; ++j) {
        const MegaOp* __restrict__ opp = P.ops + j;
        if (opp->kind != MK_GEMV) continue;
        int cprime = cta - opp->cta_off;
        if (cprime < 0) cprime += grid;
        const CUtensorMap* tm = P.tmaps + opp->tmap;
        const int nseg = opp->nseg, rowblocks = opp->rowblocks, ksteps = opp->ksteps, seg_rows = opp->seg_rows;
        const int kb_n = max(1, 4 / (opp->rpw * nseg));
        const int box_r = kNW * opp->rpw;
        const int box_bytes = box_r * kRowBytes;
        const int stages = P.stages;
        for (int rb = cprime; rb < rowblocks; rb += grid) {
          const int row0 = rb * box_r;
          for (int ks = 0; ks < ksteps; ++ks) {
            mbar_wait_guarded(&sm.empty[ring.s], ring.ph);
            mbar_arrive_expect_tx(&sm.full[ring.s], kStageBytes);
            uint8_t* dst = sm.ring + (size_t)ring.s * kStageBytes;
            for (int kb = 0; kb < kb_n; ++kb)
              for (int seg = 0; seg < nseg; ++seg)
                tma_load_2d(dst + (kb * nseg + seg) * box_bytes, tm, (ks * kb_n + kb) * kBoxK, seg * seg_rows + row0,
                            &sm.full[ring.s]);
            ring_next(ring, stages);
          }
        }
      }
    }
    return;
  }

  // -------------------------------------------------------------------- consumers
  const int ctid = threadIdx.x;
  Ring ring{0, 0u};
  int pos = *P.pos;
  long long tok = *P.cur_tok;
  for (int t = 0; t < P.n_tokens; ++t) {
    const unsigned long long tok_seq = P.tok_seq0 + (unsigned long long)t;
    const bool trace = P.trace != nullptr && t == P.n_tokens - 1;
    if (trace && ctid == 0) *trace_slot(P, P.trace_ops - 1, 0) = global_timer_ns();
    // embedding row of this token → shared memory (x and residual of layer 0) [ref: OpTransformCuda.cuh:108-120]
    {
      long long id = tok;
      if (id < 0 || id >= P.V) id = 0;
      const uint4* src = reinterpret_cast<const uint4*>(P.embed + (size_t)id * P.H);
      uint4* dst = reinterpret_cast<uint4*>(sm.emb);
      for (int i = ctid; i < (P.H >> 3); i += kConsumers) dst[i] = src[i];
    }
    cbar();
    Best best{-INFINITY, -1};
    for (int j = 0; j < P.n_ops; ++j) {
      const MegaOp* __restrict__ opp = P.ops + j;
      if (trace && ctid == 0) *trace_slot(P, j, 0) = global_timer_ns();
      if (opp->kind == MK_GEMV) {
        gemv_op(opp, P, sm, ring, tok_seq, j, best, cta, grid, ctid, warp, lane, trace);
      } else if (P.hd == 64) {
        attn_op<64>(opp, P, sm, tok_seq, j, pos, cta, grid, ctid, warp, lane, trace);
      } else {
        attn_op<128>(opp, P, sm, tok_seq, j, pos, cta, grid, ctid, warp, lane, trace);
      }
      if (trace && ctid == 0) *trace_slot(P, j, 2) = global_timer_ns();
    }
    if (P.with_head) {
      // ---- greedy token: CTA candidate → LL exchange → every CTA picks the winner itself
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, best.v, o);
        const int oi = __shfl_xor_sync(0xffffffffu, best.i, o);
        best_merge(best, ov, oi);
      }
      if (lane == 0) {
        sm.cbv[warp] = best.v;
        sm.cbi[warp] = best.i;
      }
      cbar();
      const uint32_t tag_c = ll_tag(tok_seq, P.events_per_token, P.events_per_token - 1);
      if (ctid == 0) {
        Best b{-INFINITY, -1};
        for (int w = 0; w < kNW; ++w) best_merge(b, sm.cbv[w], sm.cbi[w]);
        st_volatile_u2(P.cand + 2 * cta, make_uint2(__float_as_uint(b.v), tag_c));
        st_volatile_u2(P.cand + 2 * cta + 1, make_uint2((uint32_t)b.i, tag_c));
      }
      Best w{-INFINITY, -1};
      for (int c = ctid; c < grid; c += kConsumers) {
        SpinGuard g;
        for (;;) {
          const uint2 a = ld_volatile_u2(P.cand + 2 * c);
          const uint2 b = ld_volatile_u2(P.cand + 2 * c + 1);
          if (a.y == tag_c && b.y == tag_c) {
            best_merge(w, __uint_as_float(a.x), (int)b.x);
            break;
          }
          g.tick();
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, w.v, o);
        const int oi = __shfl_xor_sync(0xffffffffu, w.i, o);
        best_merge(w, ov, oi);
      }
      cbar();   // every thread has read cbv/cbi of the CTA candidate
      if (lane == 0) {
        sm.cbv[warp] = w.v;
        sm.cbi[warp] = w.i;
      }
      cbar();
      Best win{-INFINITY, -1};
#pragma unroll
      for (int q = 0; q < kNW; ++q) best_merge(win, sm.cbv[q], sm.cbi[q]);
      cbar();   // cbv/cbi are rewritten by the next token
      tok = win.i;
      if (cta == 0 && ctid == 0) {
        *P.cur_tok = (int64_t)win.i;
        const unsigned long long c = *P.gen_count;
        P.gen_log[c % (unsigned long long)P.gen_cap] = (int64_t)win.i;
        *P.gen_count = c + 1;
        if (P.mailbox != nullptr) {   // {sequence tag, token} in one posted 8-byte store to pinned host memory
          const unsigned long long word = (((c + 1ull) & 0xffffffffull) << 32) | (unsigned long long)(unsigned int)win.i;
          asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(P.mailbox + (c % P.mailbox_cap)), "l"(word) : "memory");
        }
      }
    }
    pos += 1;
    if (cta == 0 && ctid == 0) *P.pos = pos;
    if (trace && ctid == 0) *trace_slot(P, P.trace_ops - 1, 1) = global_timer_ns();
  }
}

}  // namespace

int mega_setup_attributes() {
  static std::once_flag once;
  static int rc = B200_OK;
  std::call_once(once, [] {
    cudaError_t e = cudaFuncSetAttribute(token_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(token_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) {
      set_error("cudaFuncSetAttribute(token_kernel) failed: %s", cudaGetErrorString(e));
      rc = B200_ERR_CUDA;
      (void)cudaGetLastError();
    }
  });
  return rc;
}

size_t mega_smem_fixed(int xs_elems, int H) { return smem_fixed_bytes(xs_elems, H); }

int mega_launch(const MegaPlan& plan, int n_tokens, bool with_head, unsigned long long tok_seq0, cudaStream_t st) {
  MegaParams p = plan.p;
  p.n_tokens = n_tokens;
  p.with_head = with_head ? 1 : 0;
  p.n_ops = with_head ? plan.n_ops_body + 1 : plan.n_ops_body;
  p.tok_seq0 = tok_seq0;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(plan.grid);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = (size_t)plan.smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;   // all CTAs co-resident or the launch fails: they spin on each other
  attr[0].val.cooperative = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  g_launches.fetch_add(1, std::memory_order_relaxed);
  B200_CUDA(cudaLaunchKernelEx(&cfg, token_kernel, p));
  return B200_OK;
}

}  // namespace b200
