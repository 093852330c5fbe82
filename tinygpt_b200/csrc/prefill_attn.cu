// prefill_attn.cu — causal flash attention of a PREFILL chunk on the tensor cores (mma.sync m16n8k16 bf16, fp32
// accumulate), sm_100a.  The default prefill attention (common.cuh Defaults::kPrefillAttnMma; parity against the oracle
// and the reference's CUDA path: tests/test_prefill_gpu.py); B200_PREFILL_ATTN=cuda selects the CUDA-core kernel in
// prefill.cu, which also serves the shapes this one does not tile.
//
// Replaces tfa::flashAttn for Sq > 1 [ref: TFA/mma/kernel.cuh:18-203 (tile loop, last → first), TFA/mma/softmax.cuh:67-131
// (online softmax, base-2 exponent, scale = log2e/√hd), TFA/mma/layout.cuh:88-97 (P rounded to bf16 before P·V),
// TFA/mma/memory.cuh:84-97 (O·(1/rowsum) rounded once)] with the same arithmetic: bf16 Q·Kᵀ products accumulated in fp32,
// P → bf16, bf16 P·V products accumulated in fp32.  Decode (Sq = 1) never comes here (attn.cu).
//
// Shape of the work: CTA = 64 query rows of one query head, 4 warps × 16 rows; keys are visited in tiles of 64 from the
// diagonal tile down to tile 0 (fully masked tiles are never touched), K and V tiles are double-buffered in shared
// memory with cp.async, 16-byte chunks XOR-swizzled by (row & 7) so that every ldmatrix phase is conflict-free.
// Query row r of the chunk sits at position p0 + r and sees keys 0 … p0 + r (the cached prefix plus the causal part
// of the chunk); K/V come straight from the in-place cache [max_ctx, Hkv, hd], q from the merged qkv rows.
// Late query blocks (most keys) are scheduled first.
#include "common.cuh"
#include "ops.cuh"

#include <cstdlib>
#include <mutex>

namespace b200 {

namespace {

constexpr int kBM = 64, kBN = 64, kPfWarps = 4, kPfThreads = kPfWarps * 32;
constexpr float kLog2eP = 1.4426950408889634f;

template <int HD>
struct PfScale;
template <>
struct PfScale<64> {
  static constexpr float value = 0.125f * kLog2eP;
};
template <>
struct PfScale<128> {
  static constexpr float value = 0.08838834764831845f * kLog2eP;
};

__device__ __forceinline__ void pf_cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void pf_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void pf_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr)
               : "memory");
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr)
               : "memory");
}
// D(16×8, fp32) += A(16×16, bf16, row) · B(16×8, bf16, col)
__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

struct PrefillAttnParams {
  __nv_bfloat16* o;             // [B][S][Hq][HD]          (row stride o_rs, batch stride o_bs, elements)
  const __nv_bfloat16* q;       // head h of row r at q + b*q_bs + r*q_rs + h*HD
  const __nv_bfloat16* k;       // key j of KV head g at k + b*kv_bs + j*kv_rs + g*HD
  const __nv_bfloat16* v;
  int64_t o_rs, o_bs, q_rs, q_bs, kv_rs, kv_bs;
  int S, Hq, Hkv, p0;
};

// byte offset of 16-byte chunk c of row r inside a [rows][HD] bf16 tile (chunks XOR-swizzled within groups of 8)
template <int HD>
__device__ __forceinline__ uint32_t sw_off(int r, int c) {
  return (uint32_t)(r * (HD * 2) + ((c ^ (r & 7)) << 4));
}

template <int HD>
__global__ void __launch_bounds__(kPfThreads) attn_prefill_mma_kernel(const PrefillAttnParams p) {
  constexpr int CH = HD / 8;    // 16-byte chunks per row
  constexpr int KS = HD / 16;   // k-steps of Q·Kᵀ
  constexpr int ND = HD / 8;    // 8-wide output column blocks of O
  constexpr float kScale = PfScale<HD>::value;

  extern __shared__ __align__(128) uint8_t pf_smem[];
  uint8_t* const Qs = pf_smem;                       // [kBM][HD]
  uint8_t* const Ks = Qs + kBM * HD * 2;             // [2][kBN][HD]
  uint8_t* const Vs = Ks + 2 * kBN * HD * 2;         // [2][kBN][HD]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, tq = lane & 3;
  const int qb = (int)gridDim.x - 1 - (int)blockIdx.x;   // heavy blocks first
  const int h = blockIdx.y, b = blockIdx.z;
  const int kvh = h / (p.Hq / p.Hkv);
  const int q0 = qb * kBM;
  const int rows_valid = min(kBM, p.S - q0);
  const int nkeys = p.p0 + q0 + rows_valid;              // keys the last valid row of this block sees
  const int T = (nkeys + kBN - 1) / kBN;

  const __nv_bfloat16* qg = p.q + (size_t)b * p.q_bs + (size_t)h * HD;
  const __nv_bfloat16* kg = p.k + (size_t)b * p.kv_bs + (size_t)kvh * HD;
  const __nv_bfloat16* vg = p.v + (size_t)b * p.kv_bs + (size_t)kvh * HD;

  // ---- Q tile (group 0)
  for (int i = tid; i < kBM * CH; i += kPfThreads) {
    const int r = i / CH, c = i % CH;
    uint8_t* dst = Qs + sw_off<HD>(r, c);
    if (r < rows_valid) pf_cp_async16(dst, qg + (size_t)(q0 + r) * p.q_rs + c * 8);
    else *reinterpret_cast<uint4*>(dst) = make_uint4(0, 0, 0, 0);
  }
  pf_commit();
  auto load_kv = [&](int t, int buf) {
    const int k0 = t * kBN;
    uint8_t* kd = Ks + buf * (kBN * HD * 2);
    uint8_t* vd = Vs + buf * (kBN * HD * 2);
    for (int i = tid; i < kBN * CH; i += kPfThreads) {
      const int r = i / CH, c = i % CH;
      const uint32_t off = sw_off<HD>(r, c);
      const int key = k0 + r;
      if (key < nkeys) {
        pf_cp_async16(kd + off, kg + (size_t)key * p.kv_rs + c * 8);
        pf_cp_async16(vd + off, vg + (size_t)key * p.kv_rs + c * 8);
      } else {  // beyond the last visible key: zeros (masked to -inf in S; P = 0 must not meet a NaN in V)
        *reinterpret_cast<uint4*>(kd + off) = make_uint4(0, 0, 0, 0);
        *reinterpret_cast<uint4*>(vd + off) = make_uint4(0, 0, 0, 0);
      }
    }
    pf_commit();
  };
  load_kv(T - 1, 0);

  uint32_t qf[KS][4];
  float o_acc[ND][4];
#pragma unroll
  for (int n = 0; n < ND; ++n)
#pragma unroll
    for (int e = 0; e < 4; ++e) o_acc[n][e] = 0.f;
  float m_run[2] = {-INFINITY, -INFINITY};
  float l_run[2] = {0.f, 0.f};            // per-thread partial row sums (quad-reduced once at the end)
  // positions of this thread's two rows; padding rows of a ragged last block behave like the last valid row
  int rpos[2];
  rpos[0] = p.p0 + q0 + min(warp * 16 + g, rows_valid - 1);
  rpos[1] = p.p0 + q0 + min(warp * 16 + g + 8, rows_valid - 1);

  const uint32_t qs_u = smem_u32(Qs), ks_u = smem_u32(Ks), vs_u = smem_u32(Vs);

  for (int it = 0; it < T; ++it) {
    const int t = T - 1 - it, buf = it & 1;
    if (it + 1 < T) {
      load_kv(t - 1, buf ^ 1);
      pf_wait<1>();
    } else {
      pf_wait<0>();
    }
    __syncthreads();
    if (it == 0) {
#pragma unroll
      for (int kk = 0; kk < KS; ++kk)
        ldsm_x4(qs_u + sw_off<HD>(warp * 16 + (lane & 15), 2 * kk + (lane >> 4)), qf[kk]);
    }

    // ---- S = Q·Kᵀ for this warp's 16 rows × 64 keys
    float s[8][4];
#pragma unroll
    for (int n = 0; n < 8; ++n)
#pragma unroll
      for (int e = 0; e < 4; ++e) s[n][e] = 0.f;
    const uint32_t kt_u = ks_u + buf * (kBN * HD * 2);
#pragma unroll
    for (int kk = 0; kk < KS; ++kk) {
#pragma unroll
      for (int n2 = 0; n2 < 4; ++n2) {
        uint32_t kb[4];
        ldsm_x4(kt_u + sw_off<HD>(n2 * 16 + (lane & 7) + ((lane >> 4) << 3), 2 * kk + ((lane >> 3) & 1)), kb);
        mma_bf16_16816(s[2 * n2], qf[kk], kb[0], kb[1]);
        mma_bf16_16816(s[2 * n2 + 1], qf[kk], kb[2], kb[3]);
      }
    }

    // ---- causal mask (only tiles that reach past the block's first row can contain masked keys)
    const int kbase = t * kBN;
    if (kbase + kBN - 1 > p.p0 + q0) {
#pragma unroll
      for (int n = 0; n < 8; ++n) {
        const int j = kbase + n * 8 + 2 * tq;
        if (j > rpos[0]) s[n][0] = -INFINITY;
        if (j + 1 > rpos[0]) s[n][1] = -INFINITY;
        if (j > rpos[1]) s[n][2] = -INFINITY;
        if (j + 1 > rpos[1]) s[n][3] = -INFINITY;
      }
    }

    // ---- online softmax, base 2, for rows g (rr = 0) and g + 8 (rr = 1)
    uint32_t pa[4][4];  // P as A fragments: [16-key block][a0..a3]
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
      float mx = -INFINITY;
#pragma unroll
      for (int n = 0; n < 8; ++n) mx = fmaxf(mx, fmaxf(s[n][2 * rr], s[n][2 * rr + 1]));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
      const float m_new = fmaxf(m_run[rr], mx);
      const float alpha = (m_run[rr] == -INFINITY) ? 0.f : exp2f((m_run[rr] - m_new) * kScale);
      const float m_scaled = (m_new == -INFINITY) ? 0.f : m_new * kScale;
      float sum = 0.f;
#pragma unroll
      for (int n = 0; n < 8; ++n) {
        const float p0v = exp2f(s[n][2 * rr] * kScale - m_scaled);
        const float p1v = exp2f(s[n][2 * rr + 1] * kScale - m_scaled);
        sum += p0v + p1v;                      // row sum of the fp32 probabilities (like the reference)
        pa[n >> 1][(n & 1) * 2 + rr] = pack2(p0v, p1v);   // a0/a2 = row g, a1/a3 = row g+8
      }
      l_run[rr] = l_run[rr] * alpha + sum;
      m_run[rr] = m_new;
#pragma unroll
      for (int n = 0; n < ND; ++n) {
        o_acc[n][2 * rr] *= alpha;
        o_acc[n][2 * rr + 1] *= alpha;
      }
    }

    // ---- O += P·V
    const uint32_t vt_u = vs_u + buf * (kBN * HD * 2);
#pragma unroll
    for (int k2 = 0; k2 < 4; ++k2) {
#pragma unroll
      for (int d2 = 0; d2 < ND / 2; ++d2) {
        uint32_t vb[4];
        ldsm_x4_trans(vt_u + sw_off<HD>(k2 * 16 + (lane & 7) + (((lane >> 3) & 1) << 3), 2 * d2 + (lane >> 4)), vb);
        mma_bf16_16816(o_acc[2 * d2], pa[k2], vb[0], vb[1]);
        mma_bf16_16816(o_acc[2 * d2 + 1], pa[k2], vb[2], vb[3]);
      }
    }
    __syncthreads();  // every warp is done with `buf` before the next iteration's prefetch overwrites it
  }

  // ---- normalise and store
#pragma unroll
  for (int rr = 0; rr < 2; ++rr) {
    float l = l_run[rr];
    l += __shfl_xor_sync(0xffffffffu, l, 1);
    l += __shfl_xor_sync(0xffffffffu, l, 2);
    const float inv = l > 0.f ? 1.f / l : 0.f;
    const int r = warp * 16 + g + 8 * rr;
    if (r < rows_valid) {
      __nv_bfloat16* dst = p.o + (size_t)b * p.o_bs + (size_t)(q0 + r) * p.o_rs + (size_t)h * HD + 2 * tq;
#pragma unroll
      for (int n = 0; n < ND; ++n)
        *reinterpret_cast<uint32_t*>(dst + n * 8) = pack2(o_acc[n][2 * rr] * inv, o_acc[n][2 * rr + 1] * inv);
    }
  }
}

template <int HD>
int launch_pf(const PrefillAttnParams& p, int B, cudaStream_t st) {
  constexpr int smem = (kBM + 4 * kBN) * HD * 2;
  static std::once_flag once;   // one per instantiation; the reference's server has a worker thread next to main
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] {
    attr_err = cudaFuncSetAttribute(attn_prefill_mma_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  });
  B200_CUDA(attr_err);
  dim3 grid((unsigned)((p.S + kBM - 1) / kBM), (unsigned)p.Hq, (unsigned)B);
  g_launches.fetch_add(1);
  attn_prefill_mma_kernel<HD><<<grid, kPfThreads, smem, st>>>(p);
  B200_CUDA(cudaGetLastError());
  return B200_OK;
}

}  // namespace

bool prefill_attn_mma_enabled() {  // read per call (a getenv per prefill launch is noise; tests toggle it in-process)
  return env_choice("B200_PREFILL_ATTN", 'm', Defaults::kPrefillAttnMma);  // "mma" | "cuda"
}

// Engine prefill: q inside the merged qkv rows [S, qdim + 2 kvdim], K/V in the in-place cache [max_ctx, Hkv, hd].
int launch_attn_prefill_mma(void* o, const void* qkv, const void* kcache, const void* vcache, int S, int Hq, int Hkv, int hd,
                            int p0, cudaStream_t st) {
  B200_CHECK_ARG(hd == 64 || hd == 128, "prefill attention: head_dim %d not built", hd);
  B200_CHECK_ARG(S > 0 && Hkv > 0 && Hq % Hkv == 0 && Hq < 65536 && p0 >= 0, "prefill attention: bad shape");
  PrefillAttnParams p{};
  p.o = (__nv_bfloat16*)o;
  p.q = (const __nv_bfloat16*)qkv;
  p.k = (const __nv_bfloat16*)kcache;
  p.v = (const __nv_bfloat16*)vcache;
  p.o_rs = (int64_t)Hq * hd;
  p.q_rs = (int64_t)(Hq + 2 * Hkv) * hd;
  p.kv_rs = (int64_t)Hkv * hd;
  p.S = S;
  p.Hq = Hq;
  p.Hkv = Hkv;
  p.p0 = p0;
  return hd == 64 ? launch_pf<64>(p, 1, st) : launch_pf<128>(p, 1, st);
}

// Boundary B (b200_attn_bf16, causal, Sq == Skv): separate BSHD tensors, the reference's top-left-aligned mask.
int launch_attn_causal_mma(void* o, const void* q, const void* k, const void* v, int64_t B, int64_t S, int64_t Hq,
                           int64_t Hkv, int64_t hd, cudaStream_t st) {
  B200_CHECK_ARG(hd == 64 || hd == 128, "attention: head_dim %lld not built (64 and 128 are)", (long long)hd);
  B200_CHECK_ARG(B > 0 && B < 65536 && S > 0 && S < (1ll << 30) && Hkv > 0 && Hq % Hkv == 0 && Hq < 65536,
                 "attention: bad shape");
  PrefillAttnParams p{};
  p.o = (__nv_bfloat16*)o;
  p.q = (const __nv_bfloat16*)q;
  p.k = (const __nv_bfloat16*)k;
  p.v = (const __nv_bfloat16*)v;
  p.o_rs = p.q_rs = Hq * hd;
  p.kv_rs = Hkv * hd;
  p.o_bs = p.q_bs = S * Hq * hd;
  p.kv_bs = S * Hkv * hd;
  p.S = (int)S;
  p.Hq = (int)Hq;
  p.Hkv = (int)Hkv;
  p.p0 = 0;
  return hd == 64 ? launch_pf<64>(p, (int)B, st) : launch_pf<128>(p, (int)B, st);
}

}  // namespace b200
