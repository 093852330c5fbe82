// gemv.cu — the weight-bound bf16 GEMV that dominates batch-1 decode, written for sm_100a.
//
// Replaces, for m = 1:  Linear::forward → op::matmul(x, W, false, true, bias) → cublasGemmStridedBatchedEx
//   [ref: third_party/TinyTorch/src/Operation/OpLinalg.cpp:244-277, OpLinalgCuda.cuh:276-293]
// and fuses what the reference runs as separate launches around it:
//   RMSNorm prologue      [ref: TT/Operation/OpNNLayerCuda.cuh:252-357]
//   bias add epilogue     [ref: TT/Operation/OpLinalg.cpp:273-275 → OpElemWiseCuda.cuh:395-405]   (second rounding kept)
//   residual add epilogue [ref: src/layer/DecoderLayer.h:38-43 → TT/Operation/OpElemWiseCuda.cuh:133-144]
//   SiLU·mul epilogue     [ref: TT/Operation/OpFusedCuda.cuh:15-29, OpElemWiseCuda.cuh:124-131]   (both roundings kept)
//
// Design (HBM-bound: 2 bytes of W per FMA, so the only goal is to keep HBM saturated):
//   * W[n,k] row-major is cut into boxes of (8·RPW rows) × (256 columns) = RPW·4 KB.  One elected producer thread
//     streams the boxes of this CTA's row blocks with TMA (cp.async.bulk.tensor.2d → SASS UTMALDG) into a ring of
//     `stages` shared-memory slots guarded by full/empty mbarriers; no registers are spent on loads in flight and
//     ~80 KB per SM stay outstanding.
//   * 8 consumer warps: warp w owns rows w·RPW … w·RPW+RPW-1 of the box; a lane reads one 16-byte vector of x and
//     RPW 16-byte vectors of W per box (conflict-free: 32 lanes × 16 B = one 512-byte row), 8 FMAs each, fp32
//     accumulators carried across the k loop; one xor-shuffle tree per row at the end.
//   * Programmatic dependent launch: barrier init and the first `stages` TMA boxes are issued BEFORE
//     griddepcontrol.wait — weights do not depend on the previous kernel — so the HBM stream of kernel i+1 starts
//     while kernel i drains.  Only the activation vector is read after the wait.
//   * grid = min(row blocks, #SMs) persistent CTAs, row blocks dealt round-robin.
#include "gemv.cuh"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <mutex>

namespace b200 {

using namespace gemvk;

namespace {

__device__ __forceinline__ uint4 ld_volatile_u4(const uint4* p) {
  uint4 v;
  asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_volatile_u2(uint2* p, const uint2 v) {
  asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(v.x), "r"(v.y) : "memory");
}

template <int RPW, int NSEG, int PRO, int EPI>
__global__ void __launch_bounds__(kThreads, 1)
gemv_stream_kernel(const GemvParams p, const __grid_constant__ CUtensorMap tmap) {
  constexpr int kBoxR = kNW * RPW;
  constexpr int KB = kboxes(RPW, NSEG);                 // 256-column boxes per stage (stage = 16 KB, 32 KB for 4×2)
  constexpr int kBoxBytes = kBoxR * kRowBytes;
  constexpr int kStageBytes = KB * NSEG * kBoxBytes;

  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* const stage_base = smem;
  __nv_bfloat16* const xs = reinterpret_cast<__nv_bfloat16*>(smem + (size_t)p.stages * kStageBytes);
  uint64_t* const full = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(xs) + (size_t)p.k_pad * 2);
  uint64_t* const empty = full + p.stages;
  float* const red = reinterpret_cast<float*>(empty + p.stages);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int ksteps = p.k_pad / (kBoxK * KB);
  const int my_rbs = (p.rowblocks - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

  if (p.trace != nullptr && blockIdx.x == 0 && threadIdx.x == 0) p.trace[0] = global_timer_ns();
  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], kNW);
    }
    fence_mbar_init();
  }
  __syncthreads();
  pdl_trigger();  // dependents only prefetch weights before their own griddepcontrol.wait

  if (warp == kNW) {
    // ------------------------------------------------------------------ producer: one thread drives the TMA ring
    if (lane == 0) {
      tma_prefetch_desc(&tmap);
      const uint64_t pol = l2_policy_evict_first();   // every weight byte is read once per token
      int s = 0;
      uint32_t ph = 1;  // first pass over the ring: slots are free (wait on the "previous" phase returns at once)
      for (int i = 0; i < my_rbs; ++i) {
        const int row0 = ((int)blockIdx.x + i * (int)gridDim.x) * kBoxR;
        for (int ks = 0; ks < ksteps; ++ks) {
          mbar_wait(&empty[s], ph);
          mbar_arrive_expect_tx(&full[s], kStageBytes);
          uint8_t* dst = stage_base + (size_t)s * kStageBytes;
          // one box per segment: kBoxR rows × (KB · 256) columns = KB · 512 contiguous bytes per row (the tensor map
          // describes W in 2/4/8-byte elements so that the inner box dimension stays ≤ 256, gemv_make_plan)
#pragma unroll
          for (int seg = 0; seg < NSEG; ++seg)
            tma_load_2d_hint(dst + seg * (KB * kBoxBytes), &tmap, ks * kBoxK, seg * p.seg_rows + row0, &full[s], pol);
          if (++s == p.stages) {
            s = 0;
            ph ^= 1;
          }
        }
      }
    }
    return;
  }

  // -------------------------------------------------------------------- consumers
  const int ctid = threadIdx.x;  // 0 … 255
  const int nvec = p.k >> 3, nvec_pad = p.k_pad >> 3;
  constexpr int kMaxHoist = 2;  // 16-byte pieces of the norm weight a thread keeps in registers (k ≤ 4096)

  // Everything that does not depend on the producer kernel is fetched BEFORE griddepcontrol.wait: the RMSNorm weight,
  // the bias of this CTA's first row block.
  uint4 wn[kMaxHoist];
  if constexpr (PRO != PRO_PLAIN) {
    const uint4* wg = reinterpret_cast<const uint4*>(p.norm_w);
#pragma unroll
    for (int j = 0; j < kMaxHoist; ++j) {
      const int i = ctid + j * kConsumers;
      wn[j] = (i < nvec) ? wg[i] : make_uint4(0, 0, 0, 0);
    }
  }
  const int my_row0 = (int)blockIdx.x * kBoxR + warp * RPW + ((lane < RPW) ? lane : 0);
  __nv_bfloat16 bias_v = f_to_bf16(0.f);
  if constexpr (EPI == EPI_PLAIN) {
    if (p.bias != nullptr && lane < RPW && my_rbs > 0 && my_row0 < p.n) bias_v = p.bias[my_row0];
  }

  // the producer kernel's output (x, residual) is complete and visible from here on
  pdl_wait();
  auto ldx = [](const uint4* q) -> uint4 { return *q; };
  unsigned int tp_tag = 0;
  if constexpr (EPI == EPI_TP_PUSH) tp_tag = (unsigned int)(*p.tp_epoch + 1ull);
  const bool tracing = p.trace != nullptr && blockIdx.x == 0 && ctid == 0;
  long long c0 = 0;
  if (tracing) {
    p.trace[1] = global_timer_ns();
    c0 = clock64();   // SM cycles from here: slots 4-7 = first stage landed, first row block summed / stored, last stored
  }

  if constexpr (PRO == PRO_PLAIN) {
    const uint4* xg = reinterpret_cast<const uint4*>(p.x);
    uint4* xv = reinterpret_cast<uint4*>(xs);
    for (int i = ctid; i < nvec_pad; i += kConsumers) xv[i] = (i < nvec) ? ldx(xg + i) : make_uint4(0, 0, 0, 0);
  } else {
    // hidden vector h (k elements): either x itself, or residual + Σ_r partial_r (tensor parallel)
    float ss = 0.f;
    uint4 xr[kMaxHoist];
    if constexpr (PRO == PRO_TP_RMSNORM) {
      // h = bf16(residual + bf16(Σ_r partial_r)) in rank order (bitwise identical on every rank).  Each partial is an
      // 8-byte {value, tag} word written by the owning rank's GEMV epilogue; spin until the tag is this token's.
      const unsigned int want = (unsigned int)(*p.tp_epoch + 1ull);
      const int nq = p.k >> 2;
      for (int i = ctid; i < nq; i += kConsumers) {
        // all ranks' words are requested before any is inspected (one L2 round trip, not `world` of them)
        uint4 w0[kMaxTpWorld], w1[kMaxTpWorld];
        unsigned int spins = 0;
        unsigned long long t0 = 0;
        for (;;) {
#pragma unroll
          for (int r = 0; r < kMaxTpWorld; ++r)
            if (r < p.tp_world) {
              const uint4* src = reinterpret_cast<const uint4*>(p.tp_partials + (size_t)r * p.tp_stride + 4 * i);
              w0[r] = ld_volatile_u4(src);
              w1[r] = ld_volatile_u4(src + 1);
            }
          bool ok = true;
#pragma unroll
          for (int r = 0; r < kMaxTpWorld; ++r)
            if (r < p.tp_world) ok = ok && w0[r].y == want && w0[r].w == want && w1[r].y == want && w1[r].w == want;
          if (ok) break;
          if ((++spins & 0x3fffu) == 0) {  // a peer died: fail loudly after ~4 s, never hang the GPU
            const unsigned long long now = global_timer_ns();
            if (t0 == 0) t0 = now;
            if (now - t0 > 4000000000ull) __trap();
          }
        }
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int r = 0; r < kMaxTpWorld; ++r)
          if (r < p.tp_world) {  // rank order: bitwise identical on every rank
            a.x += __uint_as_float(w0[r].x);
            a.y += __uint_as_float(w0[r].z);
            a.z += __uint_as_float(w1[r].x);
            a.w += __uint_as_float(w1[r].z);
          }
        const uint2 rr = *reinterpret_cast<const uint2*>(p.tp_residual + 4 * i);
        const __nv_bfloat162 r01 = *reinterpret_cast<const __nv_bfloat162*>(&rr.x);
        const __nv_bfloat162 r23 = *reinterpret_cast<const __nv_bfloat162*>(&rr.y);
        const __nv_bfloat162 h01 = __hadd2(r01, __floats2bfloat162_rn(a.x, a.y));
        const __nv_bfloat162 h23 = __hadd2(r23, __floats2bfloat162_rn(a.z, a.w));
        uint2 hv;
        hv.x = *reinterpret_cast<const uint32_t*>(&h01);
        hv.y = *reinterpret_cast<const uint32_t*>(&h23);
        *reinterpret_cast<uint2*>(xs + 4 * i) = hv;
        if (blockIdx.x == 0) *reinterpret_cast<uint2*>(p.tp_h_out + 4 * i) = hv;
        const float2 f01 = __bfloat1622float2(h01), f23 = __bfloat1622float2(h23);
        ss += f01.x * f01.x + f01.y * f01.y + f23.x * f23.x + f23.y * f23.y;
      }
    } else {
      const uint4* xg = reinterpret_cast<const uint4*>(p.x);
#pragma unroll
      for (int j = 0; j < kMaxHoist; ++j) {
        const int i = ctid + j * kConsumers;
        xr[j] = (i < nvec) ? ldx(xg + i) : make_uint4(0, 0, 0, 0);
        float xf[8];
        unpack8(xr[j], xf);
#pragma unroll
        for (int e = 0; e < 8; ++e) ss += xf[e] * xf[e];
      }
      for (int i = ctid + kMaxHoist * kConsumers; i < nvec; i += kConsumers) {
        float xf[8];
        unpack8(ldx(xg + i), xf);
#pragma unroll
        for (int e = 0; e < 8; ++e) ss += xf[e] * xf[e];
      }
    }
    ss = warp_sum(ss);
    if (lane == 0) red[warp] = ss;
    named_bar_sync(1, kConsumers);
    float tot = 0.f;
#pragma unroll
    for (int w = 0; w < kNW; ++w) tot += red[w];
    const float inv = rsqrtf(tot / (float)p.k + p.eps);
    const uint4* wg = reinterpret_cast<const uint4*>(p.norm_w);
    uint4* xv = reinterpret_cast<uint4*>(xs);
    auto scale = [&](const uint4& xq, const uint4& wq) {
      float xf[8], wf[8];
      unpack8(xq, xf);
      unpack8(wq, wf);
      uint4 o;  // reference order: normed = x * inv; normed *= w; round once
      o.x = pack2(xf[0] * inv * wf[0], xf[1] * inv * wf[1]);
      o.y = pack2(xf[2] * inv * wf[2], xf[3] * inv * wf[3]);
      o.z = pack2(xf[4] * inv * wf[4], xf[5] * inv * wf[5]);
      o.w = pack2(xf[6] * inv * wf[6], xf[7] * inv * wf[7]);
      return o;
    };
#pragma unroll
    for (int j = 0; j < kMaxHoist; ++j) {
      const int i = ctid + j * kConsumers;
      if (i < nvec_pad) {
        if constexpr (PRO == PRO_TP_RMSNORM) {
          xv[i] = (i < nvec) ? scale(xv[i], wn[j]) : make_uint4(0, 0, 0, 0);
        } else {
          xv[i] = (i < nvec) ? scale(xr[j], wn[j]) : make_uint4(0, 0, 0, 0);
        }
      }
    }
    for (int i = ctid + kMaxHoist * kConsumers; i < nvec_pad; i += kConsumers) {
      uint4 o = make_uint4(0, 0, 0, 0);
      if (i < nvec) {
        if constexpr (PRO == PRO_TP_RMSNORM) {
          o = scale(xv[i], wg[i]);
        } else {
          o = scale(ldx(reinterpret_cast<const uint4*>(p.x) + i), wg[i]);
        }
      }
      xv[i] = o;
    }
  }
  named_bar_sync(1, kConsumers);
  if (p.trace != nullptr && blockIdx.x == 0 && ctid == 0) p.trace[3] = global_timer_ns();

  // ---------------------------------------------------------------------------------------------- main k loop
  int s = 0;
  uint32_t ph = 0;
  const uint8_t* const my_rows = stage_base + (size_t)(warp * RPW) * (KB * kRowBytes) + lane * 16;
  for (int i = 0; i < my_rbs; ++i) {
    const int row_base = ((int)blockIdx.x + i * (int)gridDim.x) * kBoxR + warp * RPW;
    // operands of the epilogue are requested now so that their latency hides behind the k loop
    __nv_bfloat16 res_v = f_to_bf16(0.f);
    if constexpr (EPI == EPI_RESIDUAL) {
      if (lane < RPW && row_base + lane < p.n) res_v = p.residual[row_base + lane];
    }
    if constexpr (EPI == EPI_PLAIN) {
      if (i > 0 && p.bias != nullptr && lane < RPW && row_base + lane < p.n) bias_v = p.bias[row_base + lane];
    }

    float acc[NSEG][RPW], acc_b[NSEG][RPW];
#pragma unroll
    for (int seg = 0; seg < NSEG; ++seg)
#pragma unroll
      for (int r = 0; r < RPW; ++r) acc[seg][r] = acc_b[seg][r] = 0.f;

    for (int ks = 0; ks < ksteps; ++ks) {
      uint4 xq[KB];
#pragma unroll
      for (int kb = 0; kb < KB; ++kb)
        xq[kb] = *reinterpret_cast<const uint4*>(xs + (ks * KB + kb) * kBoxK + lane * 8);
      mbar_wait(&full[s], ph);
      if (tracing && i == 0 && ks == 0) p.trace[4] = (unsigned long long)(clock64() - c0);
      const uint8_t* st = my_rows + (size_t)s * kStageBytes;
      uint4 wv[KB][NSEG][RPW];
#pragma unroll
      for (int kb = 0; kb < KB; ++kb)
#pragma unroll
        for (int seg = 0; seg < NSEG; ++seg)
#pragma unroll
          for (int r = 0; r < RPW; ++r)
            wv[kb][seg][r] = *reinterpret_cast<const uint4*>(st + seg * (KB * kBoxBytes) + r * (KB * kRowBytes) + kb * kRowBytes);
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
      if (lane == 0) mbar_arrive(&empty[s]);
      if (++s == p.stages) {
        s = 0;
        ph ^= 1;
      }
    }

    if (tracing && i == 0) p.trace[5] = (unsigned long long)(clock64() - c0);
#pragma unroll
    for (int seg = 0; seg < NSEG; ++seg)
#pragma unroll
      for (int r = 0; r < RPW; ++r) acc[seg][r] = warp_sum(acc[seg][r] + acc_b[seg][r]);

    // lane r finishes row r of this warp
    float a0 = acc[0][0], a1 = acc[NSEG - 1][0];
#pragma unroll
    for (int r = 1; r < RPW; ++r) {
      if (lane == r) {
        a0 = acc[0][r];
        a1 = acc[NSEG - 1][r];
      }
    }
    const int row = row_base + lane;
    if constexpr (EPI == EPI_TP_PUSH && RPW > 1) {
      // after the xor-shuffle trees every lane holds all RPW row sums: lane q serves peer q and stores the warp's RPW
      // adjacent {value, tag} words as 16-byte vectors (RPW/2 NVLink writes per peer and warp instead of RPW 8-byte
      // ones issued one peer after the other by RPW lanes)
      if (row_base + RPW <= p.n) {
        if (lane < p.tp_world) {
          uint2* dst = p.tp_push[lane] + row_base;
#pragma unroll
          for (int r = 0; r < RPW; r += 2)
            asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(dst + r), "r"(__float_as_uint(acc[0][r])),
                         "r"(tp_tag), "r"(__float_as_uint(acc[0][r + 1])), "r"(tp_tag)
                         : "memory");
        }
        if (tracing) p.trace[i == 0 ? 6 : 7] = (unsigned long long)(clock64() - c0);
        continue;
      }
    }
    if (lane < RPW && row < p.n) {
      if constexpr (EPI == EPI_PLAIN) {
        __nv_bfloat16 v = f_to_bf16(a0);
        if (p.bias != nullptr) v = __hadd(v, bias_v);
        p.y[row] = v;
      } else if constexpr (EPI == EPI_RESIDUAL) {
        p.y[row] = __hadd(res_v, f_to_bf16(a0));
      } else if constexpr (EPI == EPI_SILU_MUL) {
        const float g = round_bf16(a0);
        const __nv_bfloat16 sg = f_to_bf16(g / (1.f + expf(-g)));
        p.y[row] = __hmul(sg, f_to_bf16(a1));
      } else {  // EPI_TP_PUSH: one 8-byte {value, tag} store per rank, no fence needed
        const uint2 pk = make_uint2(__float_as_uint(a0), tp_tag);
        for (int q = 0; q < p.tp_world; ++q) st_volatile_u2(p.tp_push[q] + row, pk);
      }
    }
    if (tracing) p.trace[i == 0 ? 6 : 7] = (unsigned long long)(clock64() - c0);
  }

  if (p.pos_inc != nullptr && blockIdx.x == 0 && ctid == 0) *p.pos_inc += 1;
  if (p.trace != nullptr && blockIdx.x == 0 && ctid == 0) p.trace[2] = global_timer_ns();

}

// ------------------------------------------------------------------------------------------------------ dispatch
using KernelFn = void (*)(const GemvParams, const CUtensorMap);

template <int RPW>
KernelFn pick_kernel(int nseg, int pro, int epi) {
  if (nseg == 2) {
    if (epi != EPI_SILU_MUL) return nullptr;
    switch (pro) {
      case PRO_PLAIN: return gemv_stream_kernel<RPW, 2, PRO_PLAIN, EPI_SILU_MUL>;
      case PRO_RMSNORM: return gemv_stream_kernel<RPW, 2, PRO_RMSNORM, EPI_SILU_MUL>;
      case PRO_TP_RMSNORM: return gemv_stream_kernel<RPW, 2, PRO_TP_RMSNORM, EPI_SILU_MUL>;
    }
    return nullptr;
  }
#define B200_PICK(P, E) \
  if (pro == P && epi == E) return gemv_stream_kernel<RPW, 1, P, E>;
  B200_PICK(PRO_PLAIN, EPI_PLAIN)
  B200_PICK(PRO_PLAIN, EPI_RESIDUAL)
  B200_PICK(PRO_PLAIN, EPI_TP_PUSH)
  B200_PICK(PRO_RMSNORM, EPI_PLAIN)
  B200_PICK(PRO_TP_RMSNORM, EPI_PLAIN)
#undef B200_PICK
  return nullptr;
}

KernelFn pick(int rpw, int nseg, int pro, int epi) {
  switch (rpw) {
    case 1: return pick_kernel<1>(nseg, pro, epi);
    case 2: return pick_kernel<2>(nseg, pro, epi);
    case 4: return pick_kernel<4>(nseg, pro, epi);
  }
  return nullptr;
}

}  // namespace

int gemv_setup_attributes() {
  static std::once_flag once;
  static int rc = B200_OK;
  std::call_once(once, [] {
    const int rpws[3] = {1, 2, 4};
    for (int rpw : rpws)
      for (int nseg = 1; nseg <= 2; ++nseg)
        for (int pro = 0; pro < 3; ++pro)
          for (int epi = 0; epi < 4; ++epi) {
            KernelFn f = pick(rpw, nseg, pro, epi);
            if (!f) continue;
            cudaError_t e = cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, kGemvMaxSmem + 4096);
            // one carveout for every kernel of the token: a change of carveout between launches drains the SM and
            // would serialise the PDL overlap
            if (e == cudaSuccess)
              e = cudaFuncSetAttribute(f, cudaFuncAttributePreferredSharedMemoryCarveout,
                                       cudaSharedmemCarveoutMaxShared);
            if (e != cudaSuccess) {
              set_error("cudaFuncSetAttribute(gemv smem) failed: %s", cudaGetErrorString(e));
              rc = B200_ERR_CUDA;
              (void)cudaGetLastError();
              return;
            }
          }
  });
  return rc;
}

namespace {
struct Shape {
  int rpw, box_r, kb, k_pad, stage_bytes, fixed;
  int64_t rbs;
};
Shape pick_shape(int64_t n, int64_t k, int nseg, int num_sms) {
  // rows per warp: maximise the share of busy SM-slots, prefer bigger TMA boxes when within 3 %
  int best_rpw = 1;
  double best_eff = -1.0;
  const char* force = std::getenv("B200_GEMV_RPW");
  const int cands[3] = {4, 2, 1};
  for (int rpw : cands) {
    if (force && std::atoi(force) != rpw) continue;
    const int64_t rbs = (n + 8 * rpw - 1) / (8 * rpw);
    const int64_t g = std::min<int64_t>(rbs, num_sms);
    // time ∝ (row blocks of the busiest CTA) × rows per block; ideal = n / 8 / num_sms
    const double eff = ((double)n / 8.0 / (double)num_sms) / (double)(((rbs + g - 1) / g) * rpw);
    if (eff > best_eff + 0.03) {
      best_eff = eff;
      best_rpw = rpw;
    }
  }
  Shape s;
  s.rpw = best_rpw;
  s.box_r = 8 * s.rpw;
  s.rbs = (n + s.box_r - 1) / s.box_r;
  s.kb = kboxes(s.rpw, nseg);
  s.k_pad = (int)((k + kBoxK * s.kb - 1) / (kBoxK * s.kb) * (kBoxK * s.kb));
  s.stage_bytes = s.kb * nseg * s.box_r * kRowBytes;
  s.fixed = s.k_pad * 2 + 64;  // x vector + reduction scratch
  return s;
}
}  // namespace

int gemv_smem_wanted(int64_t n, int64_t k, int nseg, int num_sms) {
  const Shape s = pick_shape(n, k, nseg, num_sms);
  const int64_t g = std::min<int64_t>(s.rbs, num_sms);
  const int64_t stages = std::min<int64_t>(kMaxStages, ((s.rbs + g - 1) / g) * (s.k_pad / (kBoxK * s.kb)));
  return (int)std::min<int64_t>(kGemvMaxSmem, stages * (s.stage_bytes + 16) + s.fixed);
}

int gemv_make_plan(GemvPlan* plan, const void* W, int64_t rows_total, int64_t n, int64_t k, int nseg, int pro, int epi,
                   int num_sms, int smem_budget) {
  B200_CHECK_ARG(W != nullptr && n > 0 && k > 0, "gemv: null weight or empty shape (n=%lld k=%lld)", (long long)n,
                 (long long)k);
  B200_CHECK_ARG(k % 8 == 0, "gemv: k=%lld must be a multiple of 8 (16-byte rows)", (long long)k);
  B200_CHECK_ARG((reinterpret_cast<uintptr_t>(W) & 15) == 0, "gemv: W must be 16-byte aligned");
  B200_CHECK_ARG(nseg == 1 || nseg == 2, "gemv: nseg must be 1 or 2");
  B200_CHECK_ARG(n < (1ll << 30) && k < (1ll << 30) && rows_total >= n * nseg, "gemv: shape out of range");
  smem_budget = std::max(32 * 1024, std::min(smem_budget, kGemvMaxSmem));

  const char* cps = std::getenv("B200_GEMV_CTAS_PER_SM");
  const int ctas_per_sm = cps ? std::max(1, std::min(4, std::atoi(cps))) : 1;
  if (ctas_per_sm > 1) smem_budget = std::max(32 * 1024, smem_budget / ctas_per_sm);
  const Shape sh = pick_shape(n, k, nseg, num_sms);
  const int64_t g = std::min<int64_t>(sh.rbs, (int64_t)num_sms * ctas_per_sm);
  const int64_t most = ((sh.rbs + g - 1) / g) * (sh.k_pad / (kBoxK * sh.kb));  // stages of the busiest CTA
  int stages = (smem_budget - sh.fixed) / (sh.stage_bytes + 16);
  stages = (int)std::max<int64_t>(2, std::min<int64_t>(std::min<int64_t>(stages, kMaxStages), std::max<int64_t>(most, 2)));
  const char* fs = std::getenv("B200_GEMV_STAGES");
  if (fs) stages = std::max(2, std::min(std::atoi(fs), kMaxStages));

  *plan = GemvPlan{};
  plan->rpw = sh.rpw;
  plan->nseg = nseg;
  plan->pro = pro;
  plan->epi = epi;
  plan->grid = (int)g;
  plan->smem = stages * sh.stage_bytes + sh.k_pad * 2 + stages * 16 + 64;
  plan->p.n = (int)n;
  plan->p.k = (int)k;
  plan->p.k_pad = sh.k_pad;
  plan->p.seg_rows = (nseg == 2) ? (int)n : 0;
  plan->p.rowblocks = (int)sh.rbs;
  plan->p.stages = stages;
  plan->p.tp_world = 1;
  B200_CHECK_ARG(pick(sh.rpw, nseg, pro, epi) != nullptr, "gemv: no kernel for nseg=%d pro=%d epi=%d", nseg, pro, epi);
  B200_CHECK_ARG(plan->smem <= kGemvMaxSmem + 4096, "gemv: k=%lld needs %d bytes of shared memory", (long long)k,
                 plan->smem);
  // one box = box_r rows × kb·256 columns: kb·512 contiguous bytes per row instead of kb separate 512-byte boxes (HBM
  // sees 2 KB runs for one-row-per-warp kernels).  The inner box dimension is limited to 256 ELEMENTS, so W is described
  // in 2·kb-byte elements (TMA only moves bytes; out-of-bounds columns are zero-filled whatever the element type).
  return make_tmap_2d_bf16(&plan->tmap, W, rows_total, k, sh.box_r, kBoxK * sh.kb);
}

int gemv_plan_set_batch(GemvPlan* plan, int B) {
  B200_CHECK_ARG(B >= 1 && B <= kMaxBatch, "gemv: batch %d out of range (1 … %d)", B, kMaxBatch);
  B200_CHECK_ARG((plan->pro == PRO_PLAIN || plan->pro == PRO_RMSNORM) && plan->epi != EPI_TP_PUSH,
                 "gemv: batched decode is single-GPU only");
  plan->batch = B;
  plan->p.batch = B;
  plan->p.x_stride = plan->p.k;
  plan->p.y_stride = plan->p.n;
  plan->sub = 0;
  if (B == 1) return B200_OK;
  const int kb = kboxes(plan->rpw, plan->nseg);
  const int stage_bytes = kb * plan->nseg * 8 * plan->rpw * kRowBytes;
  // The activation vectors are staged next to the ring.  When they leave fewer than 4 stages (only the down projection
  // of the widest FFN: k = 14336, B > 4), halve the sequences per launch instead and stream W once per half — a
  // 2-stage ring costs more than the second weight pass.
  int mb = B;
  int stages = 0, fixed = 0;
  for (;; mb = (mb + 1) >> 1) {
    fixed = gemv_batch_fixed_smem(*plan, mb);
    stages = plan->p.stages;
    while (stages > 2 && stages * (stage_bytes + 16) + fixed > kGemvMaxSmem + 4096) --stages;
    const bool fits = stages * (stage_bytes + 16) + fixed <= kGemvMaxSmem + 4096;
    if (fits && (stages >= 4 || stages == plan->p.stages || mb == 2)) break;
    B200_CHECK_ARG(mb > 2, "gemv: %d activation vectors of k=%d do not fit in shared memory next to the ring", B,
                   plan->p.k);
  }
  if (mb < B) plan->sub = mb;
  plan->p.stages = stages;
  plan->smem = stages * (stage_bytes + 16) + fixed;
  return B200_OK;
}

int gemv_launch(const GemvPlan& plan, cudaStream_t stream, bool pdl) {
  if (plan.batch > 1) return gemv_batch_launch(plan, stream, pdl);
  KernelFn f = pick(plan.rpw, plan.nseg, plan.pro, plan.epi);
  if (!f) {
    set_error("gemv: no kernel instantiation (nseg=%d pro=%d epi=%d)", plan.nseg, plan.pro, plan.epi);
    return B200_ERR_INVALID;
  }
  B200_CUDA(launch_pdl(f, dim3(plan.grid), dim3(kThreads), (size_t)plan.smem, stream, pdl, plan.p, plan.tmap));
  return B200_OK;
}

}  // namespace b200
