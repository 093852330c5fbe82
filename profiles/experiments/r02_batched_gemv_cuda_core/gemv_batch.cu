// gemv_batch.cu — the weight-streaming GEMV for a BATCH of up to 8 decode sequences: W is streamed from HBM once and
// dotted with every sequence's activation vector.
//
// Why: GPTEngine::generateSync feeds a batch of left-padded prompts (examples/inference/main.cpp:12-17 has four) and the
// reference's Linear then runs one cuBLAS GEMM with m = B  [ref: src/engine/GPTEngine.cpp:154-174;
// third_party/TinyTorch/src/Operation/OpLinalg.cpp:152-203,244-277].  A batch-1 engine per sequence costs B full weight
// passes per step; decode is weight-bound, so B sequences should cost (almost) one.
//
// Same structure as gemv_stream_kernel (gemv.cu): one TMA producer thread, a ring of 16 KB stages, 8 consumer warps,
// warp w owns rows w·RPW… of every row block.  Differences: the B activation vectors are staged side by side in shared
// memory ([MB][k_pad] bf16, MB = 2 / 4 / 8 ≥ B), every 16-byte weight vector a lane loads is multiplied with all of
// them (MB × the FMAs per byte of shared-memory traffic: the ALU-bound consumer gets cheaper per sequence, not dearer),
// and the fused prologue / epilogue run per sequence.  Per sequence the arithmetic is EXACTLY gemv_stream_kernel's —
// same partition of the RMSNorm sum of squares, same FMA order per row, same rounding points — so a batched step
// reproduces B independent batch-1 steps bit for bit (tests/test_batch_gpu.py).  Single-GPU prologues / epilogues only.
#include "gemv.cuh"

#include <mutex>

namespace b200 {

using namespace gemvk;

namespace {

template <int RPW, int NSEG, int PRO, int EPI, int MB>
__global__ void __launch_bounds__(kThreads, 1)
gemv_batch_kernel(const GemvParams p, const __grid_constant__ CUtensorMap tmap) {
  constexpr int kBoxR = kNW * RPW;
  constexpr int KB = kboxes(RPW, NSEG);
  constexpr int kBoxBytes = kBoxR * kRowBytes;
  constexpr int kStageBytes = KB * NSEG * kBoxBytes;
  static_assert(PRO == PRO_PLAIN || PRO == PRO_RMSNORM, "single-GPU prologues only");
  static_assert(EPI == EPI_PLAIN || EPI == EPI_RESIDUAL || EPI == EPI_SILU_MUL, "single-GPU epilogues only");

  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* const stage_base = smem;
  __nv_bfloat16* const xs = reinterpret_cast<__nv_bfloat16*>(smem + (size_t)p.stages * kStageBytes);   // [MB][k_pad]
  uint64_t* const full = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(xs) + (size_t)MB * p.k_pad * 2);
  uint64_t* const empty = full + p.stages;
  float* const red = reinterpret_cast<float*>(empty + p.stages);                                          // [MB][kNW]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int ksteps = p.k_pad / (kBoxK * KB);
  const int my_rbs = (p.rowblocks - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int nb = p.batch;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], kNW);
    }
    fence_mbar_init();
  }
  __syncthreads();
  pdl_trigger();

  if (warp == kNW) {
    if (lane == 0) {   // producer: gemv_stream_kernel's
      tma_prefetch_desc(&tmap);
      int s = 0;
      uint32_t ph = 1;
      for (int i = 0; i < my_rbs; ++i) {
        const int row0 = ((int)blockIdx.x + i * (int)gridDim.x) * kBoxR;
        for (int ks = 0; ks < ksteps; ++ks) {
          mbar_wait(&empty[s], ph);
          mbar_arrive_expect_tx(&full[s], kStageBytes);
          uint8_t* dst = stage_base + (size_t)s * kStageBytes;
#pragma unroll
          for (int seg = 0; seg < NSEG; ++seg)
            tma_load_2d(dst + seg * (KB * kBoxBytes), &tmap, ks * kBoxK, seg * p.seg_rows + row0, &full[s]);
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
  const int ctid = threadIdx.x;
  const int nvec = p.k >> 3, nvec_pad = p.k_pad >> 3;
  pdl_wait();   // the producer kernel's output (x, residual) is complete and visible from here on

  // ---- stage the activation vectors; with the RMSNorm prologue: raw x first, sums of squares in gemv_stream_kernel's
  // partition (thread t: vectors t, t + 256, … in order; xor-shuffle; warps in order), then every thread rescales the
  // vectors it staged itself
#pragma unroll
  for (int b = 0; b < MB; ++b) {
    if (b < nb) {
      const uint4* xg = reinterpret_cast<const uint4*>(p.x + (size_t)b * p.x_stride);
      uint4* xv = reinterpret_cast<uint4*>(xs + (size_t)b * p.k_pad);
      float ss = 0.f;
      for (int i = ctid; i < nvec_pad; i += kConsumers) {
        uint4 q = make_uint4(0, 0, 0, 0);
        if (i < nvec) {
          q = xg[i];
          if constexpr (PRO == PRO_RMSNORM) {
            float xf[8];
            unpack8(q, xf);
#pragma unroll
            for (int e = 0; e < 8; ++e) ss += xf[e] * xf[e];
          }
        }
        xv[i] = q;
      }
      if constexpr (PRO == PRO_RMSNORM) {
        ss = warp_sum(ss);
        if (lane == 0) red[b * kNW + warp] = ss;
      }
    }
  }
  if constexpr (PRO == PRO_RMSNORM) {
    named_bar_sync(1, kConsumers);
    const uint4* wg = reinterpret_cast<const uint4*>(p.norm_w);
#pragma unroll
    for (int b = 0; b < MB; ++b) {
      if (b < nb) {
        float tot = 0.f;
#pragma unroll
        for (int w = 0; w < kNW; ++w) tot += red[b * kNW + w];
        const float inv = rsqrtf(tot / (float)p.k + p.eps);
        uint4* xv = reinterpret_cast<uint4*>(xs + (size_t)b * p.k_pad);
        for (int i = ctid; i < nvec; i += kConsumers) {
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
    }
  }
  named_bar_sync(1, kConsumers);

  // ---------------------------------------------------------------------------------------------- main k loop
  int s = 0;
  uint32_t ph = 0;
  const uint8_t* const my_rows = stage_base + (size_t)(warp * RPW) * (KB * kRowBytes) + lane * 16;
  for (int i = 0; i < my_rbs; ++i) {
    const int row_base = ((int)blockIdx.x + i * (int)gridDim.x) * kBoxR + warp * RPW;
    const int row = row_base + lane;
    const bool mine = lane < RPW && row < p.n;
    // operands of the epilogue are requested now so that their latency hides behind the k loop
    __nv_bfloat16 res_v[MB];
    __nv_bfloat16 bias_v = f_to_bf16(0.f);
#pragma unroll
    for (int b = 0; b < MB; ++b) {
      res_v[b] = f_to_bf16(0.f);
      if constexpr (EPI == EPI_RESIDUAL) {
        if (mine && b < nb) res_v[b] = p.residual[(size_t)b * p.y_stride + row];
      }
    }
    if constexpr (EPI == EPI_PLAIN) {
      if (mine && p.bias != nullptr) bias_v = p.bias[row];
    }

    float acc[MB][NSEG][RPW], acc_b[MB][NSEG][RPW];
#pragma unroll
    for (int b = 0; b < MB; ++b)
#pragma unroll
      for (int seg = 0; seg < NSEG; ++seg)
#pragma unroll
        for (int r = 0; r < RPW; ++r) acc[b][seg][r] = acc_b[b][seg][r] = 0.f;

    for (int ks = 0; ks < ksteps; ++ks) {
      mbar_wait(&full[s], ph);
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
      for (int b = 0; b < MB; ++b) {
        if (b < nb) {
          const __nv_bfloat16* xb = xs + (size_t)b * p.k_pad + ks * (KB * kBoxK) + lane * 8;
#pragma unroll
          for (int kb = 0; kb < KB; ++kb) {
            float xf[8];
            unpack8(*reinterpret_cast<const uint4*>(xb + kb * kBoxK), xf);
#pragma unroll
            for (int seg = 0; seg < NSEG; ++seg)
#pragma unroll
              for (int r = 0; r < RPW; ++r) dot8x2(wv[kb][seg][r], xf, acc[b][seg][r], acc_b[b][seg][r]);
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[s]);
      if (++s == p.stages) {
        s = 0;
        ph ^= 1;
      }
    }

#pragma unroll
    for (int b = 0; b < MB; ++b) {
      if (b < nb) {
#pragma unroll
        for (int seg = 0; seg < NSEG; ++seg)
#pragma unroll
          for (int r = 0; r < RPW; ++r) acc[b][seg][r] = warp_sum(acc[b][seg][r] + acc_b[b][seg][r]);
        // lane r finishes row r of this warp
        float a0 = acc[b][0][0], a1 = acc[b][NSEG - 1][0];
#pragma unroll
        for (int r = 1; r < RPW; ++r) {
          if (lane == r) {
            a0 = acc[b][0][r];
            a1 = acc[b][NSEG - 1][r];
          }
        }
        if (mine) {
          __nv_bfloat16* y = p.y + (size_t)b * p.y_stride;
          if constexpr (EPI == EPI_PLAIN) {
            __nv_bfloat16 v = f_to_bf16(a0);
            if (p.bias != nullptr) v = __hadd(v, bias_v);
            y[row] = v;
          } else if constexpr (EPI == EPI_RESIDUAL) {
            y[row] = __hadd(res_v[b], f_to_bf16(a0));
          } else {  // EPI_SILU_MUL
            const float g = round_bf16(a0);
            const __nv_bfloat16 sg = f_to_bf16(g / (1.f + expf(-g)));
            y[row] = __hmul(sg, f_to_bf16(a1));
          }
        }
      }
    }
  }
  if (p.pos_inc != nullptr && blockIdx.x == 0 && ctid == 0) *p.pos_inc += 1;
}

using KernelFn = void (*)(const GemvParams, const CUtensorMap);

template <int RPW, int MB>
KernelFn pick_rm(int nseg, int pro, int epi) {
  if (nseg == 2) return (pro == PRO_RMSNORM && epi == EPI_SILU_MUL) ? gemv_batch_kernel<RPW, 2, PRO_RMSNORM, EPI_SILU_MUL, MB> : nullptr;
  if (pro == PRO_RMSNORM && epi == EPI_PLAIN) return gemv_batch_kernel<RPW, 1, PRO_RMSNORM, EPI_PLAIN, MB>;
  if (pro == PRO_PLAIN && epi == EPI_RESIDUAL) return gemv_batch_kernel<RPW, 1, PRO_PLAIN, EPI_RESIDUAL, MB>;
  return nullptr;
}
template <int RPW>
KernelFn pick_r(int nseg, int pro, int epi, int mb) {
  switch (mb) {
    case 2: return pick_rm<RPW, 2>(nseg, pro, epi);
    case 4: return pick_rm<RPW, 4>(nseg, pro, epi);
    case 8: return pick_rm<RPW, 8>(nseg, pro, epi);
  }
  return nullptr;
}
KernelFn pick(int rpw, int nseg, int pro, int epi, int mb) {
  switch (rpw) {
    case 1: return pick_r<1>(nseg, pro, epi, mb);
    case 2: return pick_r<2>(nseg, pro, epi, mb);
    case 4: return pick_r<4>(nseg, pro, epi, mb);
  }
  return nullptr;
}

}  // namespace

int gemv_batch_setup_attributes() {
  static std::once_flag once;
  static int rc = B200_OK;
  std::call_once(once, [] {
    const int rpws[3] = {1, 2, 4}, mbs[3] = {2, 4, 8};
    for (int rpw : rpws)
      for (int mb : mbs)
        for (int nseg = 1; nseg <= 2; ++nseg)
          for (int pro = 0; pro < 2; ++pro)
            for (int epi = 0; epi < 3; ++epi) {
              KernelFn f = pick(rpw, nseg, pro, epi, mb);
              if (!f) continue;
              cudaError_t e = cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, kGemvMaxSmem + 4096);
              if (e == cudaSuccess)
                e = cudaFuncSetAttribute(f, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
              if (e != cudaSuccess) {
                set_error("cudaFuncSetAttribute(batched gemv smem) failed: %s", cudaGetErrorString(e));
                rc = B200_ERR_CUDA;
                (void)cudaGetLastError();
                return;
              }
            }
  });
  return rc;
}

int gemv_batch_launch(const GemvPlan& plan, cudaStream_t stream, bool pdl) {
  const int per = plan.sub > 0 ? plan.sub : plan.batch;   // sequences per launch (gemv_plan_set_batch)
  const int mb = per <= 2 ? 2 : per <= 4 ? 4 : 8;
  KernelFn f = pick(plan.rpw, plan.nseg, plan.pro, plan.epi, mb);
  if (!f) {
    set_error("batched gemv: no kernel instantiation (rpw=%d nseg=%d pro=%d epi=%d batch=%d)", plan.rpw, plan.nseg, plan.pro,
              plan.epi, plan.batch);
    return B200_ERR_INVALID;
  }
  for (int b0 = 0; b0 < plan.batch; b0 += per) {
    GemvParams p = plan.p;
    p.batch = plan.batch - b0 < per ? plan.batch - b0 : per;
    p.x += (size_t)b0 * p.x_stride;
    p.y += (size_t)b0 * p.y_stride;
    if (p.residual) p.residual += (size_t)b0 * p.y_stride;
    const bool last = b0 + per >= plan.batch;
    if (!last) p.pos_inc = nullptr;                       // the position advances once, after the last launch
    // every launch waits on its predecessor (griddepcontrol.wait), so the chain stays ordered through the halves
    B200_CUDA(launch_pdl(f, dim3(plan.grid), dim3(kThreads), (size_t)plan.smem, stream, pdl, p, plan.tmap));
  }
  return B200_OK;
}

}  // namespace b200
