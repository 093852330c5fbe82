// gemv_batch.cu — the weight-streaming GEMV for a BATCH of up to 8 decode sequences, on the tensor cores:
// y[b] = W · x[b] with ONE pass over W, the B activation vectors being the n = 8 operand of mma.sync.m16n8k16.
//
// Why: GPTEngine::generateSync feeds a batch of left-padded prompts (examples/inference/main.cpp:12-17 has four) and the
// reference's Linear then runs one cuBLAS GEMM with m = B  [ref: src/engine/GPTEngine.cpp:154-174;
// third_party/TinyTorch/src/Operation/OpLinalg.cpp:152-203,244-277].  A batch-1 engine per sequence costs B full weight
// passes per step; decode is weight-bound, so B sequences should cost (almost) one.  [B ≤ 8, k] × [k, n] is GEMM-shaped
// work: on the tensor cores the consumer side costs ~25 warp instructions per 16 KB stage whatever B is, and the step
// stays bound by the weight stream.  (A first version repeated gemv_stream_kernel's CUDA-core arithmetic per sequence —
// bit-identical to batch 1, but B × the FMAs per weight vector: a B = 8 step cost 4.1-4.5 × a batch-1 step.  It is kept,
// with its numbers, under profiles/experiments/r02_batched_gemv_cuda_core/.)  The price: the k-sum of a row is taken in
// tensor-core order (fp32 accumulation of exact bf16 products, 8 k-slices added in warp order), so a batched step agrees
// with batch-1 steps to summation-order noise instead of bit for bit — exactly the relation between the reference's own
// m = 1 and m = B paths (profiles/r02_ref_cuda_parity.json, "reference_cuda_decode_vs_its_own_batched_path").
//
// Structure: gemv_stream_kernel's producer, ring and plan geometry, unchanged (one TMA thread, 16 KB stages of
// [NSEG][8·RPW rows][KB·256 k] bf16).  The 8 consumer warps split every stage along K — warp w takes the KB·32 k-elements
// w·KB·32 … of ALL the stage's rows — and keep one 16 × 8 fp32 accumulator tile per 16 stage rows.  At the end of a row
// block the 8 k-slices are added through shared memory in warp order and the fused epilogue (bias / residual / SiLU·mul,
// the batch-1 kernel's rounding points) runs on [rows][B] outputs.
//
// Fragment trick (no ldmatrix, no swizzle, no repacking of W): a dot product does not care in which order k is visited
// as long as A and B agree.  Lane (g = lane/4, t = lane%4) loads ONE 16-byte vector W[row g][k0 + 8t … 8t+7] and ONE
// x[seq g][k0 + 8t … 8t+7]; the first halves feed one MMA (logical k {2t, 2t+1, 2t+8, 2t+9} := physical {8t … 8t+3}), the
// second halves the next — two m16n8k16 per pair of LDS.128.  x rows are padded by 64 bytes so that the two sequences of
// a quarter-warp hit different banks; W rows (dense TMA boxes, stride ≡ 0 mod 128) take a 2-way conflict, which at
// ≤ 23 bytes/clk/SM of HBM feed is far from the shared-memory limit.
#include "gemv.cuh"

#include <mutex>

namespace b200 {

using namespace gemvk;

namespace {

// D(16×8, fp32) += A(16×16, bf16, row) · B(16×8, bf16, col)
__device__ __forceinline__ void mma_16816(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                          uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

constexpr int kXPad = 32;   // bf16 elements (64 bytes) between the staged activation vectors: conflict-free B fragments

template <int RPW, int NSEG, int PRO, int EPI>
__global__ void __launch_bounds__(kThreads, 1)
gemv_batch_kernel(const GemvParams p, const __grid_constant__ CUtensorMap tmap) {
  constexpr int kBoxR = kNW * RPW;                  // rows of one segment in a stage
  constexpr int KB = kboxes(RPW, NSEG);
  constexpr int kBoxBytes = kBoxR * kRowBytes;
  constexpr int kStageBytes = KB * NSEG * kBoxBytes;
  constexpr int kRowsStage = NSEG * kBoxR;          // 8, 16, 32 or 64 stage rows = A rows
  constexpr int MT = (kRowsStage + 15) / 16;        // m16 tiles
  constexpr bool kHalfTile = kRowsStage == 8;       // rows 8-15 of the only tile do not exist: a1 = a3 = 0
  constexpr int kSliceK = KB * 32;                  // k-elements of a stage one warp owns
  static_assert(PRO == PRO_PLAIN || PRO == PRO_RMSNORM, "single-GPU prologues only");
  static_assert(EPI == EPI_PLAIN || EPI == EPI_RESIDUAL || EPI == EPI_SILU_MUL, "single-GPU epilogues only");

  const int nb = p.batch;
  const int xstride = p.k_pad + kXPad;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* const stage_base = smem;
  __nv_bfloat16* const xs = reinterpret_cast<__nv_bfloat16*>(smem + (size_t)p.stages * kStageBytes);   // [nb][xstride]
  uint64_t* const full = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(xs) + (size_t)nb * xstride * 2);
  uint64_t* const empty = full + p.stages;
  float* const red = reinterpret_cast<float*>(empty + p.stages);               // red[b] = 1/rms of sequence b (64 floats reserved)
  float* const tiles = red + kMaxBatch * kNW;                                                              // [kNW][MT][16][8]
  __nv_bfloat16* const wns = reinterpret_cast<__nv_bfloat16*>(tiles + kNW * MT * 16 * 8);   // RMSNorm weight [k_pad] (PRO_RMSNORM)

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int ksteps = p.k_pad / (kBoxK * KB);
  const int my_rbs = (p.rowblocks - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

  if (p.trace != nullptr && blockIdx.x == 0 && threadIdx.x == 0) p.trace[0] = global_timer_ns();   // B200_TRACE=1
  // The RMSNorm weight does not depend on the producer kernel: its loads (k ≤ 4096 stays in registers) are issued
  // first, before the dependency wait (like gemv_stream_kernel).
  constexpr int kMaxHoist = 2;
  uint4 wn[kMaxHoist];
  if constexpr (PRO == PRO_RMSNORM) {
    const uint4* wg = reinterpret_cast<const uint4*>(p.norm_w);
#pragma unroll
    for (int j = 0; j < kMaxHoist; ++j) {
      const int i = (int)threadIdx.x + j * kConsumers;
      wn[j] = (warp < kNW && i < (p.k >> 3)) ? wg[i] : make_uint4(0, 0, 0, 0);
    }
  }
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
      const uint64_t pol = l2_policy_evict_first();   // every weight byte is read once per token
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
  const int ctid = threadIdx.x;
  const int nvec = p.k >> 3, nvec_pad = p.k_pad >> 3;
  pdl_wait();   // the producer kernel's output (x, residual) is complete and visible from here on
  const bool tracing = p.trace != nullptr && blockIdx.x == 0 && ctid == 0;
  long long c0 = 0;
  if (tracing) {
    p.trace[1] = global_timer_ns();
    c0 = clock64();
  }

  // ---- stage the activation vectors
  if constexpr (PRO == PRO_PLAIN) {
    // vector index outside, sequence inside and unrolled: up to 8 independent 16-byte loads in flight per thread (one L2
    // round trip per pass, not one per sequence)
    for (int i = ctid; i < nvec_pad; i += kConsumers) {
      uint4 q[kMaxBatch];
#pragma unroll
      for (int b = 0; b < kMaxBatch; ++b) {
        q[b] = make_uint4(0, 0, 0, 0);
        if (b < nb && i < nvec) q[b] = reinterpret_cast<const uint4*>(p.x + (size_t)b * p.x_stride)[i];
      }
#pragma unroll
      for (int b = 0; b < kMaxBatch; ++b)
        if (b < nb) reinterpret_cast<uint4*>(xs + (size_t)b * xstride)[i] = q[b];
    }
  } else {
    // RMSNorm, one WARP per sequence: lane-strided loads, sum of squares by shuffle only, no cross-warp reduction; the
    // norm weight goes through shared memory.  The partition of the sum of squares differs from the batch-1 kernel's,
    // which is inside the summation-order tolerance this kernel already has.
    // OPEN (profiles/r02_trace_batch_prologue.log): this prologue takes 2.5-2.7 µs per kernel against 0.8 µs at batch 1
    // and 0.65 µs for the plain staging above.  Ruled out by measurement: the norm-weight load (hoisting it above the
    // wait or above the first barrier, L2 evict_last, a copy on hot pages — no change).  Diagnosis from the stamps: with
    // k/8/32 < 8 the unroll-by-8 body of the x loop below never executes and the remainder loop runs its 3-4
    // iterations as dependent load → use round trips (~0.4 µs each).  Issuing the lane's loads together is the next
    // change; it was not made because it could not be validated on hardware any more this round.
    uint4* const wv = reinterpret_cast<uint4*>(wns);
    for (int b = warp; b < nb; b += kNW) {
      const uint4* xg = reinterpret_cast<const uint4*>(p.x + (size_t)b * p.x_stride);
      uint4* xv = reinterpret_cast<uint4*>(xs + (size_t)b * xstride);
      float ss = 0.f;
#pragma unroll 8
      for (int i = lane; i < nvec; i += 32) {
        const uint4 q = xg[i];
        float xf[8];
        unpack8(q, xf);
#pragma unroll
        for (int e = 0; e < 8; ++e) ss += xf[e] * xf[e];
        xv[i] = q;
      }
      ss = warp_sum(ss);
      if (lane == 0) red[b] = rsqrtf(ss / (float)p.k + p.eps);
    }
#pragma unroll
    for (int j = 0; j < kMaxHoist; ++j) {
      const int i = ctid + j * kConsumers;
      if (i < nvec) wv[i] = wn[j];
    }
    for (int i = ctid + kMaxHoist * kConsumers; i < nvec; i += kConsumers) wv[i] = reinterpret_cast<const uint4*>(p.norm_w)[i];
    named_bar_sync(1, kConsumers);
    for (int b = warp; b < nb; b += kNW) {
      uint4* xv = reinterpret_cast<uint4*>(xs + (size_t)b * xstride);
      const float inv = red[b];
#pragma unroll 4
      for (int i = lane; i < nvec; i += 32) {
        float xf[8], wf[8];
        unpack8(xv[i], xf);
        unpack8(wv[i], wf);
        uint4 o;  // reference order: normed = x * inv; normed *= w; one rounding
        o.x = pack2(xf[0] * inv * wf[0], xf[1] * inv * wf[1]);
        o.y = pack2(xf[2] * inv * wf[2], xf[3] * inv * wf[3]);
        o.z = pack2(xf[4] * inv * wf[4], xf[5] * inv * wf[5]);
        o.w = pack2(xf[6] * inv * wf[6], xf[7] * inv * wf[7]);
        xv[i] = o;
      }
      for (int i = nvec + lane; i < nvec_pad; i += 32) xv[i] = make_uint4(0, 0, 0, 0);
    }
  }
  named_bar_sync(1, kConsumers);
  if (tracing) p.trace[3] = global_timer_ns();

  // ---------------------------------------------------------------------------------------------- main k loop
  const int g = lane >> 2, t = lane & 3;
  // byte offset of stage row j (segment-major: [seg][kBoxR rows][KB·256 k]) — rows g and g + 8 of every m16 tile
  int row_off[MT][2];
#pragma unroll
  for (int mt = 0; mt < MT; ++mt)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int j = mt * 16 + h * 8 + g;
      row_off[mt][h] = (j / kBoxR) * (KB * kBoxBytes) + (j % kBoxR) * (KB * kRowBytes);
    }
  const int kslice = warp * kSliceK + t * 8;                         // my k offset inside a stage (elements)
  const bool has_x = g < nb;
  const __nv_bfloat16* const my_x = xs + (size_t)(has_x ? g : 0) * xstride + kslice;

  // epilogue work item of this thread: output (row r of the row block, sequence b) — [kBoxR][8] items over 256 threads
  constexpr int kItems = kBoxR * 8;
  const int it_r = ctid >> 3, it_b = ctid & 7;
  const bool it_on = ctid < kItems && it_b < nb;

  int s = 0;
  uint32_t ph = 0;
  for (int i = 0; i < my_rbs; ++i) {
    const int row = ((int)blockIdx.x + i * (int)gridDim.x) * kBoxR + it_r;
    const bool mine = it_on && row < p.n;
    // operands of the epilogue are requested now so that their latency hides behind the k loop
    __nv_bfloat16 res_v = f_to_bf16(0.f), bias_v = f_to_bf16(0.f);
    if constexpr (EPI == EPI_RESIDUAL) {
      if (mine) res_v = p.residual[(size_t)it_b * p.y_stride + row];
    }
    if constexpr (EPI == EPI_PLAIN) {
      if (mine && p.bias != nullptr) bias_v = p.bias[row];
    }

    float acc[MT][4];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[mt][e] = 0.f;

    for (int ks = 0; ks < ksteps; ++ks) {
      mbar_wait(&full[s], ph);
      if (tracing && i == 0 && ks == 0) p.trace[4] = (unsigned long long)(clock64() - c0);
      const uint8_t* st = stage_base + (size_t)s * kStageBytes + kslice * 2;
      const __nv_bfloat16* xk = my_x + ks * (KB * kBoxK);
#pragma unroll
      for (int c = 0; c < KB; ++c) {
        uint4 xv = make_uint4(0, 0, 0, 0);
        if (has_x) xv = *reinterpret_cast<const uint4*>(xk + c * 32);
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
          const uint4 lo = *reinterpret_cast<const uint4*>(st + row_off[mt][0] + c * 64);
          uint4 hi = make_uint4(0, 0, 0, 0);
          if constexpr (!kHalfTile) hi = *reinterpret_cast<const uint4*>(st + row_off[mt][1] + c * 64);
          mma_16816(acc[mt], lo.x, hi.x, lo.y, hi.y, xv.x, xv.y);
          mma_16816(acc[mt], lo.z, hi.z, lo.w, hi.w, xv.z, xv.w);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[s]);
      if (++s == p.stages) {
        s = 0;
        ph ^= 1;
      }
    }

    if (tracing && i == 0) p.trace[5] = (unsigned long long)(clock64() - c0);
    // ---- add the 8 k-slices in warp order, then the fused epilogue on [kBoxR rows][nb sequences]
    named_bar_sync(1, kConsumers);            // the previous row block's epilogue has read `tiles`
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
      float* tw = tiles + ((warp * MT + mt) * 16) * 8;
      *reinterpret_cast<float2*>(tw + g * 8 + 2 * t) = make_float2(acc[mt][0], acc[mt][1]);
      if constexpr (!kHalfTile) *reinterpret_cast<float2*>(tw + (g + 8) * 8 + 2 * t) = make_float2(acc[mt][2], acc[mt][3]);
    }
    named_bar_sync(1, kConsumers);
    if (mine) {
      float a[NSEG];
#pragma unroll
      for (int seg = 0; seg < NSEG; ++seg) {
        const int j = seg * kBoxR + it_r;
        const float* tp = tiles + ((j >> 4) * 16 + (j & 15)) * 8 + it_b;
        float v = 0.f;
#pragma unroll
        for (int w = 0; w < kNW; ++w) v += tp[w * MT * 16 * 8];
        a[seg] = v;
      }
      __nv_bfloat16* y = p.y + (size_t)it_b * p.y_stride;
      if constexpr (EPI == EPI_PLAIN) {
        __nv_bfloat16 v = f_to_bf16(a[0]);
        if (p.bias != nullptr) v = __hadd(v, bias_v);
        y[row] = v;
      } else if constexpr (EPI == EPI_RESIDUAL) {
        y[row] = __hadd(res_v, f_to_bf16(a[0]));
      } else {  // EPI_SILU_MUL
        const float gt = round_bf16(a[0]);
        const __nv_bfloat16 sg = f_to_bf16(gt / (1.f + expf(-gt)));
        y[row] = __hmul(sg, f_to_bf16(a[NSEG - 1]));
      }
    }
    if (tracing) p.trace[i == 0 ? 6 : 7] = (unsigned long long)(clock64() - c0);
  }
  if (tracing) p.trace[2] = global_timer_ns();
  if (p.pos_inc != nullptr && blockIdx.x == 0 && ctid == 0) *p.pos_inc += 1;
}

using KernelFn = void (*)(const GemvParams, const CUtensorMap);

template <int RPW>
KernelFn pick_r(int nseg, int pro, int epi) {
  if (nseg == 2) return (pro == PRO_RMSNORM && epi == EPI_SILU_MUL) ? gemv_batch_kernel<RPW, 2, PRO_RMSNORM, EPI_SILU_MUL> : nullptr;
  if (pro == PRO_RMSNORM && epi == EPI_PLAIN) return gemv_batch_kernel<RPW, 1, PRO_RMSNORM, EPI_PLAIN>;
  if (pro == PRO_PLAIN && epi == EPI_RESIDUAL) return gemv_batch_kernel<RPW, 1, PRO_PLAIN, EPI_RESIDUAL>;
  return nullptr;
}
KernelFn pick(int rpw, int nseg, int pro, int epi) {
  switch (rpw) {
    case 1: return pick_r<1>(nseg, pro, epi);
    case 2: return pick_r<2>(nseg, pro, epi);
    case 4: return pick_r<4>(nseg, pro, epi);
  }
  return nullptr;
}

}  // namespace

// Shared memory beside the ring: `nb` padded activation vectors, the RMSNorm partials and the k-slice tiles.
int gemv_batch_fixed_smem(const GemvPlan& plan, int nb) {
  const int mt = (plan.nseg * kNW * plan.rpw + 15) / 16;
  return nb * (plan.p.k_pad + kXPad) * 2 + kMaxBatch * kNW * 4 + kNW * mt * 16 * 8 * 4 + 64 +
         (plan.pro == PRO_RMSNORM ? plan.p.k_pad * 2 : 0);
}

int gemv_batch_setup_attributes() {
  static std::once_flag once;
  static int rc = B200_OK;
  std::call_once(once, [] {
    const int rpws[3] = {1, 2, 4};
    for (int rpw : rpws)
      for (int nseg = 1; nseg <= 2; ++nseg)
        for (int pro = 0; pro < 2; ++pro)
          for (int epi = 0; epi < 3; ++epi) {
            KernelFn f = pick(rpw, nseg, pro, epi);
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
  KernelFn f = pick(plan.rpw, plan.nseg, plan.pro, plan.epi);
  if (!f) {
    set_error("batched gemv: no kernel instantiation (rpw=%d nseg=%d pro=%d epi=%d)", plan.rpw, plan.nseg, plan.pro, plan.epi);
    return B200_ERR_INVALID;
  }
  for (int b0 = 0; b0 < plan.batch; b0 += per) {
    GemvParams p = plan.p;
    p.batch = plan.batch - b0 < per ? plan.batch - b0 : per;
    p.x += (size_t)b0 * p.x_stride;
    p.y += (size_t)b0 * p.y_stride;
    if (p.residual) p.residual += (size_t)b0 * p.y_stride;
    if (b0 + per < plan.batch) p.pos_inc = nullptr;       // the position advances once, after the last launch
    B200_CUDA(launch_pdl(f, dim3(plan.grid), dim3(kThreads), (size_t)plan.smem, stream, pdl, p, plan.tmap));
  }
  return B200_OK;
}

}  // namespace b200
