// gemm.cu — the batched PREFILL GEMM on the 5th-generation tensor cores (tcgen05 + TMEM + TMA), sm_100a only.
//
//   C[M,N] = bf16( A[M,K] · B[N,K]ᵀ )      A = activations of M prompt tokens, B = Linear.weight [out,in] (K-contiguous)
//
// Replaces, for m > 1, op::matmul → cublasGemmStridedBatchedEx / cublasGemmEx
//   [ref: third_party/TinyTorch/src/Operation/OpLinalg.cpp:244-277, OpLinalgCuda.cuh:193-216,276-293]
// with the reference's numerics: bf16 × bf16 products, fp32 accumulation, ONE rounding to bf16 (bias / residual are
// applied by separate kernels in the prefill path, exactly where the reference rounds).  Decode (m = 1) never comes
// here — it is HBM-bound and runs on the CUDA-core GEMV (gemv.cu).
//
// Structure (one 128×128 output tile per CTA, 192 threads, warp-specialised):
//   warp 0   TMA producer: per 64-column k-block one 128×64 box of A and one of B (128-byte swizzle) into a 4-stage
//            shared-memory ring, completion on `full[s]` (expect_tx)
//   warp 1   MMA issuer: one elected thread issues 4 × tcgen05.mma.cta_group::1.kind::f16 (M128 N128 K16) per k-block,
//            accumulator in TMEM (128 lanes × 128 fp32 columns); tcgen05.commit releases the stage to the producer
//            (`empty[s]`) and, after the last k-block, signals the epilogue (`tmem_full`)
//   warps 2-5 epilogue: tcgen05.ld 32x32b.x16 (each warp owns the TMEM lane quarter warp_id % 4), fp32 → bf16,
//            32-byte stores of 16 consecutive columns per thread
// TMA zero-fills out-of-range rows/columns, so M, N, K need not be multiples of the tile.
#include "common.cuh"
#include "ops.cuh"

#include <algorithm>
#include <cstdlib>
#include <mutex>

namespace b200 {

namespace {

constexpr int kBM = 128, kBN = 128, kBK = 64;   // tile; kBK bf16 = 128 bytes = one swizzle row
constexpr int kGemmStages = 4;
constexpr int kGemmThreads = 192;
constexpr int kTileBytes = kBM * kBK * 2;        // 16 KB per operand per stage
constexpr int kTmemCols = 128;
constexpr int kGemmSmem = kGemmStages * 2 * kTileBytes + 1024 /*align*/ + 256 /*barriers*/;

// kind::f16 instruction descriptor: D = F32 (bits 4-5 = 1), A = B = BF16 (bits 7-9, 10-12 = 1), both K-major
// (bits 15, 16 = 0), N >> 3 at bits 17-22, M >> 4 at bits 24-28.
constexpr uint32_t kIdesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kBN >> 3) << 17) | ((uint32_t)(kBM >> 4) << 24);

// shared-memory matrix descriptor, K-major, SWIZZLE_128B: start address >> 4 (bits 0-13), LBO = 0 (bits 16-29, unused
// for swizzled K-major), SBO = 1024 B >> 4 (bits 32-45: 8 rows × 128 B between core-matrix groups), version 1
// (bits 46-47), layout type 2 = SWIZZLE_128B (bits 61-63).
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr) {
  return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}

__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                    __nv_bfloat16* __restrict__ C, int M, int N, int K) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;                                   // [stages][128 × 64 bf16], 1024-byte aligned (swizzle atom)
  uint8_t* sB = smem + kGemmStages * kTileBytes;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + 2 * kGemmStages * kTileBytes);
  uint64_t* empty = full + kGemmStages;
  uint64_t* tmem_full = empty + kGemmStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * kBN, m0 = blockIdx.y * kBM;
  const int kblocks = (K + kBK - 1) / kBK;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    for (int s = 0; s < kGemmStages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(tmem_full, 1);
    fence_mbar_init();
  }
  if (warp == 2) {  // TMEM allocation: one whole warp, the same warp frees it
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 1;
      for (int kb = 0; kb < kblocks; ++kb) {
        mbar_wait(&empty[s], ph);
        mbar_arrive_expect_tx(&full[s], 2 * kTileBytes);
        tma_load_2d(sA + s * kTileBytes, &tmap_a, kb * kBK, m0, &full[s]);
        tma_load_2d(sB + s * kTileBytes, &tmap_b, kb * kBK, n0, &full[s]);
        if (++s == kGemmStages) {
          s = 0;
          ph ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------------------------------- MMA issuer
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int kb = 0; kb < kblocks; ++kb) {
        mbar_wait(&full[s], ph);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(sA + s * kTileBytes), b_addr = smem_u32(sB + s * kTileBytes);
#pragma unroll
        for (int k = 0; k < kBK / 16; ++k) {
          // 16 bf16 = 32 bytes further along K inside the 128-byte swizzle row
          umma_bf16(tmem_base, umma_desc(a_addr + k * 32), umma_desc(b_addr + k * 32), kIdesc,
                    (kb > 0 || k > 0) ? 1u : 0u);
        }
        umma_commit(&empty[s]);  // frees the stage once these MMAs have read it
        if (++s == kGemmStages) {
          s = 0;
          ph ^= 1;
        }
      }
      umma_commit(tmem_full);  // accumulator complete
    }
  } else {
    // ---------------------------------------------------------------------------------------------- epilogue
    mbar_wait(tmem_full, 0);
    tc_fence_after();
    const int q = warp & 3;                 // TMEM lane quarter this warp may access
    const int row = m0 + q * 32 + lane;
#pragma unroll 1
    for (int c = 0; c < kBN; c += 16) {
      uint32_t r[16];
      tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c, r);
      if (row < M) {
        __nv_bfloat16* dst = C + (size_t)row * N + n0 + c;
        if (n0 + c + 16 <= N && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
          uint4 o0, o1;
          o0.x = pack2(__uint_as_float(r[0]), __uint_as_float(r[1]));
          o0.y = pack2(__uint_as_float(r[2]), __uint_as_float(r[3]));
          o0.z = pack2(__uint_as_float(r[4]), __uint_as_float(r[5]));
          o0.w = pack2(__uint_as_float(r[6]), __uint_as_float(r[7]));
          o1.x = pack2(__uint_as_float(r[8]), __uint_as_float(r[9]));
          o1.y = pack2(__uint_as_float(r[10]), __uint_as_float(r[11]));
          o1.z = pack2(__uint_as_float(r[12]), __uint_as_float(r[13]));
          o1.w = pack2(__uint_as_float(r[14]), __uint_as_float(r[15]));
          reinterpret_cast<uint4*>(dst)[0] = o0;
          reinterpret_cast<uint4*>(dst)[1] = o1;
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (n0 + c + j < N) dst[j] = f_to_bf16(__uint_as_float(r[j]));
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------ persistent variant
// The kernel for shapes whose tiles fill the GPU (launch_gemm_bf16 chooses; tests/test_prefill_gpu.py: bit-identical to
// the 128 × 128 kernel on ≥ 99.9 % of the elements, ≤ 1 ulp against the oracle).  Same roles and the same descriptors /
// TMEM access shapes as gemm_tcgen05_kernel, re-arranged for throughput:
//   * 128 × 256 output tiles (one tcgen05.mma is M128 N256 K16): per 64-wide k-block the MMA pipe is busy 512 cycles
//     and reads 48 KB of shared memory (96 B/clk; the 128 × 128 tile sits exactly on the 128 B/clk limit),
//   * persistent CTAs (grid = min(tiles, #SMs)) walking tiles m-fastest, so concurrently running CTAs share weight rows
//     in L2 and barrier init / TMEM allocation are paid once per CTA instead of once per tile,
//   * the fp32 accumulator is double-buffered in TMEM (2 × 256 of the 512 columns): the epilogue of tile i drains
//     buffer i & 1 while the MMA warp already accumulates tile i + 1 (tmem_full / tmem_empty barriers).
constexpr int kPBN = 256;
constexpr int kPStages = 4;
constexpr int kPTileA = kBM * kBK * 2;    // 16 KB
constexpr int kPTileB = kPBN * kBK * 2;   // 32 KB
constexpr int kPTmemCols = 512;
constexpr int kPGemmSmem = kPStages * (kPTileA + kPTileB) + 1024 /*align*/ + 256 /*barriers*/;
constexpr uint32_t kPIdesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kPBN >> 3) << 17) | ((uint32_t)(kBM >> 4) << 24);

// mbarrier wait that traps after ~2 s instead of spinning forever: a protocol bug in this not-yet-run kernel must fail
// the launch, not hang the GPU
__device__ __forceinline__ void mbar_wait_bounded(uint64_t* bar, uint32_t parity) {
  unsigned int spins = 0;
  unsigned long long t0 = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 0xffffu) == 0) {
      const unsigned long long now = global_timer_ns();
      if (t0 == 0) t0 = now;
      if (now - t0 > 2000000000ull) __trap();
    }
  }
}

__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_tcgen05_persistent_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                               __nv_bfloat16* __restrict__ C, int M, int N, int K) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;                                   // [stages][128 × 64 bf16]
  uint8_t* sB = smem + kPStages * kPTileA;              // [stages][256 × 64 bf16]
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + kPStages * (kPTileA + kPTileB));
  uint64_t* empty = full + kPStages;
  uint64_t* tmem_full = empty + kPStages;               // [2]
  uint64_t* tmem_empty = tmem_full + 2;                 // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int mt = (M + kBM - 1) / kBM, nt = (N + kPBN - 1) / kPBN;
  const int tiles = mt * nt;
  const int kblocks = (K + kBK - 1) / kBK;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    for (int s = 0; s < kPStages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full[a], 1);
      mbar_init(&tmem_empty[a], 128);   // every epilogue thread arrives once it holds its part of the tile in registers
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(kPTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 1;
      for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
        const int m0 = (t % mt) * kBM, n0 = (t / mt) * kPBN;
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait_bounded(&empty[s], ph);
          mbar_arrive_expect_tx(&full[s], kPTileA + kPTileB);
          tma_load_2d(sA + s * kPTileA, &tmap_a, kb * kBK, m0, &full[s]);
          tma_load_2d(sB + s * kPTileB, &tmap_b, kb * kBK, n0, &full[s]);
          if (++s == kPStages) {
            s = 0;
            ph ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------------------------------- MMA issuer
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      int i = 0;
      for (int t = blockIdx.x; t < tiles; t += gridDim.x, ++i) {
        const int acc = i & 1;
        mbar_wait_bounded(&tmem_empty[acc], (uint32_t)(((i >> 1) & 1) ^ 1));   // the epilogue drained this buffer's last use
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(acc * kPBN);
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait_bounded(&full[s], ph);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(sA + s * kPTileA), b_addr = smem_u32(sB + s * kPTileB);
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k)
            umma_bf16(tmem_d, umma_desc(a_addr + k * 32), umma_desc(b_addr + k * 32), kPIdesc,
                      (kb > 0 || k > 0) ? 1u : 0u);
          umma_commit(&empty[s]);
          if (++s == kPStages) {
            s = 0;
            ph ^= 1;
          }
        }
        umma_commit(&tmem_full[acc]);
      }
    }
  } else {
    // ---------------------------------------------------------------------------------------------- epilogue
    const int q = warp & 3;
    int i = 0;
    for (int t = blockIdx.x; t < tiles; t += gridDim.x, ++i) {
      const int acc = i & 1;
      const int m0 = (t % mt) * kBM, n0 = (t / mt) * kPBN;
      mbar_wait_bounded(&tmem_full[acc], (uint32_t)((i >> 1) & 1));
      tc_fence_after();
      const int row = m0 + q * 32 + lane;
#pragma unroll 1
      for (int c = 0; c < kPBN; c += 16) {
        uint32_t r[16];
        tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * kPBN + c), r);
        if (row < M && n0 + c < N) {
          __nv_bfloat16* dst = C + (size_t)row * N + n0 + c;
          if (n0 + c + 16 <= N && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
            uint4 o0, o1;
            o0.x = pack2(__uint_as_float(r[0]), __uint_as_float(r[1]));
            o0.y = pack2(__uint_as_float(r[2]), __uint_as_float(r[3]));
            o0.z = pack2(__uint_as_float(r[4]), __uint_as_float(r[5]));
            o0.w = pack2(__uint_as_float(r[6]), __uint_as_float(r[7]));
            o1.x = pack2(__uint_as_float(r[8]), __uint_as_float(r[9]));
            o1.y = pack2(__uint_as_float(r[10]), __uint_as_float(r[11]));
            o1.z = pack2(__uint_as_float(r[12]), __uint_as_float(r[13]));
            o1.w = pack2(__uint_as_float(r[14]), __uint_as_float(r[15]));
            reinterpret_cast<uint4*>(dst)[0] = o0;
            reinterpret_cast<uint4*>(dst)[1] = o1;
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (n0 + c + j < N) dst[j] = f_to_bf16(__uint_as_float(r[j]));
          }
        }
      }
      // all of this thread's tcgen05.ld have completed (tmem_ld16 waits): hand the buffer back to the MMA warp
      tc_fence_before();
      mbar_arrive(&tmem_empty[acc]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kPTmemCols) : "memory");
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int make_tmap_sw128(CUtensorMap* out, const void* base, int64_t rows, int64_t cols, int box_rows = kBM) {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
      q != cudaDriverEntryPointSuccess) {
    (void)cudaGetLastError();
    set_error("cuTensorMapEncodeTiled is not available (no CUDA driver / device?)");
    return B200_ERR_NO_DEVICE;
  }
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
  cuuint32_t box[2] = {(cuuint32_t)kBK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = reinterpret_cast<EncodeTiledFn>(p)(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims,
                                                  strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(swizzle 128B) failed with CUresult %d (rows=%lld cols=%lld)", (int)r,
              (long long)rows, (long long)cols);
    return B200_ERR_CUDA;
  }
  return B200_OK;
}

}  // namespace

int launch_gemm_bf16(void* C, const void* A, const void* B, int64_t M, int64_t N, int64_t K, cudaStream_t st) {
  B200_CHECK_ARG(C && A && B && M > 0 && N > 0 && K > 0, "gemm: bad arguments");
  B200_CHECK_ARG(K % 8 == 0 && N % 8 == 0, "gemm: K=%lld and N=%lld must be multiples of 8 (16-byte rows)", (long long)K,
                 (long long)N);
  B200_CHECK_ARG(((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(B) | reinterpret_cast<uintptr_t>(C)) & 15) == 0,
                 "gemm: operands must be 16-byte aligned");
  B200_CHECK_ARG(M < (1ll << 31) && N < (1ll << 31) && K < (1ll << 31), "gemm: shape out of range");
  static bool attr_set = false;
  if (!attr_set) {
    B200_CUDA(cudaFuncSetAttribute(gemm_tcgen05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kGemmSmem));
    attr_set = true;
  }
  CUtensorMap ta, tb;
  int rc;
  // Which kernel: the persistent 128×256 one when its tiles fill the GPU (measured on B200, tools/gemm_bench.py: 0.55–0.74
  // of the cuBLAS bf16 peak at M = 2048 against 0.39–0.49 for one 128×128 tile per CTA; with fewer than ~100 tiles the
  // small-tile kernel wins).  B200_GEMM = "persistent" | "tile" forces either.
  const int64_t ptiles = ((M + kBM - 1) / kBM) * ((N + kPBN - 1) / kPBN);
  if (env_choice("B200_GEMM", 'p', ptiles >= 100)) {
    static std::once_flag p_once;
    static cudaError_t p_err = cudaSuccess;
    static int sms = 0;
    std::call_once(p_once, [] {
      p_err = cudaFuncSetAttribute(gemm_tcgen05_persistent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kPGemmSmem);
      int dev = 0;
      if (p_err == cudaSuccess) p_err = cudaGetDevice(&dev);
      if (p_err == cudaSuccess) p_err = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    });
    B200_CUDA(p_err);
    if ((rc = make_tmap_sw128(&ta, A, M, K, kBM)) != B200_OK) return rc;
    if ((rc = make_tmap_sw128(&tb, B, N, K, kPBN)) != B200_OK) return rc;
    const int64_t tiles = ((M + kBM - 1) / kBM) * ((N + kPBN - 1) / kPBN);
    g_launches.fetch_add(1);
    gemm_tcgen05_persistent_kernel<<<(unsigned)std::min<int64_t>(tiles, sms), kGemmThreads, kPGemmSmem, st>>>(
        ta, tb, (__nv_bfloat16*)C, (int)M, (int)N, (int)K);
    B200_CUDA(cudaGetLastError());
    return B200_OK;
  }
  if ((rc = make_tmap_sw128(&ta, A, M, K)) != B200_OK) return rc;
  if ((rc = make_tmap_sw128(&tb, B, N, K)) != B200_OK) return rc;
  dim3 grid((unsigned)((N + kBN - 1) / kBN), (unsigned)((M + kBM - 1) / kBM));
  g_launches.fetch_add(1);
  gemm_tcgen05_kernel<<<grid, kGemmThreads, kGemmSmem, st>>>(ta, tb, (__nv_bfloat16*)C, (int)M, (int)N, (int)K);
  B200_CUDA(cudaGetLastError());
  return B200_OK;
}

}  // namespace b200

extern "C" int b200_gemm_bf16(void* C, const void* A, const void* B, int64_t M, int64_t N, int64_t K, void* stream) {
  int rc = b200_device_check();
  if (rc != B200_OK) return rc;
  return b200::launch_gemm_bf16(C, A, B, M, N, K, (cudaStream_t)stream);
}
