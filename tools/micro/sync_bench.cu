// sync_bench.cu — what does ONE grid-wide dependency cost on a B200, by mechanism?  (round-2 decision input: the
// Qwen2.5-0.5B token is 123 dependent launches at ≈3.3 µs each against 1.2 µs of HBM time per launch.)
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o sync_bench sync_bench.cu && ./sync_bench
//
// Every variant runs the SAME minimal body per step on 148 CTAs × 288 threads with 100 KB of dynamic shared memory
// (the footprint of the real GEMV): read a 1 792-byte vector produced by the previous step (every CTA reads all of
// it, like the activation vector), reduce it in the CTA, write this CTA's 12 bytes of it back.  Reported: µs per step.
//   A  graph of kernels, full dependencies                       (no PDL)
//   B  graph of kernels, programmatic dependent launch, griddepcontrol.wait            (what the engine does today)
//   C  graph of kernels, PDL launch + completion-counter polling instead of the wait   (B200_FLAGSYNC=1)
//   D  one persistent kernel, counter grid barrier (red.release.gpu + ld.acquire.gpu polling by one thread per CTA)
//   E  as D, the vector replicated 8× so that CTAs read different L2 lines (hot-line contention test)
//   F  as D without the body's global traffic (pure barrier cost)
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x)                                                                      \
  do {                                                                             \
    cudaError_t e_ = (x);                                                          \
    if (e_ != cudaSuccess) {                                                       \
      printf("%s failed: %s (line %d)\n", #x, cudaGetErrorString(e_), __LINE__);   \
      exit(1);                                                                     \
    }                                                                              \
  } while (0)

constexpr int kThreads = 288;
constexpr int kVec = 896;          // bf16-sized elements → 1 792 bytes, stored as 448 floats here
constexpr int kWords = kVec / 2;   // floats
constexpr int kReplicas = 8;

__device__ __forceinline__ unsigned long long ld_acquire(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release(unsigned long long* p) {
  asm volatile("red.release.gpu.global.add.u64 [%0], %1;" ::"l"(p), "l"(1ull) : "memory");
}
__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// the body: every CTA reads the whole vector (through L2), block-reduces, rewrites its own slice
__device__ __forceinline__ void body(float* vec, int replicas, float* red) {
  const float* src = vec + (size_t)(blockIdx.x % replicas) * kWords;
  float s = 0.f;
  for (int i = threadIdx.x; i < kWords; i += kThreads) s += __ldcg(src + i);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 3) {
    float t = 0.f;
    for (int w = 0; w < kThreads / 32; ++w) t += red[w];
    const int i = blockIdx.x * 3 + threadIdx.x;
    if (i < kWords)
      for (int r = 0; r < replicas; ++r) vec[(size_t)r * kWords + i] = t * 1e-6f + 0.5f;
  }
  __syncthreads();
}

// mode 0: plain / PDL with griddepcontrol.wait; mode 1: PDL launch + counter polling
__global__ void __launch_bounds__(kThreads, 1)
step_kernel(float* vec, unsigned long long* ctr, const unsigned long long* epoch, int step, int nsteps, int mode) {
  extern __shared__ float dyn[];
  float* red = dyn;
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (mode == 0) {
    asm volatile("griddepcontrol.wait;" ::: "memory");
  } else {
    if (step > 0 && step < nsteps - 1) {  // first/last node: full dependency, nothing to poll (and the last one
                                          // advances the epoch, which none of its own CTAs may still have to read)
      if (threadIdx.x == 0) {
        const unsigned long long target = (*epoch + 1ull) * gridDim.x;
        unsigned int spins = 0;
        unsigned long long t0 = 0;
        while (ld_acquire(ctr + 16 * (step - 1)) < target) {
          if ((++spins & 0xfffu) == 0) {
            const unsigned long long now = gtime();
            if (t0 == 0) t0 = now;
            if (now - t0 > 2000000000ull) __trap();
          }
        }
      }
      __syncthreads();
    } else {
      asm volatile("griddepcontrol.wait;" ::: "memory");
    }
  }
  body(vec, 1, red);
  if (mode == 1) {
    if (threadIdx.x == 0) red_release(ctr + 16 * step);
    // the last node (full dependency) advances the epoch: every CTA of this "token" has read it
    if (step == nsteps - 1 && blockIdx.x == 0 && threadIdx.x == 0) *(unsigned long long*)epoch += 1ull;
  }
}

__global__ void __launch_bounds__(kThreads, 1)
persistent_kernel(float* vec, unsigned long long* bar, int steps, int replicas, int do_body, unsigned long long* stamps) {
  extern __shared__ float dyn[];
  float* red = dyn;
  unsigned long long target = 0;
  for (int p = 0; p < steps; ++p) {
    if (do_body) body(vec, replicas, red);
    else __syncthreads();
    if (threadIdx.x == 0) {
      target += gridDim.x;
      red_release(bar);
      unsigned int spins = 0;
      unsigned long long t0 = 0;
      while (ld_acquire(bar) < target) {
        if ((++spins & 0xfffu) == 0) {  // never hang the GPU: a CTA that is not co-resident would stall everyone
          const unsigned long long now = gtime();
          if (t0 == 0) t0 = now;
          if (now - t0 > 2000000000ull) __trap();
        }
      }
    }
    __syncthreads();
    if (stamps != nullptr && blockIdx.x == 0 && threadIdx.x == 0 && p < 64) stamps[p] = gtime();
  }
}

static float time_graph(cudaStream_t st, float* vec, unsigned long long* ctr, unsigned long long* epoch, int nodes, bool pdl,
                        int mode) {
  cudaGraph_t g;
  cudaGraphExec_t ge;
  CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
  for (int i = 0; i < nodes; ++i) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(148);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = 100 * 1024;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    // first and last node of a graph: full dependency
    at[0].val.programmaticStreamSerializationAllowed = (pdl && i > 0 && i < nodes - 1) ? 1 : 0;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    CK(cudaLaunchKernelEx(&cfg, step_kernel, vec, ctr, (const unsigned long long*)epoch, i, nodes, mode));
  }
  CK(cudaStreamEndCapture(st, &g));
  CK(cudaGraphInstantiate(&ge, g, 0));
  for (int w = 0; w < 5; ++w) CK(cudaGraphLaunch(ge, st));
  CK(cudaStreamSynchronize(st));
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a));
  CK(cudaEventCreate(&b));
  const int reps = 50;
  CK(cudaEventRecord(a, st));
  for (int r = 0; r < reps; ++r) CK(cudaGraphLaunch(ge, st));
  CK(cudaEventRecord(b, st));
  CK(cudaStreamSynchronize(st));
  float ms = 0.f;
  CK(cudaEventElapsedTime(&ms, a, b));
  CK(cudaGraphExecDestroy(ge));
  CK(cudaGraphDestroy(g));
  return ms * 1000.f / (reps * nodes);
}

int main() {
  const int nodes = 123;
  float* vec;
  unsigned long long *ctr, *epoch, *bar, *stamps;
  CK(cudaMalloc(&vec, kReplicas * kWords * sizeof(float)));
  CK(cudaMemset(vec, 0, kReplicas * kWords * sizeof(float)));
  CK(cudaMalloc(&ctr, (size_t)nodes * 128));
  CK(cudaMemset(ctr, 0, (size_t)nodes * 128));
  CK(cudaMalloc(&epoch, 128));
  CK(cudaMemset(epoch, 0, 128));
  CK(cudaMalloc(&bar, 128));
  CK(cudaMalloc(&stamps, 64 * 8));
  cudaStream_t st;
  CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  CK(cudaFuncSetAttribute(step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
  CK(cudaFuncSetAttribute(step_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  CK(cudaFuncSetAttribute(persistent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));

  printf("A graph, full dependencies            : %6.2f us/step\n", time_graph(st, vec, ctr, epoch, nodes, false, 0));
  printf("B graph, PDL + griddepcontrol.wait    : %6.2f us/step\n", time_graph(st, vec, ctr, epoch, nodes, true, 0));
  printf("C graph, PDL launch + counter polling : %6.2f us/step\n", time_graph(st, vec, ctr, epoch, nodes, true, 1));

  const int steps = 2000;
  struct V {
    const char* name;
    int replicas, do_body;
  } vs[3] = {{"D persistent, counter barrier, body    ", 1, 1},
             {"E persistent, barrier, 8 replicas      ", kReplicas, 1},
             {"F persistent, barrier only             ", 1, 0}};
  for (const V& v : vs) {
    for (int w = 0; w < 2; ++w) {
      CK(cudaMemsetAsync(bar, 0, 128, st));
      persistent_kernel<<<148, kThreads, 100 * 1024, st>>>(vec, bar, 200, v.replicas, v.do_body, nullptr);
    }
    CK(cudaMemsetAsync(bar, 0, 128, st));
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a));
    CK(cudaEventCreate(&b));
    CK(cudaEventRecord(a, st));
    persistent_kernel<<<148, kThreads, 100 * 1024, st>>>(vec, bar, steps, v.replicas, v.do_body, stamps);
    CK(cudaEventRecord(b, st));
    CK(cudaStreamSynchronize(st));
    CK(cudaGetLastError());
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, a, b));
    unsigned long long h[64];
    CK(cudaMemcpy(h, stamps, sizeof(h), cudaMemcpyDeviceToHost));
    printf("%s: %6.2f us/step   (in-kernel stamps, steps 32..63: %.2f us/step)\n", v.name, ms * 1000.f / steps,
           (double)(h[63] - h[32]) / 31.0 / 1e3);
  }
  return 0;
}
