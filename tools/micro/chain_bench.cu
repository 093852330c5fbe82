// Microbenchmark: cost per kernel of a dependent chain inside a CUDA graph, with/without PDL, and of a flag-based
// grid barrier inside one persistent kernel.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o chain_bench chain_bench.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <vector>

__global__ void __launch_bounds__(288) chain_kernel(float* buf, int n, int pdl) {
  extern __shared__ float sm[];
  if (pdl) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (pdl) asm volatile("griddepcontrol.wait;" ::: "memory");
  // read a small vector produced by the previous kernel, reduce, write own slice
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s += buf[i];
  sm[threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < blockDim.x; ++i) t += sm[i];
    buf[n + blockIdx.x] = t * 1e-9f;
    if (blockIdx.x < n) buf[blockIdx.x] = buf[blockIdx.x] * 0.999f + 1e-6f;
  }
}

// persistent kernel with a flag-based grid barrier between phases
__global__ void __launch_bounds__(288) persistent_kernel(float* buf, int n, unsigned int* bar, int phases) {
  extern __shared__ float sm[];
  unsigned int target = 0;
  for (int p = 0; p < phases; ++p) {
    float s = 0.f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) s += __ldcg(buf + i);
    sm[threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
      float t = 0.f;
      for (int i = 0; i < blockDim.x; ++i) t += sm[i];
      buf[n + blockIdx.x] = t * 1e-9f;
      if (blockIdx.x < n) buf[blockIdx.x] = buf[blockIdx.x] * 0.999f + 1e-6f;
      __threadfence();
      target += gridDim.x;
      atomicAdd(bar, 1u);
      while (*((volatile unsigned int*)bar) < target) {
      }
      __threadfence();
    }
    __syncthreads();
  }
}

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

int main() {
  float* buf;
  unsigned int* bar;
  const int n = 896, N = 120;
  CK(cudaMalloc(&buf, 1 << 20));
  CK(cudaMemset(buf, 0, 1 << 20));
  CK(cudaMalloc(&bar, 4));
  cudaStream_t st;
  CK(cudaStreamCreate(&st));
  for (int smem_kb : {2, 100}) {
    CK(cudaFuncSetAttribute(chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024));
    CK(cudaFuncSetAttribute(chain_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    for (int pdl = 0; pdl <= 1; ++pdl) {
      for (int grid : {1, 148}) {
        cudaGraph_t g;
        cudaGraphExec_t ge;
        CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
        for (int i = 0; i < N; ++i) {
          cudaLaunchConfig_t cfg{};
          cfg.gridDim = dim3(grid);
          cfg.blockDim = dim3(288);
          cfg.dynamicSmemBytes = smem_kb * 1024;
          cfg.stream = st;
          cudaLaunchAttribute at[1];
          at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
          at[0].val.programmaticStreamSerializationAllowed = (pdl && i > 0 && i < N - 1) ? 1 : 0;
          cfg.attrs = at;
          cfg.numAttrs = 1;
          CK(cudaLaunchKernelEx(&cfg, chain_kernel, buf, n, pdl));
        }
        CK(cudaStreamEndCapture(st, &g));
        CK(cudaGraphInstantiate(&ge, g, 0));
        for (int w = 0; w < 3; ++w) CK(cudaGraphLaunch(ge, st));
        cudaEvent_t a, b;
        cudaEventCreate(&a);
        cudaEventCreate(&b);
        CK(cudaEventRecord(a, st));
        const int reps = 20;
        for (int r = 0; r < reps; ++r) CK(cudaGraphLaunch(ge, st));
        CK(cudaEventRecord(b, st));
        CK(cudaStreamSynchronize(st));
        float ms;
        cudaEventElapsedTime(&ms, a, b);
        printf("graph chain: smem %3d KB  pdl %d  grid %3d : %.2f us per kernel\n", smem_kb, pdl, grid,
               ms * 1000 / (reps * N));
        cudaGraphExecDestroy(ge);
        cudaGraphDestroy(g);
      }
    }
  }
  CK(cudaFuncSetAttribute(persistent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024));
  for (int grid : {148}) {
    for (int w = 0; w < 2; ++w) {
      CK(cudaMemsetAsync(bar, 0, 4, st));
      persistent_kernel<<<grid, 288, 100 * 1024, st>>>(buf, n, bar, N);
    }
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    CK(cudaMemsetAsync(bar, 0, 4, st));
    CK(cudaEventRecord(a, st));
    persistent_kernel<<<grid, 288, 100 * 1024, st>>>(buf, n, bar, N * 10);
    CK(cudaEventRecord(b, st));
    CK(cudaStreamSynchronize(st));
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    printf("persistent kernel, flag grid barrier, grid %d: %.2f us per phase\n", grid, ms * 1000 / (N * 10));
  }
  return 0;
}
