// ll_bench.cu — cost of ONE grid-wide all-to-all dependency inside a persistent kernel when the activation vector
// itself carries the flag ("LL" words {tag:16 | bf16:16}, relaxed stores/loads, no fence, no counter), next to the
// counter barrier of sync_bench.cu (variant D/F: red.release.gpu + ld.acquire.gpu polling = 1.25–2 µs per step).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ll_bench ll_bench.cu && ./ll_bench
//
// Per step every CTA polls the whole vector (kVec words) until all tags are this step's, block-reduces it (the RMSNorm
// sum of squares of the real prologue), then writes ITS slice of the next vector into the other of two buffers (the
// writer of step s+1 can only run after it has seen all of step s, i.e. after every CTA has finished reading step
// s-1: two buffers are enough).  Variants: vector length 896 / 4864, reduction on / off, 1 or 2 dependent exchanges.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x)                                                                    \
  do {                                                                           \
    cudaError_t e_ = (x);                                                        \
    if (e_ != cudaSuccess) {                                                     \
      printf("%s failed: %s (line %d)\n", #x, cudaGetErrorString(e_), __LINE__); \
      exit(1);                                                                   \
    }                                                                            \
  } while (0)

constexpr int kThreads = 288;
constexpr int kPoll = 256;

__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ uint4 ld_vol4(const uint4* p) {
  uint4 v;
  asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_vol(unsigned int* p, unsigned int v) {
  asm volatile("st.volatile.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__device__ __forceinline__ uint4 ld_rlx4(const uint4* p) {
  uint4 v;
  asm volatile("ld.relaxed.gpu.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_rlx(unsigned int* p, unsigned int v) {
  asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// mode bit 0: block reduction after the poll; bit 1: relaxed.gpu instead of volatile.
// `replicas` copies of the vector (stride `rstride` words): CTA c polls copy c % replicas, writers store every copy — the
// push model: with replicas == gridDim.x every L2 line has exactly one polling CTA.
__global__ void __launch_bounds__(kThreads, 1)
ll_kernel(unsigned int* buf0, unsigned int* buf1, int nvec, int steps, int mode, unsigned long long* stamps, int replicas,
          int rstride) {
  extern __shared__ float dyn[];
  float* red = dyn;
  const int per = (nvec + gridDim.x - 1) / gridDim.x;
  const int lo = blockIdx.x * per, hi = min(nvec, lo + per);
  const size_t roff = (size_t)(blockIdx.x % replicas) * rstride;
  for (int s = 0; s < steps; ++s) {
    const unsigned int tag = (unsigned int)(s % 65535) + 1u;
    const unsigned int* src = ((s & 1) ? buf1 : buf0) + roff;
    unsigned int* dst = (s & 1) ? buf0 : buf1;
    float acc = 0.f;
    if (threadIdx.x < kPoll) {
      for (int i = threadIdx.x; i < nvec / 4; i += kPoll) {
        uint4 v;
        unsigned int spins = 0;
        unsigned long long t0 = 0;
        for (;;) {
          v = (mode & 2) ? ld_rlx4(reinterpret_cast<const uint4*>(src) + i) : ld_vol4(reinterpret_cast<const uint4*>(src) + i);
          if ((v.x >> 16) == tag && (v.y >> 16) == tag && (v.z >> 16) == tag && (v.w >> 16) == tag) break;
          if ((++spins & 0xfffu) == 0) {
            const unsigned long long now = gtime();
            if (t0 == 0) t0 = now;
            if (now - t0 > 2000000000ull) __trap();
          }
        }
        acc += __uint_as_float(v.x << 16) + __uint_as_float(v.y << 16) + __uint_as_float(v.z << 16) + __uint_as_float(v.w << 16);
      }
    }
    float tot = acc;
    if (mode & 1) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
      __syncthreads();
      tot = 0.f;
      for (int w = 0; w < kThreads / 32; ++w) tot += red[w];
      __syncthreads();
    } else {
      __syncthreads();
    }
    const unsigned int ntag = (unsigned int)((s + 1) % 65535) + 1u;
    const int nown = hi - lo;
    for (int j = threadIdx.x; j < nown * replicas; j += kThreads) {
      const int r = j / nown, i = lo + j % nown;
      const unsigned int val = (__float_as_uint(tot * 1e-3f + 1.f) >> 16) & 0xffffu;
      if (mode & 2) st_rlx(dst + (size_t)r * rstride + i, (ntag << 16) | val);
      else st_vol(dst + (size_t)r * rstride + i, (ntag << 16) | val);
    }
    if (stamps != nullptr && blockIdx.x == 0 && threadIdx.x == 0 && s < 256) stamps[s] = gtime();
  }
}

int main() {
  int dev = 0;
  CK(cudaSetDevice(dev));
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  CK(cudaFuncSetAttribute(ll_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
  const int steps = 2000;
  unsigned int *b0, *b1;
  unsigned long long* stamps;
  const size_t kBufWords = (size_t)148 * 16384;
  CK(cudaMalloc(&b0, kBufWords * 4));
  CK(cudaMalloc(&b1, kBufWords * 4));
  CK(cudaMalloc(&stamps, 256 * 8));
  const int vecs[3] = {896, 4864, 14336};
  const int reps[5] = {1, 4, 16, 37, 148};
  for (int vi = 0; vi < 3; ++vi) {
   for (int ri = 0; ri < 5; ++ri) {
    for (int mode = 0; mode < 4; mode += (ri == 0 ? 1 : 2)) {
      const int nvec = vecs[vi];
      int replicas = reps[ri] > sms ? sms : reps[ri];
      int rstride = (nvec + 31) / 32 * 32;
      std::vector<unsigned int> init((size_t)replicas * rstride, (1u << 16) | 0x3f80u);  // tag 1 (step 0), value 1.0
      CK(cudaMemset(b0, 0, kBufWords * 4));
      CK(cudaMemset(b1, 0, kBufWords * 4));
      CK(cudaMemcpy(b0, init.data(), init.size() * 4, cudaMemcpyHostToDevice));
      cudaEvent_t e0, e1;
      CK(cudaEventCreate(&e0));
      CK(cudaEventCreate(&e1));
      void* args[] = {&b0, &b1, (void*)&nvec, (void*)&steps, &mode, &stamps, &replicas, &rstride};
      CK(cudaEventRecord(e0));
      CK(cudaLaunchCooperativeKernel((void*)ll_kernel, dim3(sms), dim3(kThreads), args, 100 * 1024, 0));
      CK(cudaEventRecord(e1));
      CK(cudaDeviceSynchronize());
      float ms = 0.f;
      CK(cudaEventElapsedTime(&ms, e0, e1));
      std::vector<unsigned long long> st(256);
      CK(cudaMemcpy(st.data(), stamps, 256 * 8, cudaMemcpyDeviceToHost));
      printf("LL exchange, %5d-word vector, %3d replicas, %s, %s : %6.3f us/step   (in-kernel stamps, steps 64..255: %6.3f us/step)\n",
             nvec, replicas, (mode & 2) ? "relaxed.gpu" : "volatile   ", (mode & 1) ? "poll + block reduce" : "poll only          ",
             ms * 1e3 / steps, (double)(st[255] - st[64]) / 191.0 * 1e-3);
    }
   }
  }
  return 0;
}
