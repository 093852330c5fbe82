// common.cu — error state, launch counter, device check and TMA descriptor encoding for the C ABI.
#include "common.cuh"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <mutex>

namespace b200 {

static thread_local char t_err[512] = "";
std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(t_err, sizeof(t_err), fmt, ap);
  va_end(ap);
}

const char* last_error() { return t_err; }

bool env_flag(const char* name, bool dflt) {
  const char* e = std::getenv(name);
  return (e == nullptr || e[0] == 0) ? dflt : e[0] == '1';
}
bool env_choice(const char* name, char yes_initial, bool dflt) {
  const char* e = std::getenv(name);
  return (e == nullptr || e[0] == 0) ? dflt : e[0] == yes_initial;
}
int env_int(const char* name, int dflt, int lo, int hi) {
  const char* e = std::getenv(name);
  if (e == nullptr || e[0] == 0) return dflt;
  char* end = nullptr;
  const long v = std::strtol(e, &end, 10);
  if (end == e) return dflt;
  return (int)std::max<long>(lo, std::min<long>(hi, v));
}

// cuTensorMapEncodeTiled is a driver entry point; resolve it through the runtime so the library does not link libcuda
// (the build machine has no driver).
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
    else
      (void)cudaGetLastError();
  });
  return fn;
}

int make_tmap_2d_bf16(CUtensorMap* out, const void* base, int64_t rows, int64_t cols, int box_rows, int box_cols) {
  EncodeTiledFn enc = get_encode();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled is not available (no CUDA driver / device?)");
    return B200_ERR_NO_DEVICE;
  }
  // box_cols > 256 bf16: describe the rows in 4- or 8-byte elements (inner box dimension ≤ 256 elements); the column
  // coordinate the kernel passes is then in those elements (box index × 256 either way)
  int epb = 1;  // bf16 per element
  while (box_cols / epb > 256) epb *= 2;
  if (epb > 4 || cols % epb != 0 || box_cols % epb != 0) {
    set_error("tensor map: box of %d columns not expressible (cols=%lld)", box_cols, (long long)cols);
    return B200_ERR_INVALID;
  }
  const CUtensorMapDataType dt = epb == 1 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
                                 : epb == 2 ? CU_TENSOR_MAP_DATA_TYPE_UINT32 : CU_TENSOR_MAP_DATA_TYPE_UINT64;
  cuuint64_t dims[2] = {(cuuint64_t)(cols / epb), (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
  cuuint32_t box[2] = {(cuuint32_t)(box_cols / epb), (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, dt, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (rows=%lld cols=%lld box=%dx%d base=%p)", (int)r,
              (long long)rows, (long long)cols, box_rows, box_cols, base);
    return B200_ERR_CUDA;
  }
  return B200_OK;
}

const char* last_error();

}  // namespace b200

extern "C" {

int b200_abi_version(void) { return B200_ABI_VERSION; }
const char* b200_last_error(void) { return b200::last_error(); }
int64_t b200_launch_count(void) { return b200::g_launches.load(); }

int b200_device_check(void) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) {
    (void)cudaGetLastError();
    b200::set_error("no CUDA device: %s", cudaGetErrorString(e));
    return B200_ERR_NO_DEVICE;
  }
  int major = 0;
  e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  if (e != cudaSuccess || major != 10) {
    (void)cudaGetLastError();
    b200::set_error("device %d has compute capability major %d; this library contains sm_100a code only", dev, major);
    return B200_ERR_NO_DEVICE;
  }
  return B200_OK;
}

}  // extern "C"
