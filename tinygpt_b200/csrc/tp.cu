// tp.cu — tensor-parallel exchange windows (one process per GPU; peers map each other's window through CUDA IPC and
// the GEMV epilogue/prologue exchange fp32 partial hidden vectors over NVLink, see gemv.cu EPI_TP_PUSH /
// PRO_TP_RMSNORM).  [ref: the collective this replaces is ncclAllReduce at
// third_party/TinyTorch/src/Distributed/BackendNCCL.cpp:25-35, which TinyGPT itself never calls (README.md:32).]
#include "common.cuh"

#include <cstring>

extern "C" {

int64_t b200_tp_window_bytes(const b200_model_desc* d) {
  if (!d || d->tp_world < 1 || d->tp_world > 8 || d->hidden <= 0 || d->layers <= 0) return -1;
  // layout (engine.cu tp_slot_off / tp_flag_off / tp_cand_off): per exchange point (2 per layer, +1 for the argmax
  // candidates, +1 spare) `world` partial hidden vectors of H floats (256-byte aligned); then one 256-byte line of
  // arrival counters per point; then `world` 16-byte (max logit, global index) candidates.
  const int64_t points = 2ll * d->layers + 2;
  const int64_t vec = ((int64_t)d->hidden * 8 + 255) / 256 * 256;  // {value, tag} 8-byte words
  return points * d->tp_world * vec + points * 256 + 8 * 16 + 256;
}

int b200_tp_window_create(int64_t bytes, void** window_out, uint8_t handle_out[B200_IPC_HANDLE_BYTES]) {
  using namespace b200;
  B200_CHECK_ARG(bytes > 0 && window_out && handle_out, "tp_window_create: bad argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == B200_IPC_HANDLE_BYTES, "IPC handle size");
  void* p = nullptr;
  B200_CUDA(cudaMalloc(&p, (size_t)bytes));
  B200_CUDA(cudaMemset(p, 0, (size_t)bytes));
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) {
    cudaFree(p);
    set_error("cudaIpcGetMemHandle failed: %s", cudaGetErrorString(e));
    return B200_ERR_CUDA;
  }
  std::memcpy(handle_out, &h, sizeof(h));
  *window_out = p;
  return B200_OK;
}

int b200_tp_window_open(const uint8_t handle[B200_IPC_HANDLE_BYTES], void** peer_window_out) {
  using namespace b200;
  B200_CHECK_ARG(handle && peer_window_out, "tp_window_open: bad argument");
  cudaIpcMemHandle_t h;
  std::memcpy(&h, handle, sizeof(h));
  B200_CUDA(cudaIpcOpenMemHandle(peer_window_out, h, cudaIpcMemLazyEnablePeerAccess));
  return B200_OK;
}

int b200_tp_window_close(void* peer_window) {
  using namespace b200;
  if (peer_window) B200_CUDA(cudaIpcCloseMemHandle(peer_window));
  return B200_OK;
}

int b200_tp_window_destroy(void* window) {
  using namespace b200;
  if (window) B200_CUDA(cudaFree(window));
  return B200_OK;
}

}  // extern "C"
