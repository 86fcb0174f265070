// Peer-visible device buffers for the multi-GPU output exchange (one process per GPU on one NVSwitch box).
// The reference has no distributed code; on N > 1 ranks the only inference-path exchange is the gather of
// the core-node outputs.  Instead of a separate all-gather after the head kernel, every rank's head
// epilogue stores its rows straight into every peer's gather buffer over NVLink (fitgnn_gemm_head_rows_peers); these
// helpers provide the buffers: plain cudaMalloc allocations exported / imported as CUDA IPC handles (the 64-byte
// handles travel through torch.distributed).  Opening a handle enables peer access to the exporting device.
#include <string.h>
#include "common.cuh"

using namespace fitgnn;

static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handles are exchanged as 64-byte blobs");

extern "C" int fitgnn_peer_alloc(size_t bytes, void** dev_ptr, uint8_t* handle_out) {
  FG_REQUIRE(dev_ptr && handle_out && bytes > 0, FITGNN_EINVAL, "peer_alloc: bad arguments");
  void* p = nullptr;
  FG_CUDA(cudaMalloc(&p, bytes));
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) {
    cudaFree(p);
    set_error("peer_alloc: cudaIpcGetMemHandle -> %s", cudaGetErrorString(e));
    return FITGNN_ECUDA;
  }
  FG_CUDA(cudaMemset(p, 0, bytes));
  memcpy(handle_out, &h, sizeof(h));
  *dev_ptr = p;
  return FITGNN_OK;
}

extern "C" int fitgnn_peer_open(const uint8_t* handle, void** dev_ptr) {
  FG_REQUIRE(handle && dev_ptr, FITGNN_EINVAL, "peer_open: bad arguments");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof(h));
  void* p = nullptr;
  FG_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  *dev_ptr = p;
  return FITGNN_OK;
}

extern "C" int fitgnn_peer_close(void* dev_ptr) {
  if (dev_ptr) FG_CUDA(cudaIpcCloseMemHandle(dev_ptr));
  return FITGNN_OK;
}

extern "C" int fitgnn_peer_free(void* dev_ptr) {
  if (dev_ptr) FG_CUDA(cudaFree(dev_ptr));
  return FITGNN_OK;
}
