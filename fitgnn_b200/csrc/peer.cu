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

// ---------------------------------------------------------------------------------------------------------
// fitgnn_peer_push: this rank's slot of the gather buffer -> the same slot of every peer's buffer, as ONE small kernel
// that runs on a side stream behind the NEXT step's compute.  Round 1 measured the two alternatives on 8 x B200:
// the head kernel storing to all 7 peers itself is NVLink-egress-bound (0.64 ms inside the critical path), per-peer
// cudaMemcpyAsync pushes reach only ~310 GB/s aggregate (copy engines), so 0.41 ms of the exchange stayed exposed.  Here a
// few CTAs (default 16 of 148 SMs) stream the slot through shared memory with bulk copies: one cp.async.bulk global->shared
// per chunk, then one cp.async.bulk shared->global per peer (full-line writes over NVLink, the access pattern that reached
// 635 GB/s from the head kernel), double-buffered so loads and stores overlap.
// ---------------------------------------------------------------------------------------------------------
namespace {

constexpr int PUSH_CHUNK = 32 * 1024;  // bytes per bulk copy
constexpr int PUSH_STAGES = 7;   // shared-memory stages (224 KB: the CTA owns its SM)
constexpr int PUSH_DIST = 4;     // loads issued ahead of the stores (128 KB in flight per CTA)
constexpr int PUSH_PENDING = PUSH_STAGES - PUSH_DIST - 1;  // store groups that may still be reading shared memory
constexpr int PUSH_MAX_PEERS = 8;

struct PushArgs {
  char* dst[PUSH_MAX_PEERS];
  int n_dst;
};

__device__ __forceinline__ uint32_t push_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(32)
peer_push_kernel(const char* __restrict__ src, PushArgs pa, size_t bytes) {
  extern __shared__ __align__(128) unsigned char push_smem[];
  __shared__ __align__(8) uint64_t full[PUSH_STAGES];
  const size_t n_chunks = (bytes + PUSH_CHUNK - 1) / PUSH_CHUNK;
  if (threadIdx.x == 0) {
    for (int s = 0; s < PUSH_STAGES; ++s)
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(push_smem_u32(&full[s])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  if (threadIdx.x != 0) return;  // one thread drives the TMA engine
  uint32_t phase[PUSH_STAGES] = {};
  // chunk i of this CTA is global chunk blockIdx.x + i * gridDim.x
  auto load = [&](size_t i, int s) {
    const size_t off = (blockIdx.x + i * gridDim.x) * (size_t)PUSH_CHUNK;
    const uint32_t n = (uint32_t)(bytes - off < (size_t)PUSH_CHUNK ? bytes - off : (size_t)PUSH_CHUNK);
    const uint32_t bar = push_smem_u32(&full[s]);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(n) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     push_smem_u32(push_smem + (size_t)s * PUSH_CHUNK)),
                 "l"(src + off), "r"(n), "r"(bar)
                 : "memory");
  };
  const size_t mine = n_chunks > blockIdx.x ? (n_chunks - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  for (size_t i = 0; i < mine && i < (size_t)PUSH_DIST; ++i) load(i, (int)(i % PUSH_STAGES));
  for (size_t i = 0; i < mine; ++i) {
    const int s = (int)(i % PUSH_STAGES);
    if (i + PUSH_DIST < mine) {
      // the stage of chunk i + DIST was last read by the stores of chunk i + DIST - STAGES: only the PENDING newest
      // store groups (chunks i - PENDING .. i - 1) may still be reading shared memory
      asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(PUSH_PENDING) : "memory");
      load(i + PUSH_DIST, (int)((i + PUSH_DIST) % PUSH_STAGES));
    }
    const uint32_t bar = push_smem_u32(&full[s]);
    uint32_t ok = 0;
    while (!ok) {
      asm volatile(
          "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
          : "=r"(ok)
          : "r"(bar), "r"(phase[s])
          : "memory");
    }
    phase[s] ^= 1;
    const size_t off = (blockIdx.x + i * gridDim.x) * (size_t)PUSH_CHUNK;
    const uint32_t n = (uint32_t)(bytes - off < (size_t)PUSH_CHUNK ? bytes - off : (size_t)PUSH_CHUNK);
#pragma unroll
    for (int p = 0; p < PUSH_MAX_PEERS; ++p)
      if (p < pa.n_dst)
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(pa.dst[p] + off),
                     "r"(push_smem_u32(push_smem + (size_t)s * PUSH_CHUNK)), "r"(n)
                     : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  }
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // all stores complete (visible after the stream-ordered barrier)
}

}  // namespace

// src: this rank's slot (device, 16-byte aligned); host_dst[p]: the same slot in peer p's buffer (peer-mapped device
// pointers, 16-byte aligned); bytes: multiple of 16.  n_ctas: CTAs (= SMs) spent on the transfer (0 = 16).
extern "C" int fitgnn_peer_push(const void* src, void* const* host_dst, int n_dst, size_t bytes, int n_ctas, void* stream) {
  FG_REQUIRE(src && host_dst && n_dst >= 1 && n_dst <= PUSH_MAX_PEERS, FITGNN_EINVAL, "peer_push: 1..8 destinations");
  FG_REQUIRE(bytes % 16 == 0 && ((uintptr_t)src & 15) == 0, FITGNN_EUNSUP, "peer_push: 16-byte aligned source and size");
  if (bytes == 0) return FITGNN_OK;
  PushArgs pa{};
  pa.n_dst = n_dst;
  for (int p = 0; p < n_dst; ++p) {
    FG_REQUIRE(host_dst[p] && ((uintptr_t)host_dst[p] & 15) == 0, FITGNN_EUNSUP, "peer_push: destination %d unaligned / null", p);
    pa.dst[p] = static_cast<char*>(host_dst[p]);
  }
  if (n_ctas <= 0) n_ctas = 16;
  const size_t n_chunks = (bytes + PUSH_CHUNK - 1) / PUSH_CHUNK;
  if ((size_t)n_ctas > n_chunks) n_ctas = (int)n_chunks;
  const size_t smem = (size_t)PUSH_STAGES * PUSH_CHUNK;
  static bool configured = false;
  if (!configured) {
    FG_CUDA(cudaFuncSetAttribute(peer_push_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  if (n_ctas % 2 == 0) {
    // clusters of two: a cluster's CTAs share a TPC, so the kernel occupies WHOLE TPCs and the CTA-pair GEMM running beside it
    // (tuning sm_reserve) loses exactly n_ctas SMs, not up to 2 x n_ctas half-blocked TPCs
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)n_ctas);
    cfg.blockDim = dim3(32);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = as_stream(stream);
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    FG_CUDA(cudaLaunchKernelEx(&cfg, peer_push_kernel, static_cast<const char*>(src), pa, bytes));
    return FITGNN_OK;
  }
  peer_push_kernel<<<n_ctas, 32, smem, as_stream(stream)>>>(static_cast<const char*>(src), pa, bytes);
  FG_LAUNCH_CHECK();
  return FITGNN_OK;
}
