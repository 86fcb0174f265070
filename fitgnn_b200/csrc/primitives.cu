// Device-wide primitives for the integer builders: exclusive scan, LSD radix sort of 64-bit
// keys (optionally with a 32-bit payload), and segment pointers from sorted keys.
// Hand-written (no CUB/Thrust); all HBM-bound streaming kernels, grids sized by tile count.
#include "common.cuh"

namespace fitgnn {

// ------------------------------------------------------------------------------------------
// exclusive scan (int32), three-phase recursive: tile scan -> scan of tile sums -> add
// ------------------------------------------------------------------------------------------
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__global__ void __launch_bounds__(SCAN_THREADS)
scan_tile_kernel(const int32_t* in, int64_t n_in, int32_t* out, int64_t n_out,
                 int32_t* __restrict__ tile_sums) {
  __shared__ int32_t warp_sums[SCAN_THREADS / 32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
  int32_t v[SCAN_ITEMS];
  int32_t sum = 0;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) {
    v[i] = (base + i < n_in) ? in[base + i] : 0;
    sum += v[i];
  }
  int32_t incl = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int32_t t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) warp_sums[w] = incl;
  __syncthreads();
  if (w == 0) {
    int32_t ws = lane < SCAN_THREADS / 32 ? warp_sums[lane] : 0;
    int32_t wi = ws;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int32_t t = __shfl_up_sync(0xffffffffu, wi, o);
      if (lane >= o) wi += t;
    }
    if (lane < SCAN_THREADS / 32) warp_sums[lane] = wi - ws;
    if (lane == SCAN_THREADS / 32 - 1 && tile_sums) tile_sums[blockIdx.x] = wi;
  }
  __syncthreads();
  int32_t run = warp_sums[w] + incl - sum;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) {
    if (base + i < n_out) out[base + i] = run;
    run += v[i];
  }
}

__global__ void __launch_bounds__(SCAN_THREADS)
scan_add_kernel(int32_t* out, int64_t n_out, const int32_t* __restrict__ tile_offs) {
  const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
  const int32_t add = tile_offs[blockIdx.x];
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i)
    if (base + i < n_out) out[base + i] += add;
}

size_t scan_ws_bytes(int64_t n) {
  size_t total = 0;
  int64_t m = ceil_div(n > 0 ? n : 1, SCAN_TILE);
  while (m > 1) {
    total += align_up((size_t)m * sizeof(int32_t));
    m = ceil_div(m, SCAN_TILE);
  }
  return total + 256;
}

int scan_i32(const int32_t* in, int64_t n_in, int32_t* out, int64_t n_out, void* ws, size_t ws_bytes,
             cudaStream_t st) {
  if (n_out <= 0) return FITGNN_OK;
  const int64_t tiles = ceil_div(n_out, SCAN_TILE);
  FG_REQUIRE(tiles < (1ll << 31), FITGNN_ERANGE, "scan: too many elements (%lld)", (long long)n_out);
  if (tiles == 1) {
    scan_tile_kernel<<<1, SCAN_THREADS, 0, st>>>(in, n_in, out, n_out, nullptr);
    FG_LAUNCH_CHECK();
    return FITGNN_OK;
  }
  Bump b(ws, ws_bytes);
  int32_t* sums = b.take<int32_t>((size_t)tiles);
  FG_REQUIRE(b.ok, FITGNN_EWS, "scan: workspace too small");
  scan_tile_kernel<<<(unsigned)tiles, SCAN_THREADS, 0, st>>>(in, n_in, out, n_out, sums);
  FG_LAUNCH_CHECK();
  FG_TRY(scan_i32(sums, tiles, sums, tiles, b.here(), b.left(), st));
  scan_add_kernel<<<(unsigned)tiles, SCAN_THREADS, 0, st>>>(out, n_out, sums);
  FG_LAUNCH_CHECK();
  return FITGNN_OK;
}

// ------------------------------------------------------------------------------------------
// LSD radix sort, 8-bit digits, stable.  Per pass: per-tile digit histogram -> exclusive scan
// over [digit][tile] -> stable scatter (warp match_any ranks + cross-warp prefix in smem).
// ------------------------------------------------------------------------------------------
constexpr int RS_THREADS = 256;
constexpr int RS_ROUNDS = 16;
constexpr int RS_TILE = RS_THREADS * RS_ROUNDS;
constexpr int RS_WARPS = RS_THREADS / 32;

__global__ void __launch_bounds__(RS_THREADS)
rs_hist_kernel(const uint64_t* __restrict__ keys, int64_t n, int shift, int32_t* __restrict__ hist,
               int ntiles) {
  __shared__ int32_t h[256];
  h[threadIdx.x] = 0;
  __syncthreads();
  const int64_t base = (int64_t)blockIdx.x * RS_TILE;
#pragma unroll 4
  for (int r = 0; r < RS_ROUNDS; ++r) {
    int64_t idx = base + r * RS_THREADS + threadIdx.x;
    if (idx < n) atomicAdd(&h[(int)((keys[idx] >> shift) & 255u)], 1);
  }
  __syncthreads();
  hist[(size_t)threadIdx.x * ntiles + blockIdx.x] = h[threadIdx.x];
}

template <bool HAS_VALS>
__global__ void __launch_bounds__(RS_THREADS)
rs_scatter_kernel(const uint64_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
                  uint64_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out, int64_t n, int shift,
                  const int32_t* __restrict__ offs, int ntiles) {
  __shared__ int32_t base[256];
  __shared__ int32_t wcount[RS_WARPS][256];
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  base[tid] = offs[(size_t)tid * ntiles + blockIdx.x];
#pragma unroll
  for (int i = 0; i < RS_WARPS; ++i) wcount[i][tid] = 0;
  __syncthreads();
  const int64_t tile_base = (int64_t)blockIdx.x * RS_TILE;
  for (int r = 0; r < RS_ROUNDS; ++r) {
    const int64_t idx = tile_base + r * RS_THREADS + tid;
    if (tile_base + (int64_t)r * RS_THREADS >= n) break;  // uniform across the block
    const bool valid = idx < n;
    uint64_t key = 0;
    uint32_t val = 0;
    if (valid) {
      key = keys_in[idx];
      if (HAS_VALS) val = vals_in[idx];
    }
    const int d = valid ? (int)((key >> shift) & 255u) : 256 + lane;  // invalid lanes never match
    const unsigned peers = __match_any_sync(0xffffffffu, d);
    const int rank = __popc(peers & ((1u << lane) - 1u));
    if (valid && rank == 0) wcount[w][d] = __popc(peers);
    __syncthreads();
    // thread `tid` owns digit `tid`: exclusive prefix of the per-warp counts
    int32_t run = 0;
#pragma unroll
    for (int i = 0; i < RS_WARPS; ++i) {
      int32_t c = wcount[i][tid];
      wcount[i][tid] = run;
      run += c;
    }
    __syncthreads();
    if (valid) {
      const int32_t dst = base[d] + wcount[w][d] + rank;
      keys_out[dst] = key;
      if (HAS_VALS) vals_out[dst] = val;
    }
    __syncthreads();
    base[tid] += run;
#pragma unroll
    for (int i = 0; i < RS_WARPS; ++i) wcount[i][tid] = 0;
    __syncthreads();
  }
}

size_t sort_ws_bytes(int64_t n) {
  if (n <= 0) n = 1;
  const int64_t tiles = ceil_div(n, RS_TILE);
  return align_up((size_t)n * 8) + align_up((size_t)n * 4) + align_up((size_t)tiles * 256 * 4) +
         scan_ws_bytes(tiles * 256) + 1024;
}

int sort_u64(uint64_t* keys, uint32_t* vals, int64_t n, int key_bits, void* ws, size_t ws_bytes,
             cudaStream_t st) {
  if (n <= 1) return FITGNN_OK;
  FG_REQUIRE(key_bits >= 1 && key_bits <= 64, FITGNN_EINVAL, "sort: key_bits=%d", key_bits);
  FG_REQUIRE(n < (1ll << 31) - RS_TILE, FITGNN_ERANGE, "sort: n=%lld exceeds int32 offsets", (long long)n);
  const int64_t tiles = ceil_div(n, RS_TILE);
  Bump b(ws, ws_bytes);
  uint64_t* kalt = b.take<uint64_t>((size_t)n);
  uint32_t* valt = b.take<uint32_t>((size_t)n);
  int32_t* hist = b.take<int32_t>((size_t)tiles * 256);
  FG_REQUIRE(b.ok, FITGNN_EWS, "sort: workspace too small (%zu bytes for n=%lld)", ws_bytes, (long long)n);
  uint64_t* kin = keys;
  uint64_t* kout = kalt;
  uint32_t* vin = vals;
  uint32_t* vout = valt;
  const int passes = (key_bits + 7) / 8;
  for (int p = 0; p < passes; ++p) {
    const int shift = p * 8;
    rs_hist_kernel<<<(unsigned)tiles, RS_THREADS, 0, st>>>(kin, n, shift, hist, (int)tiles);
    FG_LAUNCH_CHECK();
    FG_TRY(scan_i32(hist, tiles * 256, hist, tiles * 256, b.here(), b.left(), st));
    if (vals)
      rs_scatter_kernel<true><<<(unsigned)tiles, RS_THREADS, 0, st>>>(kin, vin, kout, vout, n, shift, hist,
                                                                      (int)tiles);
    else
      rs_scatter_kernel<false><<<(unsigned)tiles, RS_THREADS, 0, st>>>(kin, nullptr, kout, nullptr, n, shift,
                                                                       hist, (int)tiles);
    FG_LAUNCH_CHECK();
    uint64_t* tk = kin; kin = kout; kout = tk;
    uint32_t* tv = vin; vin = vout; vout = tv;
  }
  if (kin != keys) {
    FG_CUDA(cudaMemcpyAsync(keys, kin, (size_t)n * 8, cudaMemcpyDeviceToDevice, st));
    if (vals) FG_CUDA(cudaMemcpyAsync(vals, vin, (size_t)n * 4, cudaMemcpyDeviceToDevice, st));
  }
  return FITGNN_OK;
}

// ------------------------------------------------------------------------------------------
// ptr[r] = lower bound of row r in the sorted keys (row = key >> shift), r in [0, n_rows]
// ------------------------------------------------------------------------------------------
__global__ void segment_ptr_kernel(const uint64_t* __restrict__ keys, int64_t n_keys, int shift,
                                   int64_t n_rows, int32_t* __restrict__ ptr) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i > n_keys) return;
  // rows in (prev, cur] start at i; i == n_keys closes every remaining row
  int64_t prev = (i == 0) ? -1 : (int64_t)(keys[i - 1] >> shift);
  int64_t cur = (i == n_keys) ? n_rows : (int64_t)(keys[i] >> shift);
  if (cur > n_rows) cur = n_rows;
  if (prev >= n_rows) return;
  for (int64_t r = prev + 1; r <= cur; ++r) ptr[r] = (int32_t)i;
}

int segment_ptr_from_sorted(const uint64_t* keys, int64_t n_keys, int shift, int64_t n_rows, int32_t* ptr,
                            cudaStream_t st) {
  const int threads = 256;
  segment_ptr_kernel<<<(unsigned)ceil_div(n_keys + 1, threads), threads, 0, st>>>(keys, n_keys, shift, n_rows,
                                                                                 ptr);
  FG_LAUNCH_CHECK();
  return FITGNN_OK;
}

}  // namespace fitgnn

// ------------------------------------------------------------------------------------------
// exported test hooks
// ------------------------------------------------------------------------------------------
using namespace fitgnn;

extern "C" size_t fitgnn_sort_workspace_bytes(int64_t n) { return sort_ws_bytes(n); }
extern "C" size_t fitgnn_scan_workspace_bytes(int64_t n) { return scan_ws_bytes(n + 1); }

extern "C" int fitgnn_sort_u64(uint64_t* keys, uint32_t* vals, int64_t n, int key_bits, void* ws,
                               size_t ws_bytes, void* stream) {
  FG_REQUIRE(n >= 0 && (n == 0 || keys), FITGNN_EINVAL, "sort_u64: bad arguments");
  return sort_u64(keys, vals, n, key_bits, ws, ws_bytes, as_stream(stream));
}

extern "C" int fitgnn_scan_i32(const int32_t* in, int32_t* out, int64_t n, int with_total, void* ws,
                               size_t ws_bytes, void* stream) {
  FG_REQUIRE(n >= 0 && (n == 0 || (in && out)), FITGNN_EINVAL, "scan_i32: bad arguments");
  return scan_i32(in, n, out, n + (with_total ? 1 : 0), ws, ws_bytes, as_stream(stream));
}
