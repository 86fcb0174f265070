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

// FROM_KEYS: the scanned value is the run-boundary flag of sorted keys, (i == 0 || keys[i] != keys[i-1]), formed on the fly
// (saves writing and re-reading a flag array)
template <bool FROM_KEYS>
__global__ void __launch_bounds__(SCAN_THREADS)
scan_tile_kernel(const void* in_, int64_t n_in, int32_t* out, int64_t n_out,
                 int32_t* __restrict__ tile_sums) {
  const int32_t* in = static_cast<const int32_t*>(in_);
  const uint64_t* keys = static_cast<const uint64_t*>(in_);
  __shared__ int32_t warp_sums[SCAN_THREADS / 32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
  int32_t v[SCAN_ITEMS];
  int32_t sum = 0;
  uint64_t prev = 0;
  if (FROM_KEYS && base > 0 && base - 1 < n_in) prev = keys[base - 1];
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) {
    if (FROM_KEYS) {
      const uint64_t cur = (base + i < n_in) ? keys[base + i] : 0ull;
      v[i] = (base + i < n_in) ? ((base + i == 0 || cur != prev) ? 1 : 0) : 0;
      prev = cur;
    } else {
      v[i] = (base + i < n_in) ? in[base + i] : 0;
    }
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

template <bool FROM_KEYS>
static int scan_impl(const void* in, int64_t n_in, int32_t* out, int64_t n_out, void* ws, size_t ws_bytes, cudaStream_t st);

int scan_i32(const int32_t* in, int64_t n_in, int32_t* out, int64_t n_out, void* ws, size_t ws_bytes,
             cudaStream_t st) {
  return scan_impl<false>(in, n_in, out, n_out, ws, ws_bytes, st);
}

int scan_key_boundaries(const uint64_t* keys, int64_t n_keys, int32_t* out, int64_t n_out, void* ws, size_t ws_bytes,
                        cudaStream_t st) {
  return scan_impl<true>(keys, n_keys, out, n_out, ws, ws_bytes, st);
}

template <bool FROM_KEYS>
static int scan_impl(const void* in, int64_t n_in, int32_t* out, int64_t n_out, void* ws, size_t ws_bytes, cudaStream_t st) {
  if (n_out <= 0) return FITGNN_OK;
  const int64_t tiles = ceil_div(n_out, SCAN_TILE);
  FG_REQUIRE(tiles < (1ll << 31), FITGNN_ERANGE, "scan: too many elements (%lld)", (long long)n_out);
  if (tiles == 1) {
    scan_tile_kernel<FROM_KEYS><<<1, SCAN_THREADS, 0, st>>>(in, n_in, out, n_out, nullptr);
    FG_LAUNCH_CHECK();
    return FITGNN_OK;
  }
  Bump b(ws, ws_bytes);
  int32_t* sums = b.take<int32_t>((size_t)tiles);
  FG_REQUIRE(b.ok, FITGNN_EWS, "scan: workspace too small");
  scan_tile_kernel<FROM_KEYS><<<(unsigned)tiles, SCAN_THREADS, 0, st>>>(in, n_in, out, n_out, sums);
  FG_LAUNCH_CHECK();
  FG_TRY(scan_i32(sums, tiles, sums, tiles, b.here(), b.left(), st));
  scan_add_kernel<<<(unsigned)tiles, SCAN_THREADS, 0, st>>>(out, n_out, sums);
  FG_LAUNCH_CHECK();
  return FITGNN_OK;
}

// ------------------------------------------------------------------------------------------
// LSD radix sort, digits of <= 8 bits, stable.  Per pass: per-tile digit histogram -> exclusive scan over
// [digit][tile] -> stable scatter.  The passes cover only the bits set in `mask` (contiguous runs of set bits, each
// split into near-equal digits), so a (hi << 32 | lo) key with 22 + 21 significant bits takes 6 passes, not 7.
//
// Scatter: a warp owns 512 consecutive keys of the tile (16 register-resident keys per thread, all loads in flight
// at once), ranks them with match_any against its private digit counters (warp-level syncs only), one block-level
// prefix over (digit, warp) turns the ranks into tile-local sorted positions, the keys are exchanged through shared
// memory and leave in runs of consecutive addresses per digit (coalesced stores; ~16 keys = 128 bytes per run).
// ------------------------------------------------------------------------------------------
constexpr int RS_THREADS = 256;
constexpr int RS_ITEMS = 16;
constexpr int RS_TILE = RS_THREADS * RS_ITEMS;
constexpr int RS_WARPS = RS_THREADS / 32;

// Per-tile digit histogram.  No warp collectives and no shared atomics: ncu showed MATCH.ANY (and VOTE) bound on the ADU
// pipe (~57 cycles per warp instruction and SM, pipe_adu 98 % busy) and ATOMS at 2 cycles per lane.  Every thread counts
// its 16 keys into PRIVATE byte counters (four digits per 32-bit word, column `tid` of cnt[word][256]: bank = tid % 32,
// conflict-free plain LDS/STS; a count is at most 16), then word w's column is summed with packed 16-bit adds by thread
// w, walking the columns skewed by its own index so that the lanes of a warp stay on distinct banks.
__global__ void __launch_bounds__(RS_THREADS)
rs_hist_kernel(const uint64_t* __restrict__ keys, int64_t n, int shift, int nbits, int32_t* __restrict__ hist,
               int ntiles) {
  extern __shared__ uint32_t rs_cnt[];  // [nwords][RS_THREADS]
  const int tid = threadIdx.x;
  const uint32_t dmask = (1u << nbits) - 1u;
  const int ndig = 1 << nbits, nwords = ndig >= 4 ? ndig >> 2 : 1;
  for (int i = 0; i < nwords; ++i) rs_cnt[i * RS_THREADS + tid] = 0;  // own column only: no barrier needed yet
  const int64_t base = (int64_t)blockIdx.x * RS_TILE;
  uint64_t key[RS_ITEMS];
#pragma unroll
  for (int r = 0; r < RS_ITEMS; ++r) {
    const int64_t idx = base + r * RS_THREADS + tid;
    key[r] = idx < n ? keys[idx] : 0ull;
  }
#pragma unroll
  for (int r = 0; r < RS_ITEMS; ++r) {
    if (base + r * RS_THREADS + tid < n) {
      const uint32_t d = (uint32_t)(key[r] >> shift) & dmask;
      rs_cnt[(d >> 2) * RS_THREADS + tid] += 1u << ((d & 3u) * 8u);
    }
  }
  __syncthreads();
  if (tid < nwords) {
    uint32_t even = 0, odd = 0;  // 16-bit fields: digits (4w, 4w+2) and (4w+1, 4w+3); 256 columns x 16 keys fit
#pragma unroll 8
    for (int i = 0; i < RS_THREADS; ++i) {
      const uint32_t v = rs_cnt[tid * RS_THREADS + ((i + tid) & (RS_THREADS - 1))];
      even += v & 0x00ff00ffu;
      odd += (v >> 8) & 0x00ff00ffu;
    }
    const int d0 = tid * 4;
    const int32_t c[4] = {(int32_t)(even & 0xffffu), (int32_t)(odd & 0xffffu), (int32_t)(even >> 16), (int32_t)(odd >> 16)};
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (d0 + j < ndig) hist[(size_t)(d0 + j) * ntiles + blockIdx.x] = c[j];
  }
  // digits >= ndig never occur: their rows of the [256][ntiles] table must read as zero for the scan
  for (int d = ndig + tid; d < 256; d += RS_THREADS) hist[(size_t)d * ntiles + blockIdx.x] = 0;
}

constexpr size_t rs_hist_smem(int nbits) { return (size_t)((1 << nbits) >= 4 ? (1 << nbits) >> 2 : 1) * RS_THREADS * 4; }

constexpr size_t rs_scatter_smem(bool has_vals) {
  return (size_t)RS_TILE * 8 + (size_t)(RS_WARPS * 256 + 256 + 256 + 32) * 4 + (has_vals ? (size_t)RS_TILE * 4 : 0);
}

template <bool HAS_VALS>
__global__ void __launch_bounds__(RS_THREADS, HAS_VALS ? 2 : 3)
rs_scatter_kernel(const uint64_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
                  uint64_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out, int64_t n, int shift,
                  int nbits, const int32_t* __restrict__ offs, int ntiles) {
  const uint32_t dmask = (1u << nbits) - 1u;
  extern __shared__ __align__(16) unsigned char rs_smem[];
  uint64_t* skeys = reinterpret_cast<uint64_t*>(rs_smem);   // [RS_TILE] keys at their tile-local sorted position
  int32_t* wcount = reinterpret_cast<int32_t*>(skeys + RS_TILE);  // [RS_WARPS][256]
  int32_t* lstart = wcount + RS_WARPS * 256;                // [256] tile-local start of each digit
  int32_t* gbase = lstart + 256;                            // [256] global start minus tile-local start
  int32_t* wsum = gbase + 256;                              // [32]
  uint32_t* svals = reinterpret_cast<uint32_t*>(wsum + 32); // [RS_TILE] (HAS_VALS)
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int64_t tile_base = (int64_t)blockIdx.x * RS_TILE;
  const int count = (int)((n - tile_base) < (int64_t)RS_TILE ? (n - tile_base) : (int64_t)RS_TILE);
  const int wbase = w * (RS_ITEMS * 32);
  int32_t* mycount = wcount + w * 256;

  uint64_t key[RS_ITEMS];
  uint32_t val[RS_ITEMS];
  int rank[RS_ITEMS];
#pragma unroll
  for (int r = 0; r < RS_ITEMS; ++r) {
    const int p = wbase + r * 32 + lane;
    key[r] = p < count ? keys_in[tile_base + p] : 0ull;
    if (HAS_VALS) val[r] = p < count ? vals_in[tile_base + p] : 0u;
  }
  const int32_t my_goff = offs[(size_t)tid * ntiles + blockIdx.x];
#pragma unroll
  for (int i = 0; i < RS_WARPS; ++i) wcount[i * 256 + tid] = 0;
  __syncthreads();

  // rank inside the warp's 512 keys: keys of earlier rounds first, then lower lanes (= input order)
#pragma unroll
  for (int r = 0; r < RS_ITEMS; ++r) {
    const bool valid = wbase + r * 32 + lane < count;
    const int d = valid ? (int)((uint32_t)(key[r] >> shift) & dmask) : 256 + lane;  // invalid lanes never match
    const unsigned peers = __match_any_sync(0xffffffffu, d);  // ADU-bound (see rs_hist_kernel), ~0.86 ms per 123 M keys
    const int below = __popc(peers & ((1u << lane) - 1u));
    int prev = 0;
    if (valid) prev = mycount[d];
    __syncwarp();
    if (valid && below == 0) mycount[d] = prev + __popc(peers);
    __syncwarp();
    rank[r] = prev + below;
  }
  __syncthreads();

  // thread `tid` owns digit `tid`: exclusive prefix over the warps, then over the digits
  int run = 0;
#pragma unroll
  for (int i = 0; i < RS_WARPS; ++i) {
    const int c = wcount[i * 256 + tid];
    wcount[i * 256 + tid] = run;
    run += c;
  }
  int incl = run;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) wsum[w] = incl;
  __syncthreads();
  int woff = 0;
#pragma unroll
  for (int i = 0; i < RS_WARPS; ++i) woff += (i < w) ? wsum[i] : 0;
  const int excl = woff + incl - run;
  lstart[tid] = excl;
  gbase[tid] = my_goff - excl;
  __syncthreads();

#pragma unroll
  for (int r = 0; r < RS_ITEMS; ++r) {
    if (wbase + r * 32 + lane < count) {
      const int d = (int)((uint32_t)(key[r] >> shift) & dmask);
      const int pos = lstart[d] + mycount[d] + rank[r];
      skeys[pos] = key[r];
      if (HAS_VALS) svals[pos] = val[r];
    }
  }
  __syncthreads();

#pragma unroll
  for (int i = 0; i < RS_ITEMS; ++i) {
    const int p = i * RS_THREADS + tid;
    if (p < count) {
      const uint64_t k = skeys[p];
      const int d = (int)((uint32_t)(k >> shift) & dmask);
      const int32_t dst = gbase[d] + p;
      keys_out[dst] = k;
      if (HAS_VALS) vals_out[dst] = svals[p];
    }
  }
}

size_t sort_ws_bytes(int64_t n) {
  if (n <= 0) n = 1;
  const int64_t tiles = ceil_div(n, RS_TILE);
  return align_up((size_t)n * 8) + align_up((size_t)n * 4) + align_up((size_t)tiles * 256 * 4) +
         scan_ws_bytes(tiles * 256) + 1024;
}

// pass plan: every contiguous run of set bits in `mask` is split into ceil(len / 8) digits of near-equal width
static int plan_passes(uint64_t mask, int* shifts, int* widths) {
  int np = 0;
  int b = 0;
  while (b < 64) {
    if (!((mask >> b) & 1ull)) { ++b; continue; }
    int e = b;
    while (e < 64 && ((mask >> e) & 1ull)) ++e;
    const int len = e - b, parts = (len + 7) / 8;
    int at = b;
    for (int i = 0; i < parts; ++i) {
      const int wd = len / parts + (i < len % parts ? 1 : 0);
      shifts[np] = at; widths[np] = wd; ++np;
      at += wd;
    }
    b = e;
  }
  return np;
}

int sort_u64_mask(uint64_t* keys, uint32_t* vals, int64_t n, uint64_t mask, void* ws, size_t ws_bytes,
                  cudaStream_t st) {
  if (n <= 1 || mask == 0) return FITGNN_OK;
  FG_REQUIRE(n < (1ll << 31) - RS_TILE, FITGNN_ERANGE, "sort: n=%lld exceeds int32 offsets", (long long)n);
  FG_CUDA(cudaFuncSetAttribute(rs_hist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rs_hist_smem(8)));
  if (vals)
    FG_CUDA(cudaFuncSetAttribute(rs_scatter_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)rs_scatter_smem(true)));
  else
    FG_CUDA(cudaFuncSetAttribute(rs_scatter_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)rs_scatter_smem(false)));
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
  int shifts[64], widths[64];
  const int passes = plan_passes(mask, shifts, widths);
  for (int p = 0; p < passes; ++p) {
    const int shift = shifts[p];
    const int nbits = widths[p];
    rs_hist_kernel<<<(unsigned)tiles, RS_THREADS, rs_hist_smem(nbits), st>>>(kin, n, shift, nbits, hist, (int)tiles);
    FG_LAUNCH_CHECK();
    FG_TRY(scan_i32(hist, tiles * 256, hist, tiles * 256, b.here(), b.left(), st));
    if (vals)
      rs_scatter_kernel<true><<<(unsigned)tiles, RS_THREADS, rs_scatter_smem(true), st>>>(
          kin, vin, kout, vout, n, shift, nbits, hist, (int)tiles);
    else
      rs_scatter_kernel<false><<<(unsigned)tiles, RS_THREADS, rs_scatter_smem(false), st>>>(
          kin, nullptr, kout, nullptr, n, shift, nbits, hist, (int)tiles);
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

int sort_u64(uint64_t* keys, uint32_t* vals, int64_t n, int key_bits, void* ws, size_t ws_bytes,
             cudaStream_t st) {
  FG_REQUIRE(key_bits >= 1 && key_bits <= 64, FITGNN_EINVAL, "sort: key_bits=%d", key_bits);
  return sort_u64_mask(keys, vals, n, key_bits == 64 ? ~0ull : ((1ull << key_bits) - 1ull), ws, ws_bytes, st);
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
