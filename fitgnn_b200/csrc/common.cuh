// Shared host/device helpers for libfitgnn_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include "../../include/fitgnn.h"

namespace fitgnn {

void set_error(const char* fmt, ...);

#define FG_CUDA(call)                                                                      \
  do {                                                                                     \
    cudaError_t e_ = (call);                                                               \
    if (e_ != cudaSuccess) {                                                               \
      ::fitgnn::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
      return FITGNN_ECUDA;                                                                 \
    }                                                                                      \
  } while (0)
#define FG_LAUNCH_CHECK() FG_CUDA(cudaGetLastError())
#define FG_REQUIRE(cond, code, ...)        \
  do {                                     \
    if (!(cond)) {                         \
      ::fitgnn::set_error(__VA_ARGS__);    \
      return code;                         \
    }                                      \
  } while (0)
#define FG_TRY(expr)            \
  do {                          \
    int rc_ = (expr);           \
    if (rc_ != FITGNN_OK) return rc_; \
  } while (0)

static inline size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }
static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
// number of bits needed to represent every value in [0, max_value]
static inline int bits_for(uint64_t max_value) {
  int b = 1;
  while (b < 64 && (max_value >> b) != 0) ++b;
  return b;
}

// bump allocator over a caller-provided workspace
struct Bump {
  char* base;
  size_t off;
  size_t cap;
  bool ok;
  Bump(void* p, size_t bytes) : base(static_cast<char*>(p)), off(0), cap(bytes), ok(true) {}
  template <class T>
  T* take(size_t n) {
    size_t bytes = align_up(n * sizeof(T));
    if (off + bytes > cap) {
      ok = false;
      return nullptr;
    }
    T* r = reinterpret_cast<T*>(base + off);
    off += bytes;
    return r;
  }
  size_t left() const { return cap - off; }
  void* here() const { return base + off; }
};

static inline cudaStream_t as_stream(void* s) { return static_cast<cudaStream_t>(s); }

// SM count of the current device (queried once per device; grids are sized in multiples of it)
int sm_count();
// kernel-tuning switches, read from the environment ONCE per process (FITGNN_GEMM_WS, FITGNN_HEAD_BULK,
// FITGNN_AGG_WIDE, FITGNN_GEMM_WIDE, FITGNN_GEMM_PAIR, FITGNN_SM_RESERVE, FITGNN_GEMM_PAIR_WS, FITGNN_GEMM_PREFETCH)
struct Tuning {
  int gemm_ws;    // 0 = force the streaming smem plan (default 1)
  int head_bulk;  // 0 = per-thread stores in row-mapped heads (default 1)
  int agg_wide;   // fused aggregation tile: 0 = default, 1 = 128 columns / 16 epilogue warps, 2 = 256 / 12 (fp16 output only), -1 = 256 / 8
  int gemm_wide;  // 1 = same tile for small-K wide-output transforms (default 0)
  int gemm_pair;  // 0 = no CTA pairs (default 1)
  int sm_reserve; // SMs the persistent GEMMs leave free for a concurrent exchange kernel (default 0)
  int gemm_pair_ws;  // 0 = CTA pairs never keep their weights resident (default 1: when they fit)
  int gemm_debug;    // measurement switches, never set in production: bit 0 = GEMM epilogues skip their TMA stores
  int gemm_prefetch; // 1 = L2 prefetch of the next m-block's A rows by the TMA producer (default 0: measured slower, r2aj)
};
const Tuning& tuning();

#ifdef __CUDACC__
// ---- device helpers shared by the SpMM and GEMM epilogues -------------------------------------------------------
// two fp32 -> packed bf16x2 (round to nearest even), `a` in the low half: one F2FP on the ALU pipe instead of two
// single conversions on the quarter-rate XU pipe plus a PRMT
__device__ __forceinline__ uint32_t pack_bf16x2_rn(float a, float b) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  return r;
}
// two fp32 -> packed fp16x2 (round to nearest even, overflow saturates to the largest finite fp16), `a` in the low half
__device__ __forceinline__ uint32_t pack_f16x2_rn(float a, float b) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  return r;
}
// hi/lo bf16 split of two fp32 values: hi = rn_bf16(x), lo = rn_bf16(x - hi), each packed as bf16x2 (x0 low half)
__device__ __forceinline__ void split_bf16x2(float x0, float x1, uint32_t& hi, uint32_t& lo) {
  hi = pack_bf16x2_rn(x0, x1);
  lo = pack_bf16x2_rn(x0 - __uint_as_float(hi << 16), x1 - __uint_as_float(hi & 0xffff0000u));
}
// ELU (alpha = 1, network.py:32): exp through one MUFU.EX2 (flush-to-zero), absolute error <= ~2e-7
__device__ __forceinline__ float elu_fast(float x) {
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * 1.4426950408889634f));
  return x > 0.f ? x : e - 1.0f;
}
#endif

// --- internal primitives (primitives.cu) ------------------------------------------------
size_t scan_ws_bytes(int64_t n);
// out[i] = sum_{j<i} in[j] for i < n_out, with in[j] = 0 for j >= n_in (so n_out = n_in+1 yields the total)
int scan_i32(const int32_t* in, int64_t n_in, int32_t* out, int64_t n_out, void* ws, size_t ws_bytes,
             cudaStream_t st);
// out[i] = number of run starts among sorted keys[0..i) (run start: i == 0 or keys[i] != keys[i-1]); n_out = n_keys + 1
// yields the number of distinct keys in out[n_keys]
int scan_key_boundaries(const uint64_t* keys, int64_t n_keys, int32_t* out, int64_t n_out, void* ws, size_t ws_bytes,
                        cudaStream_t st);
size_t sort_ws_bytes(int64_t n);
int sort_u64(uint64_t* keys, uint32_t* vals, int64_t n, int key_bits, void* ws, size_t ws_bytes,
             cudaStream_t st);
// same, on the bits set in `mask` only (bits outside the mask must be equal in all keys, or irrelevant to the order)
int sort_u64_mask(uint64_t* keys, uint32_t* vals, int64_t n, uint64_t mask, void* ws, size_t ws_bytes,
                  cudaStream_t st);
// mask of a (hi << 32 | lo) key whose fields hold values in [0, hi_max] and [0, lo_max]
static inline uint64_t field_mask(uint64_t hi_max, uint64_t lo_max) {
  const int hb = bits_for(hi_max), lb = bits_for(lo_max);
  const uint64_t lo = lb >= 32 ? 0xffffffffull : ((1ull << lb) - 1ull);
  const uint64_t hi = hb >= 32 ? 0xffffffffull : ((1ull << hb) - 1ull);
  return (hi << 32) | lo;
}
// ptr[r] = first index i with (keys[i] >> shift) >= r, for r in [0, n_rows]; keys sorted, n_keys valid
int segment_ptr_from_sorted(const uint64_t* keys, int64_t n_keys, int shift, int64_t n_rows, int32_t* ptr,
                            cudaStream_t st);

}  // namespace fitgnn
