// Exact-fp32 dense transform on CUDA cores: Y[M,N] = act(A[M,K] · W[N,K]^T + bias).
// Replaces GCNConv.lin / lt1 (F.linear) at /root/reference/network.py:31,34.
// This is the FITGNN_GEMM_FP32 arithmetic (bit-faithful fp32 FMA accumulation); it is the
// numerics anchor for the tcgen05 path and the path for shapes the tensor-core kernel does not
// take.  128x128x16 tiles, 256 threads, 8x8 register micro-tiles, both operands K-major in
// global memory and transposed into shared memory so the inner product reads float4 rows.
// Also: the row-wise (log-)softmax head and the fp32 -> bf16 hi/lo split.
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include "common.cuh"

namespace fitgnn {

constexpr int GM_BM = 128, GM_BN = 128, GM_BK = 16, GM_THREADS = 256;

__device__ __forceinline__ float elu1g(float x) { return x > 0.f ? x : expm1f(x); }

__global__ void __launch_bounds__(GM_THREADS)
gemm_fp32_kernel(const float* __restrict__ A, int64_t lda, const float* __restrict__ W, int64_t ldw,
                 const float* __restrict__ bias, int64_t M, int K, int N, int act, float* __restrict__ Y,
                 int64_t ldy) {
  __shared__ __align__(16) float As[2][GM_BK][GM_BM + 4];
  __shared__ __align__(16) float Ws[2][GM_BK][GM_BN + 4];
  const int tid = threadIdx.x;
  const int64_t m0 = (int64_t)blockIdx.x * GM_BM;
  const int n0 = blockIdx.y * GM_BN;
  // loader mapping: 128 rows x 16 k per tile = 2048 floats = 8 per thread: row = tid/2, k-offset (tid%2)*8
  const int lrow = tid >> 1, lk = (tid & 1) * 8;
  const bool a_vec = (lda % 4 == 0) && (((uintptr_t)A & 15) == 0);
  const bool w_vec = (ldw % 4 == 0) && (((uintptr_t)W & 15) == 0);
  // compute mapping: 16x16 threads, each 8 rows x 8 cols
  const int ty = tid >> 4, tx = tid & 15;
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  float ra[8], rw[8];
  auto load_tile = [&](int k0) {
    const int64_t am = m0 + lrow;
    const int wn = n0 + lrow;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int k = k0 + lk + 4 * h;
      if (am < M && a_vec && k + 3 < K) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(A + am * lda + k));
        ra[4 * h] = v.x; ra[4 * h + 1] = v.y; ra[4 * h + 2] = v.z; ra[4 * h + 3] = v.w;
      } else {
#pragma unroll
        for (int u = 0; u < 4; ++u) ra[4 * h + u] = (am < M && k + u < K) ? __ldg(A + am * lda + k + u) : 0.f;
      }
      if (wn < N && w_vec && k + 3 < K) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(W + (int64_t)wn * ldw + k));
        rw[4 * h] = v.x; rw[4 * h + 1] = v.y; rw[4 * h + 2] = v.z; rw[4 * h + 3] = v.w;
      } else {
#pragma unroll
        for (int u = 0; u < 4; ++u) rw[4 * h + u] = (wn < N && k + u < K) ? __ldg(W + (int64_t)wn * ldw + k + u) : 0.f;
      }
    }
  };
  auto store_tile = [&](int buf) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      As[buf][lk + u][lrow] = ra[u];
      Ws[buf][lk + u][lrow] = rw[u];
    }
  };

  const int nk = (K + GM_BK - 1) / GM_BK;
  load_tile(0);
  store_tile(0);
  __syncthreads();
  for (int kb = 0; kb < nk; ++kb) {
    const int buf = kb & 1;
    if (kb + 1 < nk) load_tile((kb + 1) * GM_BK);
#pragma unroll
    for (int k = 0; k < GM_BK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 8]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 8 + 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Ws[buf][k][tx * 8]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Ws[buf][k][tx * 8 + 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (kb + 1 < nk) {
      store_tile(buf ^ 1);
      __syncthreads();
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int64_t m = m0 + ty * 8 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int n = n0 + tx * 8 + j;
      if (n >= N) continue;
      float v = acc[i][j] + (bias ? __ldg(bias + n) : 0.f);
      if (act == FITGNN_ACT_ELU) v = elu1g(v);
      Y[m * ldy + n] = v;
    }
  }
}

// row-wise softmax / log_softmax in place, one warp per row (torch.log_softmax semantics:
// x - max - log(sum(exp(x - max))))
__global__ void row_softmax_kernel(float* __restrict__ Y, int64_t ldy, int64_t M, int N, int head) {
  const int lane = threadIdx.x & 31;
  const int64_t m = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (m >= M) return;
  float* y = Y + m * ldy;
  float mx = -INFINITY;
  for (int n = lane; n < N; n += 32) mx = fmaxf(mx, y[n]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  float s = 0.f;
  for (int n = lane; n < N; n += 32) s += expf(y[n] - mx);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (head == FITGNN_HEAD_LOG_SOFTMAX) {
    const float ls = logf(s);
    for (int n = lane; n < N; n += 32) y[n] = y[n] - mx - ls;
  } else {
    const float inv = 1.f / s;
    for (int n = lane; n < N; n += 32) y[n] = expf(y[n] - mx) * inv;
  }
}

__global__ void split_bf16_kernel(const float* __restrict__ X, int64_t ldx, int64_t rows, int cols,
                                  __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, int64_t ldo) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t r = idx / ldo;
  const int c = (int)(idx % ldo);
  if (r >= rows) return;
  const float x = c < cols ? X[r * ldx + c] : 0.f;
  const __nv_bfloat16 h = __float2bfloat16_rn(x);
  hi[idx] = h;
  lo[idx] = __float2bfloat16_rn(x - __bfloat162float(h));
}

// fp32 -> fp16 hi/lo planes (hi = rn(x), lo = rn(x - hi): 22 significant bits; lo == nullptr: the hi plane only)
__global__ void split_f16_kernel(const float* __restrict__ X, int64_t ldx, int64_t rows, int cols, __half* __restrict__ hi,
                                 __half* __restrict__ lo, int64_t ldo) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t r = idx / ldo;
  const int c = (int)(idx % ldo);
  if (r >= rows) return;
  const float x = c < cols ? X[r * ldx + c] : 0.f;
  const float xc = fminf(fmaxf(x, -65504.f), 65504.f);  // saturate instead of inf
  const __half h = __float2half_rn(xc);
  hi[idx] = h;
  if (lo) lo[idx] = __float2half_rn(xc - __half2float(h));
}

int gemm_fp32(const float* A, int64_t lda, const float* W, int64_t ldw, const float* bias, int64_t M, int K, int N,
              int act, float* Y, int64_t ldy, cudaStream_t st) {
  dim3 grid((unsigned)ceil_div(M, GM_BM), (unsigned)ceil_div(N, GM_BN));
  gemm_fp32_kernel<<<grid, GM_THREADS, 0, st>>>(A, lda, W, ldw, bias, M, K, N, act, Y, ldy);
  FG_LAUNCH_CHECK();
  return FITGNN_OK;
}

int row_softmax(float* Y, int64_t ldy, int64_t M, int N, int head, cudaStream_t st) {
  if (head == FITGNN_HEAD_IDENTITY || M == 0) return FITGNN_OK;
  row_softmax_kernel<<<(unsigned)ceil_div(M, 8), 256, 0, st>>>(Y, ldy, M, N, head);
  FG_LAUNCH_CHECK();
  return FITGNN_OK;
}

}  // namespace fitgnn

using namespace fitgnn;

extern "C" int fitgnn_split_bf16(const float* X, int64_t ldx, int64_t rows, int cols, void* hi, void* lo, int64_t ldo,
                                 void* stream) {
  FG_REQUIRE(X && hi && lo && rows >= 0 && cols >= 0 && ldo >= cols && ldx >= cols, FITGNN_EINVAL,
             "split_bf16: bad arguments");
  if (rows == 0 || ldo == 0) return FITGNN_OK;
  const int64_t total = rows * ldo;
  split_bf16_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, as_stream(stream)>>>(
      X, ldx, rows, cols, static_cast<__nv_bfloat16*>(hi), static_cast<__nv_bfloat16*>(lo), ldo);
  FG_LAUNCH_CHECK();
  return FITGNN_OK;
}

extern "C" int fitgnn_split_f16(const float* X, int64_t ldx, int64_t rows, int cols, void* hi, void* lo, int64_t ldo,
                                void* stream) {
  FG_REQUIRE(X && hi && rows >= 0 && cols >= 0 && ldo >= cols && ldx >= cols, FITGNN_EINVAL, "split_f16: bad arguments");
  if (rows == 0 || ldo == 0) return FITGNN_OK;
  const int64_t total = rows * ldo;
  split_f16_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, as_stream(stream)>>>(X, ldx, rows, cols, static_cast<__half*>(hi),
                                                                                  static_cast<__half*>(lo), ldo);
  FG_LAUNCH_CHECK();
  return FITGNN_OK;
}
