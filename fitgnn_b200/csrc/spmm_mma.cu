// Block-dense aggregation on the tensor cores: Y = act( D (M (D X[src])) + bias ) per block of whole subgraphs.
//
// Replaces, like spmm.cu, gcn_norm + MessagePassing.propagate of PyG GCNConv (call sites /root/reference/network.py:31,60,
// 90,126,161,197) — for packs whose subgraphs are DENSE: with cluster_node augmentation (utils.py:190-233) every cluster
// node of a subgraph is joined to the core nodes that see it and to every other cluster node adjacent in Ac (utils.py:224-232);
// on the ogbn-products-shaped graph that is ~25 entries per row inside subgraphs of ~96 rows (density ~25 %), 2.6e9 entries in
// all.  Gathering them entry by entry is bound by L2 traffic (spmm_pipe_kernel: 155 ms) or by shared-memory bandwidth
// (spmm_block_kernel: 84 ms, one LDS.128 per entry and 16 bytes).  But inside a block the normalised adjacency factors as
//     Â_s = D_s · M_s · D_s,     M_s = 0/1 adjacency with self loops (duplicate edges: small integer counts), D_s = diag(deg^-1/2)
// and M_s is EXACT in bf16.  So per 128 x 128 piece of M_s the aggregation is a small dense product on the tensor cores:
//     stage  Yt = D_s X[src] for the piece's source rows as bf16 hi/lo planes (Yt = hi + lo to 2^-17),
//     build  M (128 x 128 bf16) in shared memory from the CSR entries,
//     MMA    acc += M·hi + M·lo   (mma.sync m16n8k16, fp32 accumulate; 8 warps, 16 output rows each),
//     scale  rows by deg^-1/2, + bias, ELU, store fp32 or bf16 hi/lo planes.
// Work per entry drops from "16 bytes of shared-memory traffic per 4 features" to nothing (the entry is one bf16 in M); the
// cost is the 128 x 128 x width MMA per piece, ~4 TFLOP for the whole products cluster pack.  Differences to the gather
// kernels: the sum runs in MMA order and the sources are split to bf16 hi/lo, so results agree to ~1e-6 relative, not bit for
// bit.  Blocks larger than 128 rows are tiled (output tile x source chunk); entries must stay inside their block.
#include <cuda_bf16.h>
#include "common.cuh"

namespace fitgnn {
namespace {

constexpr int MM_THREADS = 256;
constexpr int MM_TILE = 128;                   // rows of an M piece (output rows) and source rows per chunk (MMA K extent)
constexpr int MM_MP = MM_TILE + 8;             // M row pitch (bf16): 272 bytes = 17 x 16 -> conflict-free ldmatrix
constexpr size_t MM_BUDGET = 112 * 1024;       // shared memory per CTA: two CTAs per SM

__host__ __device__ constexpr int mm_yp(int nt8) { return 8 * nt8 + 8; }  // Yt row pitch (bf16), conflict-free ldmatrix
__host__ __device__ constexpr size_t mm_fixed(int nt8) {
  return (size_t)MM_TILE * MM_MP * 2 + 2 * (size_t)MM_TILE * mm_yp(nt8) * 2 + 3 * MM_TILE * 4 + (MM_TILE + 4) * 4;
}
__host__ __device__ constexpr int mm_cs_cap(int nt8) { return (int)((MM_BUDGET - mm_fixed(nt8)) / 2) & ~7; }

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ float elu1(float x) { return elu_fast(x); }

// NT8 = n8 tiles per slice (feature columns = 8 * NT8 <= 128)
template <int NT8, bool SPLIT>
__global__ void __launch_bounds__(MM_THREADS, 2)
spmm_mma_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col, const float* __restrict__ dinv,
                const float* __restrict__ X, int64_t ldx, int nq, const int32_t* __restrict__ src_index,
                const int32_t* __restrict__ blk_ptr, int64_t n_blk, int n_slices, const float* __restrict__ bias, int act,
                void* Y, void* Ylo, int64_t ldy) {
  constexpr int YP = mm_yp(NT8);
  constexpr int CS_CAP = mm_cs_cap(NT8);
  extern __shared__ __align__(16) unsigned char mm_smem[];
  __nv_bfloat16* Ms = reinterpret_cast<__nv_bfloat16*>(mm_smem);  // [128][MM_MP]  M piece
  __nv_bfloat16* Yh = Ms + MM_TILE * MM_MP;                       // [128][YP]     D X hi
  __nv_bfloat16* Yl = Yh + MM_TILE * YP;                          // [128][YP]     D X lo
  float* dsm = reinterpret_cast<float*>(Yl + MM_TILE * YP);       // [128] dinv of the output tile's rows
  float* dks = dsm + MM_TILE;                                     // [128] dinv of the chunk's source rows
  int32_t* srw = reinterpret_cast<int32_t*>(dks + MM_TILE);       // [128] feature-table row of the chunk's source rows
  int32_t* rps = srw + MM_TILE;                                   // [129] CSR ranges of the tile's rows (relative to its first entry)
  uint16_t* cs = reinterpret_cast<uint16_t*>(rps + MM_TILE + 4);  // [CS_CAP] block-local column indices of the tile's rows
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int NQS = 2 * NT8;  // float4 columns per slice
  const int64_t items = n_blk * n_slices;
  for (int64_t it = blockIdx.x; it < items; it += gridDim.x) {
    const int64_t b = it / n_slices;
    const int sl = (int)(it % n_slices);
    const int base = __ldg(blk_ptr + b);
    const int rows_b = __ldg(blk_ptr + b + 1) - base;
    if (rows_b <= 0) continue;  // CTA-uniform
    const int q0 = sl * NQS;            // first float4 column of the slice
    const int nqs = min(NQS, nq - q0);  // float4 columns that exist
    for (int i0 = 0; i0 < rows_b; i0 += MM_TILE) {  // output tile
      const int ti = min(MM_TILE, rows_b - i0);
      const int ebase = __ldg(rowptr + base + i0);
      const int n_ent = __ldg(rowptr + base + i0 + ti) - ebase;
      const bool cs_ok = n_ent <= CS_CAP && rows_b <= 65535;  // CTA-uniform: the tile's column indices fit the staging area
      float acc[NT8][4];
#pragma unroll
      for (int t = 0; t < NT8; ++t) acc[t][0] = acc[t][1] = acc[t][2] = acc[t][3] = 0.f;
      for (int j0 = 0; j0 < rows_b; j0 += MM_TILE) {  // source chunk
        const int tj = min(MM_TILE, rows_b - j0);
        const int k_ext = (tj + 15) & ~15;  // MMA K extent of this chunk
        __syncthreads();                    // the previous piece's MMAs / epilogue are done with the shared arrays
        // 1. zero the M piece; fetch everything index-like with coalesced loads: the chunk's source rows and weights, and (once
        //    per output tile) its rows' dinv, CSR ranges and block-local column indices
        for (int e = tid; e < ti * (k_ext / 8); e += MM_THREADS) {
          const int r = e / (k_ext / 8), c8 = e % (k_ext / 8);
          *reinterpret_cast<uint4*>(Ms + r * MM_MP + 8 * c8) = make_uint4(0u, 0u, 0u, 0u);
        }
        for (int k = tid; k < tj; k += MM_THREADS) {
          const int r = base + j0 + k;
          srw[k] = src_index ? __ldg(src_index + r) : r;
          dks[k] = __ldg(dinv + r);
        }
        if (j0 == 0) {
          for (int i = tid; i < ti; i += MM_THREADS) dsm[i] = __ldg(dinv + base + i0 + i);
          for (int i = tid; i <= ti; i += MM_THREADS) rps[i] = __ldg(rowptr + base + i0 + i) - ebase;
          if (cs_ok)
            for (int e = tid; e < n_ent; e += MM_THREADS) cs[e] = (uint16_t)(__ldg(col + ebase + e) - base);
        }
        __syncthreads();
        // 2. (a) issue the global loads of the first two source rows this thread stages, (b) scatter the CSR entries while they
        //    are in flight, (c) convert and store the staged rows, then the remaining ones.
        const int sub = tid & 7, grp = tid >> 3;  // staging: 8 lanes per source row, 32 rows per pass
        constexpr int NV = (NQS + 7) / 8;
        constexpr int RPP = MM_THREADS / 8;
        auto load_rows = [&](int k, float4 (&x)[2][NV], float (&dk)[2]) {
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const int kk = k + u * RPP;
            const bool rowok = kk < tj;
            dk[u] = rowok ? dks[kk] : 0.f;
            const float4* xr = reinterpret_cast<const float4*>(X + (int64_t)(rowok ? srw[kk] : 0) * ldx) + q0;
#pragma unroll
            for (int v = 0; v < NV; ++v) {
              const int q = sub + 8 * v;
              x[u][v] = (rowok && q < nqs) ? __ldg(xr + q) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
          }
        };
        auto store_rows = [&](int k, const float4 (&x)[2][NV], const float (&dk)[2]) {
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const int kk = k + u * RPP;
            if (kk < k_ext) {  // rows past the chunk up to the K extent are zero (they meet zero columns of M, but must be finite)
#pragma unroll
              for (int v = 0; v < NV; ++v) {
                const int q = sub + 8 * v;
                if (q < NQS) {
                  uint2 h, l;
                  split_bf16x2(x[u][v].x * dk[u], x[u][v].y * dk[u], h.x, l.x);
                  split_bf16x2(x[u][v].z * dk[u], x[u][v].w * dk[u], h.y, l.y);
                  *reinterpret_cast<uint2*>(Yh + kk * YP + 4 * q) = h;
                  *reinterpret_cast<uint2*>(Yl + kk * YP + 4 * q) = l;
                }
              }
            }
          }
        };
        float4 x0[2][NV];
        float dk0[2];
        load_rows(grp, x0, dk0);
        // (b) scatter: one WARP per row, lanes over its entries (coalesced u16 reads from the staged indices).  Entries ascend, so
        //     duplicate edges (which PyG counts) are adjacent: the first lane of a run stores the run length; a run that crosses a
        //     32-entry step is continued by a read-modify-write of the same element (same warp: no race).
        if (cs_ok) {
          const int lo = j0, hi = j0 + tj;  // block-local column range of the chunk
          for (int i = warp; i < ti; i += MM_THREADS / 32) {
            const int beg = rps[i], end = rps[i + 1];
            __nv_bfloat16* mrow = Ms + i * MM_MP;
            int carry_c = -1;
            for (int e0 = beg; e0 < end; e0 += 32) {
              const int e = e0 + lane;
              const int c = e < end ? (int)cs[e] : 0x7fffffff;
              const bool in = c >= lo && c < hi;
              int prev = __shfl_up_sync(0xffffffffu, c, 1);
              if (lane == 0) prev = carry_c;
              const bool start = in && c != prev;
              const unsigned starts = __ballot_sync(0xffffffffu, c != prev);  // run starts of ANY column (ends of my run)
              if (in) {
                const unsigned later = lane == 31 ? 0u : (starts >> (lane + 1));
                const int run = later ? __ffs(later) : (32 - lane);  // entries of my run inside this step (from me on)
                if (start) {
                  mrow[c - lo] = __float2bfloat16_rn((float)run);
                } else if (lane == 0) {  // my run began in the previous step: add this step's part
                  mrow[c - lo] = __float2bfloat16_rn(__bfloat162float(mrow[c - lo]) + (float)run);
                }
              }
              carry_c = __shfl_sync(0xffffffffu, c, 31);
            }
          }
        } else if (tid < ti) {  // oversized tile: one thread per row on global memory
          const int beg = rps[tid], end = rps[tid + 1];
          const int lo = j0, hi = j0 + tj;
          __nv_bfloat16* mrow = Ms + tid * MM_MP;
          const int32_t* cg = col + ebase;
          int e = beg;
          while (e < end && __ldg(cg + e) - base < lo) ++e;
          while (e < end) {
            const int c = __ldg(cg + e) - base;
            if (c >= hi) break;
            int cnt = 1;
            while (e + cnt < end && __ldg(cg + e + cnt) - base == c) ++cnt;
            mrow[c - lo] = __float2bfloat16_rn((float)cnt);
            e += cnt;
          }
        }
        // (c)
        store_rows(grp, x0, dk0);
        for (int k = grp + 2 * RPP; k < k_ext; k += 2 * RPP) {
          float4 x1[2][NV];
          float dk1[2];
          load_rows(k, x1, dk1);
          store_rows(k, x1, dk1);
        }
        __syncthreads();
        // 3. MMA: warp w owns output rows [16w, 16w + 16) of the tile
        if (16 * warp < ti) {
          const uint32_t a_base = smem_u32(Ms + (16 * warp + (lane & 7) + ((lane >> 3) & 1) * 8) * MM_MP + (lane >> 4) * 8);
          const uint32_t bh_base = smem_u32(Yh + ((lane & 7) + ((lane >> 3) & 1) * 8) * YP + (lane >> 4) * 8);
          const uint32_t bl_base = smem_u32(Yl + ((lane & 7) + ((lane >> 3) & 1) * 8) * YP + (lane >> 4) * 8);
          for (int k0 = 0; k0 < k_ext; k0 += 16) {
            uint32_t a[4];
            ldmatrix_x4(a, a_base + (uint32_t)k0 * 2u);
#pragma unroll
            for (int t = 0; t < NT8; t += 2) {
              uint32_t bh[4], bl[4];
              const uint32_t off = ((uint32_t)k0 * YP + 8u * t) * 2u;
              ldmatrix_x4_trans(bh, bh_base + off);
              ldmatrix_x4_trans(bl, bl_base + off);
              mma_bf16(acc[t], a, bh[0], bh[1]);
              mma_bf16(acc[t], a, bl[0], bl[1]);
              if (t + 1 < NT8) {
                mma_bf16(acc[t + 1], a, bh[2], bh[3]);
                mma_bf16(acc[t + 1], a, bl[2], bl[3]);
              }
            }
          }
        }
      }
      // 4. epilogue: c0,c1 -> row lane/4, cols 2*(lane%4), +1 of the n8 tile; c2,c3 -> row + 8
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int i = 16 * warp + (lane >> 2) + 8 * half;
        if (i < ti) {
          const float dr = dsm[i];
          const int64_t yrow = (int64_t)(base + i0 + i) * ldy;
#pragma unroll
          for (int t = 0; t < NT8; ++t) {
            const int cidx = 8 * t + 2 * (lane & 3);  // column inside the slice
            if (cidx < 4 * nqs) {
              const int cg = 4 * q0 + cidx;
              float o0 = acc[t][2 * half] * dr, o1 = acc[t][2 * half + 1] * dr;
              if (bias) { o0 += __ldg(bias + cg); o1 += __ldg(bias + cg + 1); }
              if (act == FITGNN_ACT_ELU) { o0 = elu1(o0); o1 = elu1(o1); }
              if (SPLIT) {
                uint32_t h, l;
                split_bf16x2(o0, o1, h, l);
                *reinterpret_cast<uint32_t*>(static_cast<__nv_bfloat16*>(Y) + yrow + cg) = h;
                *reinterpret_cast<uint32_t*>(static_cast<__nv_bfloat16*>(Ylo) + yrow + cg) = l;
              } else {
                *reinterpret_cast<float2*>(static_cast<float*>(Y) + yrow + cg) = make_float2(o0, o1);
              }
            }
          }
        }
      }
    }
  }
}

}  // namespace
}  // namespace fitgnn

using namespace fitgnn;

extern "C" int fitgnn_spmm_symnorm_mma(const int32_t* rowptr, const int32_t* col, const float* dinv, const float* X, int64_t ldx,
                                       int width, const int32_t* src_index, const int32_t* blk_ptr, int64_t n_blk,
                                       const float* bias, int act, void* Y, void* Y_lo, int64_t ldy, void* stream) {
  FG_REQUIRE(rowptr && col && dinv && X && Y && blk_ptr, FITGNN_EINVAL, "spmm_mma: null pointer");
  FG_REQUIRE(n_blk >= 0 && width > 0, FITGNN_EINVAL, "spmm_mma: n_blk=%lld width=%d", (long long)n_blk, width);
  FG_REQUIRE(width % 4 == 0 && ldx % 4 == 0 && ldy % 4 == 0, FITGNN_EUNSUP,
             "spmm_mma: width (%d), ldx (%lld), ldy (%lld) must be multiples of 4", width, (long long)ldx, (long long)ldy);
  FG_REQUIRE(((uintptr_t)X % 16) == 0 && ((uintptr_t)Y % 8) == 0 && ((uintptr_t)Y_lo % 4) == 0 && (!bias || ((uintptr_t)bias % 8) == 0),
             FITGNN_EUNSUP, "spmm_mma: X must be 16-byte, Y / bias 8-byte aligned");
  FG_REQUIRE(act == FITGNN_ACT_NONE || act == FITGNN_ACT_ELU, FITGNN_EINVAL, "spmm_mma: unknown act %d", act);
  if (n_blk == 0) return FITGNN_OK;
  cudaStream_t st = as_stream(stream);
  const int nq = width / 4;
  // slices of at most 128 feature columns; narrow rows get the smallest instantiation that covers them
  const int nt8 = nq <= 8 ? 4 : nq <= 16 ? 8 : nq <= 28 ? 14 : 16;
  const int n_slices = (int)ceil_div(nq, 2 * nt8);
  const int64_t items = n_blk * n_slices;
  const int64_t max_blocks = (int64_t)sm_count() * 2;
  const unsigned blocks = (unsigned)(items < max_blocks ? items : max_blocks);
  const bool split = Y_lo != nullptr;
#define FG_MM(NT, SP)                                                                                                      \
  do {                                                                                                                     \
    FG_CUDA(cudaFuncSetAttribute(spmm_mma_kernel<NT, SP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MM_BUDGET));   \
    spmm_mma_kernel<NT, SP><<<blocks, MM_THREADS, MM_BUDGET, st>>>(rowptr, col, dinv, X, ldx, nq, src_index, blk_ptr, n_blk, \
                                                                 n_slices, bias, act, Y, Y_lo, ldy);                      \
  } while (0)
  if (nt8 == 4) { if (split) FG_MM(4, true); else FG_MM(4, false); }
  else if (nt8 == 8) { if (split) FG_MM(8, true); else FG_MM(8, false); }
  else if (nt8 == 14) { if (split) FG_MM(14, true); else FG_MM(14, false); }
  else { if (split) FG_MM(16, true); else FG_MM(16, false); }
#undef FG_MM
  FG_LAUNCH_CHECK();
  return FITGNN_OK;
}
