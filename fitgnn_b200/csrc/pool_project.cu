// Segment pooling (graph-level heads) and the coarsened-graph projections.
//   fitgnn_segment_pool      <- x[mask] + torch.cat + global_max_pool / global_mean_pool
//                               (/root/reference/network.py:129-131, :200-202, :93, :164)
//   fitgnn_project_features  <- C.dot(H_feature) (/root/reference/utils.py:161, :738, :827), C built by
//                               get_coarsening_matrix (graph_coarsening/coarsening_utils.py:212-254, :136)
//   fitgnn_project_adj_*     <- zero_diag(coarsen_matrix(W, iC)) (coarsening_utils.py:138, :201-205;
//                               graph_utils.py:79-87) consumed as Gc.W.tocoo() (utils.py:745-746)
//   fitgnn_group_by_part     <- metanode_to_node_mapping_new (utils.py:123-130)
// All HBM-bound: one pass over the inputs, segmented (atomic-free) reductions over sorted segments.
#include <math.h>
#include "common.cuh"

namespace fitgnn {

// ---------------------------------------------------------------------------------------------
// pooling: one warp per (segment, 128-column block)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
segment_pool_kernel(const float* __restrict__ X, int64_t ldx, int nq, const int32_t* __restrict__ rows,
                    const int32_t* __restrict__ seg_ptr, int64_t n_seg, int pool, float* __restrict__ Y,
                    int64_t ldy) {
  const int lane = threadIdx.x & 31;
  const int64_t g = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (g >= n_seg) return;
  const int q = blockIdx.y * 32 + lane;
  if (q >= nq) return;
  const int beg = seg_ptr[g], end = seg_ptr[g + 1];
  float4 acc = pool == FITGNN_POOL_MAX ? make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY)
                                       : make_float4(0.f, 0.f, 0.f, 0.f);
  for (int i = beg; i < end; ++i) {
    const int64_t r = rows ? rows[i] : i;
    const float4 v = __ldg(reinterpret_cast<const float4*>(X + r * ldx) + q);
    if (pool == FITGNN_POOL_MAX) {
      acc.x = fmaxf(acc.x, v.x); acc.y = fmaxf(acc.y, v.y); acc.z = fmaxf(acc.z, v.z); acc.w = fmaxf(acc.w, v.w);
    } else {
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
  }
  if (end == beg) {
    acc = make_float4(0.f, 0.f, 0.f, 0.f);
  } else if (pool == FITGNN_POOL_MEAN) {
    const float inv = 1.f / (float)(end - beg);  // PyG: sum / clamp(count, 1)
    acc.x *= inv; acc.y *= inv; acc.z *= inv; acc.w *= inv;
  }
  *(reinterpret_cast<float4*>(Y + g * ldy) + q) = acc;
}

// ---------------------------------------------------------------------------------------------
// Xc = C·X : one warp per (cluster, 128-column block); fp64 accumulate in ascending member id,
// separate multiply and add (no FMA contraction) so the result rounds like scipy's csc_matvecs.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
project_features_kernel(const int32_t* __restrict__ members, const int32_t* __restrict__ member_ptr, int64_t k,
                        const double* __restrict__ cweight, const float* __restrict__ X, int64_t ldx, int F,
                        float* __restrict__ Xc, int64_t ldxc) {
  const int lane = threadIdx.x & 31;
  const int64_t c = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (c >= k) return;
  const int col0 = (blockIdx.y * 32 + lane) * 4;
  if (col0 >= F) return;
  const int beg = member_ptr[c], end = member_ptr[c + 1];
  const bool vec = (col0 + 3 < F) && (ldx % 4 == 0) && (((uintptr_t)X & 15) == 0);
  double acc[4] = {0.0, 0.0, 0.0, 0.0};
  for (int i = beg; i < end; ++i) {
    const int j = members[i];
    const double w = cweight[j];
    float x[4];
    if (vec) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(X + (int64_t)j * ldx + col0));
      x[0] = v.x; x[1] = v.y; x[2] = v.z; x[3] = v.w;
    } else {
#pragma unroll
      for (int u = 0; u < 4; ++u) x[u] = (col0 + u < F) ? __ldg(X + (int64_t)j * ldx + col0 + u) : 0.f;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) acc[u] = __dadd_rn(acc[u], __dmul_rn(w, (double)x[u]));
  }
#pragma unroll
  for (int u = 0; u < 4; ++u)
    if (col0 + u < F) Xc[c * ldxc + col0 + u] = (float)acc[u];
}

// ---------------------------------------------------------------------------------------------
// group nodes by part: keys = part<<32 | id, sort, split
// ---------------------------------------------------------------------------------------------
__global__ void part_keys_kernel(const int32_t* __restrict__ part, int64_t N, int64_t k, uint64_t* __restrict__ keys,
                                 int32_t* err) {
  const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= N) return;
  const int32_t p = part[v];
  if (p < 0 || p >= k) {
    atomicExch(err, 1);
    keys[v] = ((uint64_t)k << 32) | (uint64_t)v;
    return;
  }
  keys[v] = ((uint64_t)p << 32) | (uint64_t)v;
}

__global__ void low32_kernel(const uint64_t* __restrict__ keys, int64_t n, int32_t* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (int32_t)(keys[i] & 0xffffffffull);
}

// ---------------------------------------------------------------------------------------------
// Ac = P_bin·A·P_bin^T minus diagonal: relabel edges to (part[src], part[dst]) keys, sort, run-length
// ---------------------------------------------------------------------------------------------
// ADJ_EPT edges per thread, the part[] gathers of all of them in flight together; one counter atomic per warp
constexpr int ADJ_EPT = 4;
__global__ void __launch_bounds__(256)
adj_keys_kernel(const int64_t* __restrict__ edge_index, int64_t E, int64_t N, const int32_t* __restrict__ part, int64_t k,
                int kb, uint64_t* __restrict__ keys, int32_t* __restrict__ n_drop) {
  const int64_t base = (int64_t)blockIdx.x * (256 * ADJ_EPT) + threadIdx.x;
  int64_t u[ADJ_EPT], v[ADJ_EPT];
#pragma unroll
  for (int j = 0; j < ADJ_EPT; ++j) {
    const int64_t e = base + (int64_t)j * 256;
    u[j] = e < E ? edge_index[e] : 0;
    v[j] = e < E ? edge_index[E + e] : 0;
  }
  int32_t a[ADJ_EPT], b[ADJ_EPT];
  bool bad = false;
#pragma unroll
  for (int j = 0; j < ADJ_EPT; ++j) {
    const bool live = base + (int64_t)j * 256 < E;
    const bool inr = live && u[j] >= 0 && u[j] < N && v[j] >= 0 && v[j] < N;
    bad |= live && !inr;
    a[j] = inr ? __ldg(part + u[j]) : 0;
    b[j] = inr ? __ldg(part + v[j]) : 0;
  }
  int dropped = 0;
#pragma unroll
  for (int j = 0; j < ADJ_EPT; ++j) {
    const int64_t e = base + (int64_t)j * 256;
    if (e < E) {
      if (a[j] < 0 || a[j] >= k || b[j] < 0 || b[j] >= k) {
        bad = true;
        a[j] = b[j] = 0;
      }
      // compact key (row << kb | col), kb = bits of k: 2·kb significant bits instead of 32 + kb -> fewer sort passes;
      // intra-cluster edges / self loops -> k << kb, which sorts after every real key
      const bool drop = a[j] == b[j];
      keys[e] = drop ? (uint64_t)k << kb : ((uint64_t)a[j] << kb) | (uint64_t)b[j];
      dropped += drop ? 1 : 0;
    }
  }
  if (bad) atomicExch(n_drop + 1, 1);
  dropped = __reduce_add_sync(0xffffffffu, dropped);
  if ((threadIdx.x & 31) == 0 && dropped) atomicAdd(n_drop, dropped);
}

// pos = exclusive scan of flags; run starts write their (row, col) and the run length
__global__ void adj_emit_kernel(const uint64_t* __restrict__ keys, const int32_t* __restrict__ pos, int64_t n, int kb,
                                int64_t* __restrict__ out_row, int64_t* __restrict__ out_col,
                                int32_t* __restrict__ run_start) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (i == 0 || keys[i] != keys[i - 1]) {
    const int32_t p = pos[i];
    out_row[p] = (int64_t)(keys[i] >> kb);
    out_col[p] = (int64_t)(keys[i] & ((1ull << kb) - 1ull));
    run_start[p] = (int32_t)i;
  }
}

__global__ void adj_count_kernel(const int32_t* __restrict__ run_start, int64_t nnz, int64_t n_valid,
                                 int32_t* __restrict__ out_cnt) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= nnz) return;
  const int32_t next = (p + 1 < nnz) ? run_start[p + 1] : (int32_t)n_valid;
  out_cnt[p] = next - run_start[p];
}

__global__ void rowptr_from_rows_kernel(const int64_t* __restrict__ rows, int64_t nnz, int64_t k,
                                        int32_t* __restrict__ rowptr) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i > nnz) return;
  const int64_t prev = (i == 0) ? -1 : rows[i - 1];
  const int64_t cur = (i == nnz) ? k : rows[i];
  for (int64_t r = prev + 1; r <= cur; ++r) rowptr[r] = (int32_t)i;
}

}  // namespace fitgnn

using namespace fitgnn;

extern "C" int fitgnn_segment_pool(const float* X, int64_t ldx, int width, const int32_t* rows, const int32_t* seg_ptr,
                                   int64_t n_seg, int pool, float* Y, int64_t ldy, void* stream) {
  FG_REQUIRE(X && seg_ptr && Y && n_seg >= 0 && width > 0, FITGNN_EINVAL, "segment_pool: bad arguments");
  FG_REQUIRE(pool == FITGNN_POOL_MAX || pool == FITGNN_POOL_MEAN, FITGNN_EINVAL, "segment_pool: unknown pool %d", pool);
  FG_REQUIRE(width % 4 == 0 && ldx % 4 == 0 && ldy % 4 == 0, FITGNN_EUNSUP,
             "segment_pool: width/ldx/ldy must be multiples of 4");
  if (n_seg == 0) return FITGNN_OK;
  const int nq = width / 4;
  dim3 grid((unsigned)ceil_div(n_seg, 8), (unsigned)ceil_div(nq, 32));
  segment_pool_kernel<<<grid, 256, 0, as_stream(stream)>>>(X, ldx, nq, rows, seg_ptr, n_seg, pool, Y, ldy);
  FG_LAUNCH_CHECK();
  return FITGNN_OK;
}

extern "C" size_t fitgnn_group_workspace_bytes(int64_t N, int64_t k) {
  (void)k;
  return align_up((size_t)(N > 0 ? N : 1) * 8) + sort_ws_bytes(N) + 512;
}

extern "C" int fitgnn_group_by_part(const int32_t* part, int64_t N, int64_t k, int32_t* members, int32_t* member_ptr,
                                    void* ws, size_t ws_bytes, void* stream) {
  FG_REQUIRE(part && members && member_ptr && N >= 0 && k >= 0, FITGNN_EINVAL, "group_by_part: bad arguments");
  FG_REQUIRE(N < (1ll << 31) && k < (1ll << 31), FITGNN_ERANGE, "group_by_part: N or k exceeds int32");
  cudaStream_t st = as_stream(stream);
  Bump b(ws, ws_bytes);
  uint64_t* keys = b.take<uint64_t>((size_t)(N > 0 ? N : 1));
  int32_t* err = b.take<int32_t>(1);
  FG_REQUIRE(b.ok, FITGNN_EWS, "group_by_part: workspace too small");
  FG_CUDA(cudaMemsetAsync(err, 0, sizeof(int32_t), st));
  if (N > 0) {
    part_keys_kernel<<<(unsigned)ceil_div(N, 256), 256, 0, st>>>(part, N, k, keys, err);
    FG_LAUNCH_CHECK();
    // the id field already ascends with the index and the sort is stable: only the part field needs passes
    FG_TRY(sort_u64_mask(keys, nullptr, N, field_mask((uint64_t)k, 0) & ~0xffffffffull, b.here(), b.left(), st));
    low32_kernel<<<(unsigned)ceil_div(N, 256), 256, 0, st>>>(keys, N, members);
    FG_LAUNCH_CHECK();
  }
  FG_TRY(segment_ptr_from_sorted(keys, N, 32, k, member_ptr, st));
  int32_t herr = 0;
  FG_CUDA(cudaMemcpyAsync(&herr, err, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  FG_CUDA(cudaStreamSynchronize(st));
  FG_REQUIRE(herr == 0, FITGNN_EINVAL, "group_by_part: part[] holds ids outside [0,%lld)", (long long)k);
  return FITGNN_OK;
}

extern "C" int fitgnn_project_features(const int32_t* members, const int32_t* member_ptr, int64_t k,
                                       const double* cweight, const float* X, int64_t ldx, int F, float* Xc,
                                       int64_t ldxc, void* stream) {
  FG_REQUIRE(members && member_ptr && cweight && X && Xc && k >= 0 && F > 0, FITGNN_EINVAL,
             "project_features: bad arguments");
  if (k == 0) return FITGNN_OK;
  dim3 grid((unsigned)ceil_div(k, 8), (unsigned)ceil_div(ceil_div(F, 4), 32));
  project_features_kernel<<<grid, 256, 0, as_stream(stream)>>>(members, member_ptr, k, cweight, X, ldx, F, Xc, ldxc);
  FG_LAUNCH_CHECK();
  return FITGNN_OK;
}

// plan header kept at the start of the workspace (host mirror returned through host_nnz only)
struct AdjPlan {
  int64_t E, n_valid, nnz;
};

extern "C" size_t fitgnn_project_adj_workspace_bytes(int64_t E) {
  const size_t e = (size_t)(E > 0 ? E : 1);
  return 512 + align_up(e * 8) + 2 * align_up((e + 1) * 4) + scan_ws_bytes(E + 1) + sort_ws_bytes(E) + 1024;
}

extern "C" int fitgnn_project_adj_plan(const int64_t* edge_index, int64_t E, int64_t N, const int32_t* part, int64_t k,
                                       void* ws, size_t ws_bytes, int64_t* host_nnz, void* stream) {
  FG_REQUIRE(E >= 0 && k >= 0 && host_nnz && ws && (E == 0 || (edge_index && part)), FITGNN_EINVAL,
             "project_adj_plan: bad arguments");
  FG_REQUIRE(E < (1ll << 31) - 8192 && k < (1ll << 31), FITGNN_ERANGE, "project_adj_plan: E or k exceeds int32");
  cudaStream_t st = as_stream(stream);
  Bump b(ws, ws_bytes);
  int64_t* hdr = b.take<int64_t>(8);
  int32_t* counter = b.take<int32_t>(4);
  const size_t e = (size_t)(E > 0 ? E : 1);
  uint64_t* keys = b.take<uint64_t>(e);
  int32_t* flags = b.take<int32_t>(e + 1);
  int32_t* run_start = b.take<int32_t>(e + 1);
  (void)run_start;
  FG_REQUIRE(b.ok, FITGNN_EWS, "project_adj_plan: workspace too small");
  FG_CUDA(cudaMemsetAsync(counter, 0, 4 * sizeof(int32_t), st));
  int64_t n_valid = 0, nnz = 0;
  const int kb = bits_for((uint64_t)k);  // k < 2^31 -> 2·kb <= 62
  if (E > 0) {
    adj_keys_kernel<<<(unsigned)ceil_div(E, 256 * ADJ_EPT), 256, 0, st>>>(edge_index, E, N, part, k, kb, keys, counter);
    FG_LAUNCH_CHECK();
    FG_TRY(sort_u64(keys, nullptr, E, 2 * kb, b.here(), b.left(), st));
    int32_t hc[2] = {0, 0};
    FG_CUDA(cudaMemcpyAsync(hc, counter, sizeof(hc), cudaMemcpyDeviceToHost, st));
    FG_CUDA(cudaStreamSynchronize(st));
    FG_REQUIRE(hc[1] == 0, FITGNN_EINVAL, "project_adj_plan: edge_index / part hold ids out of range");
    n_valid = E - hc[0];
    if (n_valid > 0) {
      // exclusive scan of the run-boundary flags, formed on the fly from the sorted keys
      FG_TRY(scan_key_boundaries(keys, n_valid, flags, n_valid + 1, b.here(), b.left(), st));
      int32_t total = 0;
      FG_CUDA(cudaMemcpyAsync(&total, flags + n_valid, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
      FG_CUDA(cudaStreamSynchronize(st));
      nnz = total;
    }
  }
  const int64_t h[8] = {E, n_valid, nnz, kb, 0, 0, 0, 0};
  FG_CUDA(cudaMemcpyAsync(hdr, h, sizeof(h), cudaMemcpyHostToDevice, st));
  FG_CUDA(cudaStreamSynchronize(st));
  *host_nnz = nnz;
  return FITGNN_OK;
}

extern "C" int fitgnn_project_adj_fill(void* ws, size_t ws_bytes, int64_t k, int64_t* out_row, int64_t* out_col,
                                       int32_t* out_cnt, int32_t* out_rowptr, void* stream) {
  FG_REQUIRE(ws && k >= 0, FITGNN_EINVAL, "project_adj_fill: bad arguments");
  cudaStream_t st = as_stream(stream);
  Bump b(ws, ws_bytes);
  int64_t* hdr = b.take<int64_t>(8);
  b.take<int32_t>(4);
  int64_t h[8];
  FG_CUDA(cudaMemcpyAsync(h, hdr, sizeof(h), cudaMemcpyDeviceToHost, st));
  FG_CUDA(cudaStreamSynchronize(st));
  const int64_t E = h[0], n_valid = h[1], nnz = h[2];
  const int kb = (int)h[3];
  const size_t e = (size_t)(E > 0 ? E : 1);
  uint64_t* keys = b.take<uint64_t>(e);
  int32_t* pos = b.take<int32_t>(e + 1);
  int32_t* run_start = b.take<int32_t>(e + 1);
  FG_REQUIRE(b.ok, FITGNN_EWS, "project_adj_fill: workspace too small");
  if (nnz > 0) {
    FG_REQUIRE(out_row && out_col && out_cnt, FITGNN_EINVAL, "project_adj_fill: null output");
    adj_emit_kernel<<<(unsigned)ceil_div(n_valid, 256), 256, 0, st>>>(keys, pos, n_valid, kb, out_row, out_col, run_start);
    FG_LAUNCH_CHECK();
    adj_count_kernel<<<(unsigned)ceil_div(nnz, 256), 256, 0, st>>>(run_start, nnz, n_valid, out_cnt);
    FG_LAUNCH_CHECK();
  }
  if (out_rowptr) {
    rowptr_from_rows_kernel<<<(unsigned)ceil_div(nnz + 1, 256), 256, 0, st>>>(out_row, nnz, k, out_rowptr);
    FG_LAUNCH_CHECK();
  }
  return FITGNN_OK;
}
