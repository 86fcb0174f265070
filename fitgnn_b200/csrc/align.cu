// Row-group alignment of a pack: a derived layout in which no subgraph straddles a multiple of `group` rows, so
// that every CSR entry of a row points into the row's own aligned group of 32 rows.  That makes the normalised
// aggregation Â·H a warp-local operation inside the tensor-core GEMM epilogue (thread = row, one TMEM lane quadrant
// = one group; gemm_tcgen05.cu, fitgnn_gcn_transform_aggregate) — the layer order "GCNConv -> ELU -> next GCNConv's
// propagate" of /root/reference/network.py:31-33 then needs no separate SpMM launch and no HBM round trip.
//
// Two placement policies (FITGNN_ALIGN_*):
//   IN_ORDER  subgraphs keep the pack's (= the reference's subgraph_list) order and are placed greedily; when the
//             next one does not fit into what is left of the current group, the group is closed with padding rows.
//   BY_DEGREE subgraphs are placed in order of (largest non-self row degree desc, size desc, index asc).  The fused
//             aggregation loops to the LARGEST row degree of a warp's 32 rows, so grouping similar degrees halves that
//             loop (3.9 -> 2.1 on the products-shaped pack), and the gap left at a group's tail is filled from the
//             other end of the order (singletons and other low-degree subgraphs, which cannot raise the group's
//             maximum) instead of with padding.
// Padding rows are empty CSR rows with dinv = 0 and orig_row = -1.
//
// Per aligned row the fill also emits an aggregation descriptor: bits [0,4) = number c of non-self entries
// (c <= 12), bits [4+5j, 9+5j) = lane (row index inside the group) of the j-th one, duplicates kept.
#include <vector>
#include "common.cuh"

namespace fitgnn {

constexpr int ALIGN_MAX_INLINE = 12;

// sort key of a subgraph for FITGNN_ALIGN_BY_DEGREE (ascending sort = degree desc, size desc, index asc)
__global__ void align_keys_kernel(fitgnn_pack in, int group, uint64_t* __restrict__ keys) {
  const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= in.n_sub) return;
  const int a = in.sub_ptr[s], b = in.sub_ptr[s + 1];
  int maxdeg = 0;
  for (int r = a; r < b; ++r) maxdeg = max(maxdeg, in.rowptr[r + 1] - in.rowptr[r] - 1);
  maxdeg = min(max(maxdeg, 0), 63);
  const int size = min(b - a, 1023);
  keys[s] = ((uint64_t)(63 - maxdeg) << 42) | ((uint64_t)(1023 - size) << 32) | (uint64_t)s;
}

__global__ void align_sizes_kernel(const int32_t* __restrict__ sub_ptr, const uint64_t* __restrict__ keys, int64_t n_sub,
                                   int32_t* __restrict__ order, int32_t* __restrict__ size_sorted) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_sub) return;
  const int32_t s = keys ? (int32_t)(keys[i] & 0xffffffffull) : (int32_t)i;
  order[i] = s;
  size_sorted[i] = sub_ptr[s + 1] - sub_ptr[s];
}

__global__ void align_init_rows_kernel(int64_t n_al, int32_t* deg_a, float* dinv_a, int32_t* gid_a, uint8_t* is_core_a,
                                       uint8_t* mask_a, int32_t* orig_row, unsigned long long* agg_desc) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_al) return;
  deg_a[r] = 0;
  dinv_a[r] = 0.f;
  gid_a[r] = 0;
  is_core_a[r] = 0;
  mask_a[r] = 0;
  orig_row[r] = -1;
  agg_desc[r] = 0ull;
}

// one thread per subgraph: moves its rows (everything but the CSR entries, which need the scanned row pointers)
__global__ void align_rows_kernel(fitgnn_pack in, const int32_t* __restrict__ new_start, int group, int32_t* deg_a,
                                  float* dinv_a, int32_t* gid_a, uint8_t* is_core_a, uint8_t* mask_a, int32_t* orig_row,
                                  int32_t* new_of_old, unsigned long long* agg_desc, int32_t* __restrict__ flags) {
  const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= in.n_sub) return;
  const int a = in.sub_ptr[s], b = in.sub_ptr[s + 1];
  const int ns = new_start[s];
  const int shift = ns - a;
  const int gbase = ns & ~(group - 1);
  for (int r = a; r < b; ++r) {
    const int nr = r + shift;
    const int e0 = in.rowptr[r], e1 = in.rowptr[r + 1];
    new_of_old[r] = nr;
    orig_row[nr] = r;
    deg_a[nr] = e1 - e0;
    dinv_a[nr] = in.dinv[r];
    gid_a[nr] = in.gid[r];
    is_core_a[nr] = in.is_core[r];
    mask_a[nr] = in.mask[r];
    unsigned long long d = 0;
    int cnt = 0;
    bool self_seen = false;
    for (int e = e0; e < e1; ++e) {
      const int c = in.col[e];
      if (c < a || c >= b) atomicOr(flags, 4);  // entry leaves its subgraph: not a block-diagonal pack
      if (c == r && !self_seen) {
        self_seen = true;  // the materialised self loop is applied from the thread's own registers
        continue;
      }
      if (cnt < ALIGN_MAX_INLINE) d |= (unsigned long long)((c + shift - gbase) & 31) << (4 + 5 * cnt);
      ++cnt;
    }
    if (!self_seen) atomicOr(flags, 2);
    if (cnt > ALIGN_MAX_INLINE) {
      atomicOr(flags, 1);
      cnt = ALIGN_MAX_INLINE;
    }
    agg_desc[nr] = d | (unsigned long long)cnt;
  }
}

// CSR entries, once rowptr_a (scan of deg_a) exists: one thread per source row
__global__ void align_cols_kernel(fitgnn_pack in, const int32_t* __restrict__ new_of_old, const int32_t* __restrict__ rowptr_a,
                                  int32_t* __restrict__ col_a) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= in.n_rows) return;
  const int nr = new_of_old[r];
  const int shift = nr - (int)r;
  const int e0 = in.rowptr[r], e1 = in.rowptr[r + 1];
  int dst = rowptr_a[nr];
  for (int e = e0; e < e1; ++e) col_a[dst++] = in.col[e] + shift;  // a subgraph's rows all move by the same shift
}

__global__ void remap_rows_kernel(const int32_t* __restrict__ rows, int64_t n, const int32_t* __restrict__ new_of_old,
                                  int32_t* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = new_of_old[rows[i]];
}

struct AlignWs {
  int64_t* status;
  int32_t* flags;
  uint64_t* keys;
  int32_t* order;
  int32_t* size_sorted;
  int32_t* deg_a;
  void* rest;
  size_t rest_bytes;
  bool ok;
};

static AlignWs carve(void* ws, size_t ws_bytes, int64_t n_sub, int64_t n_al) {
  Bump b(ws, ws_bytes);
  AlignWs w;
  const size_t ns = (size_t)(n_sub > 0 ? n_sub : 1);
  w.status = b.take<int64_t>(8);
  w.flags = b.take<int32_t>(8);
  w.keys = b.take<uint64_t>(ns);
  w.order = b.take<int32_t>(ns);
  w.size_sorted = b.take<int32_t>(ns);
  w.deg_a = n_al >= 0 ? b.take<int32_t>((size_t)n_al + 1) : nullptr;
  w.ok = b.ok;
  w.rest = b.ok ? b.here() : nullptr;
  w.rest_bytes = b.ok ? b.left() : 0;
  return w;
}

}  // namespace fitgnn

using namespace fitgnn;

extern "C" size_t fitgnn_pack_align_workspace_bytes(int64_t n_sub, int64_t n_rows_aligned_max) {
  const size_t ns = (size_t)(n_sub > 0 ? n_sub : 1);
  const size_t na = (size_t)(n_rows_aligned_max > 0 ? n_rows_aligned_max : 1);
  const size_t sort_b = sort_ws_bytes(n_sub), scan_b = scan_ws_bytes((int64_t)na + 1);
  return 4096 + align_up(ns * 8) + 2 * align_up(ns * 4) + align_up((na + 1) * 4) + (sort_b > scan_b ? sort_b : scan_b) + 1024;
}

extern "C" int fitgnn_pack_align_plan(const fitgnn_pack* in, int group, int policy, int32_t* new_sub_ptr,
                                      int64_t* host_n_rows_aligned, int* host_alignable, void* ws, size_t ws_bytes,
                                      void* stream) {
  FG_REQUIRE(in && new_sub_ptr && host_n_rows_aligned && host_alignable && in->n_sub >= 0, FITGNN_EINVAL,
             "pack_align_plan: bad arguments");
  FG_REQUIRE(group == 32, FITGNN_EUNSUP, "pack_align_plan: only groups of 32 rows (one TMEM lane quadrant) are supported");
  FG_REQUIRE(policy == FITGNN_ALIGN_IN_ORDER || policy == FITGNN_ALIGN_BY_DEGREE, FITGNN_EINVAL,
             "pack_align_plan: unknown policy %d", policy);
  AlignWs w = carve(ws, ws_bytes, in->n_sub, -1);
  FG_REQUIRE(ws && w.ok, FITGNN_EWS, "pack_align_plan: workspace too small");
  cudaStream_t st = as_stream(stream);
  const int64_t n_sub = in->n_sub;
  if (n_sub > 0) {
    const unsigned blocks = (unsigned)ceil_div(n_sub, 256);
    if (policy == FITGNN_ALIGN_BY_DEGREE) {
      align_keys_kernel<<<blocks, 256, 0, st>>>(*in, group, w.keys);
      FG_LAUNCH_CHECK();
      FG_TRY(sort_u64(w.keys, nullptr, n_sub, 48, w.rest, w.rest_bytes, st));
    }
    align_sizes_kernel<<<blocks, 256, 0, st>>>(in->sub_ptr, policy == FITGNN_ALIGN_BY_DEGREE ? w.keys : nullptr, n_sub,
                                               w.order, w.size_sorted);
    FG_LAUNCH_CHECK();
  }
  // The placement is a sequential greedy (where a subgraph goes depends on where its predecessor ended, and the gap filling
  // consumes the order from both ends), one decision per subgraph.  Round 1 ran it as a single GPU thread: 21 ms for the
  // 1.06 M subgraphs of the products pack (~20 ns per dependent global load).  The plan call synchronises with the host anyway
  // (it returns the aligned row count), so the walk now runs on the host over the two small arrays (8 bytes per subgraph down,
  // 4 bytes up): ~3 ns per step.  (The oracle's aligned_layout restates the same walk for the tests.)
  int64_t h[2] = {0, 0};
  if (n_sub > 0) {
    std::vector<int32_t> order((size_t)n_sub), sizes((size_t)n_sub), start((size_t)n_sub + 1);
    FG_CUDA(cudaMemcpyAsync(order.data(), w.order, (size_t)n_sub * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    FG_CUDA(cudaMemcpyAsync(sizes.data(), w.size_sorted, (size_t)n_sub * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    FG_CUDA(cudaStreamSynchronize(st));
    const bool fill_gaps = policy == FITGNN_ALIGN_BY_DEGREE;
    int64_t pos = 0, head = 0, tail = n_sub - 1;
    int bad = 0;
    while (head <= tail) {
      const int size = sizes[(size_t)head];
      if (size > group) { bad = 1; break; }
      const int rem = group - (int)(pos & (group - 1));
      if (size <= rem) {
        start[(size_t)order[(size_t)head]] = (int32_t)pos;
        pos += size;
        ++head;
      } else if (fill_gaps && head < tail && sizes[(size_t)tail] <= rem) {
        start[(size_t)order[(size_t)tail]] = (int32_t)pos;
        pos += sizes[(size_t)tail];
        --tail;
      } else {
        pos += rem;  // close the group with padding
      }
      if (pos > 0x7fffff00ll) { bad = 2; break; }
    }
    start[(size_t)n_sub] = (int32_t)pos;
    h[0] = bad;
    h[1] = pos;
    if (bad == 0) {
      FG_CUDA(cudaMemcpyAsync(new_sub_ptr, start.data(), ((size_t)n_sub + 1) * sizeof(int32_t), cudaMemcpyHostToDevice, st));
      FG_CUDA(cudaStreamSynchronize(st));  // `start` dies with this scope
    }
  } else {
    FG_CUDA(cudaMemsetAsync(new_sub_ptr, 0, sizeof(int32_t), st));
    FG_CUDA(cudaStreamSynchronize(st));
  }
  FG_REQUIRE(h[0] != 2, FITGNN_ERANGE, "pack_align_plan: aligned row count exceeds int32");
  *host_alignable = h[0] == 0 ? 1 : 0;
  *host_n_rows_aligned = h[1];
  return FITGNN_OK;
}

extern "C" int fitgnn_pack_align_fill(const fitgnn_pack* in, const int32_t* new_sub_ptr, int group,
                                      int64_t n_rows_aligned, const fitgnn_pack* out, int32_t* orig_row,
                                      int32_t* new_of_old, uint64_t* agg_desc, int* host_flags, void* ws,
                                      size_t ws_bytes, void* stream) {
  FG_REQUIRE(in && out && new_sub_ptr && orig_row && new_of_old && agg_desc && host_flags, FITGNN_EINVAL,
             "pack_align_fill: bad arguments");
  FG_REQUIRE(group == 32, FITGNN_EUNSUP, "pack_align_fill: only groups of 32 rows are supported");
  FG_REQUIRE(out->n_rows == n_rows_aligned && out->nnz == in->nnz && out->n_core == in->n_core, FITGNN_EINVAL,
             "pack_align_fill: output pack sizes do not match the plan");
  AlignWs w = carve(ws, ws_bytes, in->n_sub, n_rows_aligned);
  FG_REQUIRE(ws && w.ok, FITGNN_EWS, "pack_align_fill: workspace too small");
  cudaStream_t st = as_stream(stream);
  FG_CUDA(cudaMemsetAsync(w.flags, 0, sizeof(int32_t), st));
  *host_flags = 0;
  int32_t* rowptr_a = const_cast<int32_t*>(out->rowptr);
  if (n_rows_aligned > 0) {
    align_init_rows_kernel<<<(unsigned)ceil_div(n_rows_aligned, 256), 256, 0, st>>>(
        n_rows_aligned, w.deg_a, const_cast<float*>(out->dinv), const_cast<int32_t*>(out->gid),
        const_cast<uint8_t*>(out->is_core), const_cast<uint8_t*>(out->mask), orig_row,
        reinterpret_cast<unsigned long long*>(agg_desc));
    FG_LAUNCH_CHECK();
  }
  if (in->n_sub > 0) {
    align_rows_kernel<<<(unsigned)ceil_div(in->n_sub, 128), 128, 0, st>>>(
        *in, new_sub_ptr, group, w.deg_a, const_cast<float*>(out->dinv), const_cast<int32_t*>(out->gid),
        const_cast<uint8_t*>(out->is_core), const_cast<uint8_t*>(out->mask), orig_row, new_of_old,
        reinterpret_cast<unsigned long long*>(agg_desc), w.flags);
    FG_LAUNCH_CHECK();
  }
  FG_TRY(scan_i32(w.deg_a, n_rows_aligned, rowptr_a, n_rows_aligned + 1, w.rest, w.rest_bytes, st));
  if (in->n_rows > 0) {
    align_cols_kernel<<<(unsigned)ceil_div(in->n_rows, 256), 256, 0, st>>>(*in, new_of_old, rowptr_a,
                                                                           const_cast<int32_t*>(out->col));
    FG_LAUNCH_CHECK();
  }
  if (in->n_core > 0) {
    remap_rows_kernel<<<(unsigned)ceil_div(in->n_core, 256), 256, 0, st>>>(in->core_rows, in->n_core, new_of_old,
                                                                           const_cast<int32_t*>(out->core_rows));
    FG_LAUNCH_CHECK();
  }
  FG_CUDA(cudaMemcpyAsync(const_cast<int32_t*>(out->sub_ptr), new_sub_ptr, (size_t)(in->n_sub + 1) * sizeof(int32_t),
                          cudaMemcpyDeviceToDevice, st));
  FG_CUDA(cudaMemcpyAsync(host_flags, w.flags, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  FG_CUDA(cudaStreamSynchronize(st));
  return FITGNN_OK;
}
