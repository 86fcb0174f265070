// Row-group alignment of a pack: a derived layout in which no subgraph straddles a multiple of `group` rows, so
// that every CSR entry of a row points into the row's own aligned group of 32 rows.  That makes the normalised
// aggregation Â·H a warp-local operation inside the tensor-core GEMM epilogue (thread = row, one TMEM lane quadrant
// = one group; gemm_tcgen05.cu, fitgnn_gcn_transform_aggregate) — the layer order "GCNConv -> ELU -> next GCNConv's
// propagate" of /root/reference/network.py:31-33 then needs no separate SpMM launch and no HBM round trip.
//
// Subgraphs keep the pack's (= the reference's subgraph_list) order and are placed greedily; when the next one does
// not fit into what is left of the current group, the group is closed with padding rows (empty CSR rows, dinv = 0,
// orig_row = -1).  Padding therefore always sits at the tail of a group.
//
// Per aligned row the fill also emits an aggregation descriptor: bits [0,4) = number c of non-self entries
// (c <= 12), bits [4+5j, 9+5j) = lane (row index inside the group) of the j-th one, duplicates kept.
#include "common.cuh"

namespace fitgnn {

constexpr int ALIGN_MAX_INLINE = 12;

// One thread walks the subgraphs in order (the placement of subgraph s depends on where s-1 ended).
// status[0] = 1 when a subgraph has more than `group` rows (not alignable); status[1..2] = aligned row count.
__global__ void align_plan_kernel(const int32_t* __restrict__ sub_ptr, int64_t n_sub, int group,
                                  int32_t* __restrict__ new_start, int64_t* __restrict__ status) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  int64_t pos = 0;
  int bad = 0;
  int32_t prev = sub_ptr[0];
#pragma unroll 8
  for (int64_t s = 0; s < n_sub; ++s) {
    const int32_t next = __ldg(sub_ptr + s + 1);
    const int size = next - prev;
    prev = next;
    if (size > group) bad = 1;
    const int off = (int)(pos & (group - 1));  // group is a power of two
    if (off + size > group && size <= group) pos += group - off;
    new_start[s] = (int32_t)pos;
    pos += size;
    if (pos > 0x7fffff00ll) { bad = 2; break; }
  }
  new_start[n_sub] = (int32_t)pos;
  status[0] = bad;
  status[1] = pos;
}

// one thread per subgraph: moves its rows, relabels its CSR entries, writes the padding that follows it
__global__ void align_fill_kernel(fitgnn_pack in, const int32_t* __restrict__ new_start, int group, int32_t* rowptr_a,
                                  int32_t* col_a, float* dinv_a, int32_t* gid_a, uint8_t* is_core_a, uint8_t* mask_a,
                                  int32_t* orig_row, int32_t* new_of_old, unsigned long long* agg_desc,
                                  int32_t* __restrict__ flags) {
  const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= in.n_sub) return;
  const int a = in.sub_ptr[s], b = in.sub_ptr[s + 1];
  const int ns = new_start[s], nn = new_start[s + 1];
  const int shift = ns - a;
  const int gbase = ns & ~(group - 1);
  for (int r = a; r < b; ++r) {
    const int nr = r + shift;
    const int e0 = in.rowptr[r], e1 = in.rowptr[r + 1];
    new_of_old[r] = nr;
    orig_row[nr] = r;
    rowptr_a[nr] = e0;
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
      const int nc = c + shift;
      col_a[e] = nc;
      if (c == r && !self_seen) {
        self_seen = true;  // the materialised self loop is applied from the thread's own registers
        continue;
      }
      if (cnt < ALIGN_MAX_INLINE) d |= (unsigned long long)((nc - gbase) & 31) << (4 + 5 * cnt);
      ++cnt;
    }
    if (!self_seen) atomicOr(flags, 2);
    if (cnt > ALIGN_MAX_INLINE) {
      atomicOr(flags, 1);
      cnt = ALIGN_MAX_INLINE;
    }
    agg_desc[nr] = d | (unsigned long long)cnt;
  }
  const int e_end = in.rowptr[b];
  for (int p = b + shift; p < nn; ++p) {  // padding rows closing the group
    orig_row[p] = -1;
    rowptr_a[p] = e_end;
    dinv_a[p] = 0.f;
    gid_a[p] = 0;
    is_core_a[p] = 0;
    mask_a[p] = 0;
    agg_desc[p] = 0ull;
  }
  if (s == in.n_sub - 1) rowptr_a[nn] = e_end;
}

__global__ void remap_rows_kernel(const int32_t* __restrict__ rows, int64_t n, const int32_t* __restrict__ new_of_old,
                                  int32_t* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = new_of_old[rows[i]];
}

}  // namespace fitgnn

using namespace fitgnn;

extern "C" int fitgnn_pack_align_plan(const int32_t* sub_ptr, int64_t n_sub, int group, int32_t* new_sub_ptr,
                                      int64_t* host_n_rows_aligned, int* host_alignable, void* ws, size_t ws_bytes,
                                      void* stream) {
  FG_REQUIRE(sub_ptr && new_sub_ptr && host_n_rows_aligned && host_alignable && n_sub >= 0, FITGNN_EINVAL,
             "pack_align_plan: bad arguments");
  FG_REQUIRE(group == 32, FITGNN_EUNSUP, "pack_align_plan: only groups of 32 rows (one TMEM lane quadrant) are supported");
  FG_REQUIRE(ws && ws_bytes >= 64, FITGNN_EWS, "pack_align_plan: workspace too small (needs 64 bytes)");
  cudaStream_t st = as_stream(stream);
  int64_t* status = static_cast<int64_t*>(ws);
  align_plan_kernel<<<1, 32, 0, st>>>(sub_ptr, n_sub, group, new_sub_ptr, status);
  FG_LAUNCH_CHECK();
  int64_t h[2] = {0, 0};
  FG_CUDA(cudaMemcpyAsync(h, status, sizeof(h), cudaMemcpyDeviceToHost, st));
  FG_CUDA(cudaStreamSynchronize(st));
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
  FG_REQUIRE(ws && ws_bytes >= 64, FITGNN_EWS, "pack_align_fill: workspace too small (needs 64 bytes)");
  cudaStream_t st = as_stream(stream);
  int32_t* flags = static_cast<int32_t*>(ws);
  FG_CUDA(cudaMemsetAsync(flags, 0, sizeof(int32_t), st));
  *host_flags = 0;
  if (in->n_sub > 0) {
    align_fill_kernel<<<(unsigned)ceil_div(in->n_sub, 128), 128, 0, st>>>(
        *in, new_sub_ptr, group, const_cast<int32_t*>(out->rowptr), const_cast<int32_t*>(out->col),
        const_cast<float*>(out->dinv), const_cast<int32_t*>(out->gid), const_cast<uint8_t*>(out->is_core),
        const_cast<uint8_t*>(out->mask), orig_row, new_of_old, reinterpret_cast<unsigned long long*>(agg_desc), flags);
    FG_LAUNCH_CHECK();
    if (in->n_core > 0) {
      remap_rows_kernel<<<(unsigned)ceil_div(in->n_core, 256), 256, 0, st>>>(in->core_rows, in->n_core, new_of_old,
                                                                             const_cast<int32_t*>(out->core_rows));
      FG_LAUNCH_CHECK();
    }
    FG_CUDA(cudaMemcpyAsync(const_cast<int32_t*>(out->sub_ptr), new_sub_ptr, (size_t)(in->n_sub + 1) * sizeof(int32_t),
                            cudaMemcpyDeviceToDevice, st));
  } else {
    FG_CUDA(cudaMemsetAsync(const_cast<int32_t*>(out->rowptr), 0, sizeof(int32_t), st));
    FG_CUDA(cudaMemsetAsync(const_cast<int32_t*>(out->sub_ptr), 0, sizeof(int32_t), st));
  }
  FG_CUDA(cudaMemcpyAsync(host_flags, flags, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  FG_CUDA(cudaStreamSynchronize(st));
  return FITGNN_OK;
}
