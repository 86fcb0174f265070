// fitgnn_gcn_forward: the whole GCN forward over one pack in ONE C call — the schedule a non-Python host would otherwise have
// to re-implement (fitgnn_b200/engine.py, classic schedule).  Replaces, for every subgraph of the pack at once, the body of
// Classify_node.forward / Regress_node.forward (/root/reference/network.py:29-35, :58-64):
//     for i in range(num_layers): x = F.elu(conv_i(x, edge_index)); (dropout = identity in eval)
//     x = lt1(x); log_softmax / identity
// as it is driven by node_infer_Gs_GD (run.py:59-77) and the per-query loop (inference.py:672-688), returning the rows the
// callers read (core rows, run.py:73) in pack order.
//
// Schedule (all re-associations of the reference's fp32 arithmetic, same as engine.PackedForward):
//   layer 0, F > hidden  : transform the DE-DUPLICATED feature rows once (X W^T)[gid], then aggregate through gid
//   layer 0, F <= hidden : aggregate the F-wide rows first, then transform (+bias, ELU)
//   layers >= 1          : aggregate, then transform; the LAST layer only on the core rows
//   head                 : lt1 (+ log_softmax / softmax) on the core rows
// Stream-ordered, no host synchronisation (hub rows are found and consumed on the device), no allocation: everything lives
// in the caller's workspace.  Graph-capturable.
#include <cuda_bf16.h>
#include "common.cuh"

extern "C" int fitgnn_spmm_symnorm_devhub(const int32_t* rowptr, const int32_t* col, const float* dinv, const float* X,
                                          int64_t ldx, int width, const int32_t* src_index, const float* bias, int act,
                                          const int32_t* out_rows, int64_t n_out, void* Y, void* Y_lo, int64_t ldy,
                                          const int32_t* hub_list, const int32_t* hub_count, int hub_cap, int hub_deg,
                                          void* stream);

namespace fitgnn {

namespace {

constexpr int HUB_DEG = 256;

inline int pad4i(int n) { return (n + 3) / 4 * 4; }
inline int pad8i(int n) { return (n + 7) / 8 * 8; }

struct FwdPlan {
  int L, F, H, C, Fw, Fp, Hp;  // Fw = pad4(F): columns the layer-0 SpMM touches; Fp / Hp = K pitch of the bf16 planes
  bool tc, transform_first;
  int64_t n_rows, n_core, n_src, hub_cap;
  // workspace offsets (bytes)
  size_t o_hubs_all, o_hubs_core, o_cnt, o_w[16], o_wl, o_xplanes, o_a, o_h0, o_h1, total;
};

int make_plan(const fitgnn_pack* p, const fitgnn_weights* w, int precision, FwdPlan& pl) {
  FG_REQUIRE(p && w, FITGNN_EINVAL, "gcn_forward: null pack / weights");
  FG_REQUIRE(w->n_layers >= 1 && w->n_layers <= 8, FITGNN_EUNSUP, "gcn_forward: 1..8 conv layers (got %d)", w->n_layers);
  FG_REQUIRE(w->in_features > 0 && w->hidden > 0 && w->n_classes > 0, FITGNN_EINVAL, "gcn_forward: bad model dimensions");
  FG_REQUIRE(w->hidden % 4 == 0, FITGNN_EUNSUP, "gcn_forward: hidden width must be a multiple of 4");
  FG_REQUIRE(precision == FITGNN_GEMM_FP32 || precision == FITGNN_GEMM_BF16X3, FITGNN_EINVAL, "gcn_forward: unknown precision");
  pl.L = w->n_layers; pl.F = w->in_features; pl.H = w->hidden; pl.C = w->n_classes;
  pl.tc = precision == FITGNN_GEMM_BF16X3;
  FG_REQUIRE(!pl.tc || pl.H % 8 == 0, FITGNN_EUNSUP, "gcn_forward: BF16X3 needs hidden %% 8 == 0");
  pl.Fw = pad4i(pl.F);
  pl.Fp = pl.tc ? pad8i(pl.F) : pl.Fw;
  pl.Hp = pl.H;
  pl.transform_first = pl.F > pl.H;
  pl.n_rows = p->n_rows; pl.n_core = p->n_core; pl.n_src = p->n_src;
  pl.hub_cap = p->nnz / HUB_DEG + 1;
  Bump b(nullptr, (size_t)1 << 62);
  auto take = [&](size_t bytes) { size_t o = b.off; b.take<char>(bytes > 0 ? bytes : 1); return o; };
  pl.o_hubs_all = take((size_t)pl.hub_cap * 4);
  pl.o_hubs_core = take((size_t)pl.hub_cap * 4);
  pl.o_cnt = take(64);
  const size_t esz = pl.tc ? 4 : 4;  // bf16 hi + lo planes = 4 bytes per element, like fp32
  for (int i = 0; i < pl.L; ++i) pl.o_w[i] = take((size_t)pl.H * (i == 0 ? pl.Fp : pl.Hp) * esz);
  pl.o_wl = take((size_t)pl.C * pl.Hp * esz);
  // layer-0 operand: transform-first = the split feature table (tc only); aggregate-first = Â·X [n_rows0, Fp]
  const int64_t rows0 = (pl.L == 1) ? pl.n_core : pl.n_rows;  // rows layer 0 produces
  pl.o_xplanes = take(pl.transform_first && pl.tc ? (size_t)pl.n_src * pl.Fp * 4 : 0);
  // A: aggregate output of the current layer (planes or fp32), Z of the transform-first layer
  size_t a_bytes = pl.transform_first ? (size_t)pl.n_src * pl.H * 4 : (size_t)rows0 * pl.Fp * 4;
  if (pl.L > 1) a_bytes = a_bytes > (size_t)pl.n_rows * pl.Hp * 4 ? a_bytes : (size_t)pl.n_rows * pl.Hp * 4;
  pl.o_a = take(a_bytes);
  // h ping-pong: hidden state of a layer [rows, H] (fp32, or planes when it only feeds the head)
  pl.o_h0 = take((size_t)rows0 * pl.H * 4);
  pl.o_h1 = take(pl.L > 1 ? (size_t)pl.n_rows * pl.H * 4 : 0);
  pl.total = b.off + 256;
  return FITGNN_OK;
}

// weights [out, in] fp32 (ld = in) -> what the GEMM takes: bf16 hi/lo planes with the K pitch padded (tc) or a padded fp32 copy
__global__ void pad_copy_kernel(const float* __restrict__ W, int rows, int cols, int ld_out, float* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)rows * ld_out) return;
  const int r = (int)(i / ld_out), c = (int)(i % ld_out);
  out[i] = c < cols ? W[(int64_t)r * cols + c] : 0.f;
}

struct Operand {  // a dense GEMM operand in the precision of the run
  void* hi;
  void* lo;  // null for fp32
  int64_t ld;
};

int prep_weight(const float* W, int rows, int cols, int ld_pad, bool tc, char* dst, Operand& op, cudaStream_t st) {
  FG_REQUIRE(W, FITGNN_EINVAL, "gcn_forward: null weight pointer");
  op.ld = ld_pad;
  if (tc) {
    op.hi = dst;
    op.lo = dst + (size_t)rows * ld_pad * 2;
    return fitgnn_split_bf16(W, cols, rows, cols, op.hi, op.lo, ld_pad, st);
  }
  op.hi = dst;
  op.lo = nullptr;
  const int64_t n = (int64_t)rows * ld_pad;
  pad_copy_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(W, rows, cols, ld_pad, static_cast<float*>(op.hi));
  FG_LAUNCH_CHECK();
  return FITGNN_OK;
}

}  // namespace
}  // namespace fitgnn

using namespace fitgnn;

extern "C" size_t fitgnn_gcn_forward_workspace_bytes(const fitgnn_pack* pack, const fitgnn_weights* weights, int precision) {
  FwdPlan pl;
  if (make_plan(pack, weights, precision, pl) != FITGNN_OK) return 0;
  return pl.total;
}

extern "C" int fitgnn_gcn_forward(const fitgnn_pack* p, const float* X, int64_t ldx, const fitgnn_weights* w, int head,
                                  int precision, float* out, int64_t ld_out, void* ws, size_t ws_bytes, void* stream) {
  FwdPlan pl;
  FG_TRY(make_plan(p, w, precision, pl));
  FG_REQUIRE(X && out && ws, FITGNN_EINVAL, "gcn_forward: null X / out / workspace");
  FG_REQUIRE(ws_bytes >= pl.total, FITGNN_EWS, "gcn_forward: workspace needs %zu bytes (got %zu)", pl.total, ws_bytes);
  FG_REQUIRE(((uintptr_t)ws & 255) == 0, FITGNN_EINVAL, "gcn_forward: workspace must be 256-byte aligned");
  FG_REQUIRE(ldx >= pl.Fw && ldx % 4 == 0, FITGNN_EUNSUP,
             "gcn_forward: X needs a row pitch >= %d floats, a multiple of 4, with zero-filled padding columns (got %lld)", pl.Fw,
             (long long)ldx);
  FG_REQUIRE(ld_out >= pl.C, FITGNN_EINVAL, "gcn_forward: ld_out smaller than the class count");
  FG_REQUIRE(head >= FITGNN_HEAD_IDENTITY && head <= FITGNN_HEAD_SOFTMAX, FITGNN_EINVAL, "gcn_forward: unknown head %d", head);
  FG_REQUIRE(w->conv_weight && w->conv_bias && w->lt1_weight, FITGNN_EINVAL, "gcn_forward: null weight arrays");
  if (pl.n_core == 0) return FITGNN_OK;
  cudaStream_t st = as_stream(stream);
  char* base = static_cast<char*>(ws);
  const int prec = precision;
  const bool tc = pl.tc;

  // hub rows (deg >= 256) of (a) all rows, (b) the core rows: found on the device, consumed by the hub kernels without a sync
  int32_t* hubs_all = reinterpret_cast<int32_t*>(base + pl.o_hubs_all);
  int32_t* hubs_core = reinterpret_cast<int32_t*>(base + pl.o_hubs_core);
  int32_t* cnt_all = reinterpret_cast<int32_t*>(base + pl.o_cnt);
  int32_t* cnt_core = cnt_all + 8;
  const bool core_is_all = p->n_core == p->n_rows;  // mode 'none': core_rows is the identity
  const int32_t* core_rows = core_is_all ? nullptr : p->core_rows;
  FG_TRY(fitgnn_spmm_hubs(p->rowptr, nullptr, p->n_rows, HUB_DEG, hubs_all, cnt_all, (int)pl.hub_cap, stream));
  FG_TRY(fitgnn_spmm_hubs(p->rowptr, core_rows, p->n_core, HUB_DEG, hubs_core, cnt_core, (int)pl.hub_cap, stream));

  Operand W[8], Wl;
  for (int i = 0; i < pl.L; ++i)
    FG_TRY(prep_weight(w->conv_weight[i], pl.H, i == 0 ? pl.F : pl.H, i == 0 ? pl.Fp : pl.Hp, tc, base + pl.o_w[i], W[i], st));
  FG_TRY(prep_weight(w->lt1_weight, pl.C, pl.H, pl.Hp, tc, base + pl.o_wl, Wl, st));

  auto spmm = [&](const float* src, int64_t ld, int width, const int32_t* src_index, const float* bias, int act, bool last,
                  void* y, void* ylo, int64_t ldy) {
    return fitgnn_spmm_symnorm_devhub(p->rowptr, p->col, p->dinv, src, ld, width, src_index, bias, act, last ? core_rows : nullptr,
                                      last ? p->n_core : p->n_rows, y, ylo, ldy, last ? hubs_core : hubs_all,
                                      last ? cnt_core : cnt_all, (int)pl.hub_cap, HUB_DEG, stream);
  };
  auto gemm = [&](const Operand& A, const Operand& Wt, const float* bias, int64_t M, int K, int N, int act, int hd, void* y,
                  void* ylo, int64_t ldy) {
    return fitgnn_gemm_bias_act_split(prec, A.hi, A.lo, A.ld, Wt.hi, Wt.lo, Wt.ld, bias, M, K, N, act, hd, y, ylo, ldy, stream);
  };

  char* hbuf[2] = {base + pl.o_h0, base + pl.o_h1};
  Operand h{};  // current hidden state (fp32: hi only; planes when it feeds only the head)
  for (int i = 0; i < pl.L; ++i) {
    const bool last = i == pl.L - 1;
    const int64_t rows = last ? pl.n_core : pl.n_rows;
    // the last conv layer feeds only the head GEMM: emit its bf16 planes straight from the epilogue
    const bool to_planes = tc && last;
    char* hout = hbuf[i & 1];
    Operand hn{hout, to_planes ? hout + (size_t)rows * pl.H * 2 : nullptr, pl.H};
    if (i == 0 && pl.transform_first) {
      Operand A{const_cast<float*>(X), nullptr, ldx};
      if (tc) {
        A.hi = base + pl.o_xplanes;
        A.lo = base + pl.o_xplanes + (size_t)pl.n_src * pl.Fp * 2;
        A.ld = pl.Fp;
        FG_TRY(fitgnn_split_bf16(X, ldx, pl.n_src, pl.F, A.hi, A.lo, pl.Fp, stream));
      }
      float* Z = reinterpret_cast<float*>(base + pl.o_a);
      FG_TRY(gemm(A, W[0], nullptr, pl.n_src, tc ? pl.Fp : pl.F, pl.H, FITGNN_ACT_NONE, FITGNN_HEAD_IDENTITY, Z, nullptr, pl.H));
      // propagate + bias + ELU in the SpMM epilogue; planes when this is also the last layer
      FG_TRY(spmm(Z, pl.H, pl.H, p->gid, w->conv_bias[0], FITGNN_ACT_ELU, last, hn.hi, hn.lo, pl.H));
    } else {
      const float* src = i == 0 ? X : static_cast<const float*>(h.hi);
      const int64_t ld = i == 0 ? ldx : pl.H;
      const int width = i == 0 ? pl.Fw : pl.H;
      const int Kp = i == 0 ? pl.Fp : pl.Hp;
      Operand A{base + pl.o_a, tc ? base + pl.o_a + (size_t)rows * Kp * 2 : nullptr, Kp};
      if (tc && Kp != width) FG_CUDA(cudaMemsetAsync(A.hi, 0, (size_t)rows * Kp * 4, st));  // K-pad columns of the planes
      FG_TRY(spmm(src, ld, width, i == 0 ? p->gid : nullptr, nullptr, FITGNN_ACT_NONE, last, A.hi, A.lo, Kp));
      FG_TRY(gemm(A, W[i], w->conv_bias[i], rows, tc ? Kp : width, pl.H, FITGNN_ACT_ELU, FITGNN_HEAD_IDENTITY, hn.hi, hn.lo, pl.H));
    }
    h = hn;
  }
  // head on the core rows: lt1 + (log_)softmax (network.py:34-35)
  if (tc && !h.lo) {  // cannot happen (the last layer emits planes), kept as a guard
    set_error("gcn_forward: internal: head operand is not split");
    return FITGNN_EINVAL;
  }
  return gemm(h, Wl, w->lt1_bias, pl.n_core, pl.H, pl.C, FITGNN_ACT_NONE, head, out, nullptr, ld_out);
}
