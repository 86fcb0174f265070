"""Layer-2 transform of the products workload (2,449,029 x 512 x 512, fp16 plane in and out) under the three operand plans:
W hi/lo (2 MMAs, streaming CTA pairs), ONE W plane on streaming pairs, ONE W plane on W-stationary pairs.
python scripts/bench_gemm1_f16.py"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fitgnn_b200 as fg
from fitgnn_b200._lib import set_tuning
dev = torch.device("cuda:0")
M, K, N = 2449029, 512, 512
def t(fn, reps=10):
    for _ in range(3): fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize(); return a.elapsed_time(b) / reps
A = (torch.randn(M, K, device=dev) * 0.1).half()
W = torch.randn(N, K, device=dev) / K ** 0.5
bias = torch.randn(N, device=dev) * 0.1
rs = torch.rand(M, device=dev) + 0.5
W2 = fg.ops.split_f16(W)
W1 = fg.ops.split_f16(W, lo=False)
out = None
def run(Wp):
    return fg.ops.gemm_f16(A, Wp, bias, fg.ops.ACT_ELU, row_scale=rs, out_f16=True)
gb = (2 * M * K + 2 * M * N) / 1e9
fl = 2.0 * M * K * N / 1e12
NCU = os.environ.get("NCU") == "1"  # profiling: the two single-plane plans only, one warm-up + one launch each
plans = [("W hi/lo, 2 MMAs, streaming pairs", W2, 0), ("W hi/lo, W-stationary if it fits (it does not)", W2, 1),
                      ("one W plane, streaming pairs", W1, 0), ("one W plane, W-stationary pairs", W1, 1)]
if NCU:
    plans = plans[2:]
for label, Wp, ws in [(l_ + f", {e_} epilogue warps", w_, s_ + 2 * (e_ == 8)) for e_ in (12, 8) for (l_, w_, s_) in plans]:
    set_tuning("gemm_pair_ws", ws & 1)
    set_tuning("gemm_wide", -1 if ws & 2 else 0)
    if NCU:
        run(Wp); run(Wp); torch.cuda.synchronize()
        continue
    ms = t(lambda: run(Wp))
    print(f"{label:48s} {ms:.3f} ms  {gb / ms * 1e3:.0f} GB/s algorithmic  {fl / ms * 1e3:.0f} logical TFLOP/s", flush=True)
set_tuning("gemm_pair_ws", 1)
set_tuning("gemm_wide", 0)
