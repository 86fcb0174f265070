"""What does the epilogue of the layer-2 transform cost?  2,449,029 x 512 x 512, fp16 plane in / out, one W plane, W-stationary
pairs: ELU vs no activation, bias vs none, row_scale vs none.  python scripts/bench_epilogue_cost.py"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fitgnn_b200 as fg
dev = torch.device("cuda:0")
M, K, N = 2449029, 512, 512
def t(fn, reps=10):
    for _ in range(3): fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize(); return a.elapsed_time(b) / reps
A = (torch.randn(M, K, device=dev) * 0.1).half()
W1 = fg.ops.split_f16(torch.randn(N, K, device=dev) / K ** 0.5, lo=False)
bias = torch.randn(N, device=dev) * 0.1
rs = torch.rand(M, device=dev) + 0.5
for label, b, act, r in [("ELU + bias + row_scale (the forward's call)", bias, fg.ops.ACT_ELU, rs), ("ELU + bias", bias, fg.ops.ACT_ELU, None),
                         ("ELU only", None, fg.ops.ACT_ELU, None), ("bias only", bias, fg.ops.ACT_NONE, None),
                         ("nothing (convert + store)", None, fg.ops.ACT_NONE, None)]:
    ms = t(lambda: fg.ops.gemm_f16(A, W1, b, act, row_scale=r, out_f16=True))
    print(f"{label:46s} {ms:.3f} ms", flush=True)
from fitgnn_b200._lib import set_tuning
set_tuning("gemm_debug", 1)
ms = t(lambda: fg.ops.gemm_f16(A, W1, bias, fg.ops.ACT_ELU, row_scale=rs, out_f16=True))
print(f"{'ELU + bias + row_scale, NO TMA stores':46s} {ms:.3f} ms", flush=True)
set_tuning("gemm_debug", 0)
