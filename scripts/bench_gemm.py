"""Times tensor-core GEMM shapes of the products workload: python scripts/bench_gemm.py"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fitgnn_b200 as fg
dev = torch.device("cuda:0")
M = 2449029
def planes(r, c):
    return (torch.randn(r, c, device=dev).to(torch.bfloat16), (torch.randn(r, c, device=dev) * 1e-3).to(torch.bfloat16))
def t(fn, reps=8):
    for _ in range(3): fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize(); return a.elapsed_time(b) / reps
A512 = planes(M, 512)
bias = torch.randn(512, device=dev)
for (N, head, ldy, label) in [(47, 1, 47, "head N=47 log_softmax (direct stores)"), (47, 0, 47, "N=47 identity (direct stores)"),
                               (48, 0, 48, "N=48 identity (TMA stores)"), (48, 1, 48, "N=48 log_softmax (TMA stores)"),
                               (64, 0, 64, "N=64 identity"), (16, 0, 16, "N=16 identity")]:
    W = planes(N, 512)
    out = torch.empty(M, ldy, device=dev)
    ms = t(lambda: fg.ops.gemm_bias_act(A512, W, bias[:N].contiguous(), 0, head, out=out, precision=1, N=N, K=512))
    print(f"{label:45s} {ms:.3f} ms  A-read {M*512*4/ms/1e6:.0f} GB/s", flush=True)
