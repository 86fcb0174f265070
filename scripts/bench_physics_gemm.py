"""The Physics-shaped layer-1 transform (configs[2]: 34,493 unique rows x 8,415 features -> 512, transform-first on the
de-duplicated rows): the one place BASELINE.json asks for tensor-pipe utilisation at large K.
python scripts/bench_physics_gemm.py [--profiler-range]   (ncu: --profile-from-start off)"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fitgnn_b200 as fg
dev = torch.device("cuda:0")
M, K, N = 34493, 8415, 512
g = torch.Generator(device=dev).manual_seed(0)
X = torch.rand(M, K, generator=g, device=dev)
W = (torch.rand(N, K, generator=g, device=dev) - 0.5) * 0.03
Kp = fg.ops.pad8(K)
def run_tc():
    a = fg.ops.split_bf16(X, cols=K, ldo=Kp)          # fp32 features -> bf16 hi/lo planes (part of the call pattern)
    w = fg.ops.split_bf16(W, cols=K, ldo=Kp)
    return fg.ops.gemm_bias_act(a, w, None, 0, 0, K=Kp, N=N, precision=fg.ops.GEMM_BF16X3)
a_pl, w_pl = fg.ops.split_bf16(X, cols=K, ldo=Kp), fg.ops.split_bf16(W, cols=K, ldo=Kp)
def run_gemm_only():
    return fg.ops.gemm_bias_act(a_pl, w_pl, None, 0, 0, K=Kp, N=N, precision=fg.ops.GEMM_BF16X3)
Xp = torch.nn.functional.pad(X, (0, fg.ops.pad4(K) - K)); Wp = torch.nn.functional.pad(W, (0, fg.ops.pad4(K) - K))
def run_fp32():
    return fg.ops.gemm_bias_act(Xp, Wp, None, 0, 0, K=Xp.shape[1], precision=fg.ops.GEMM_FP32)
def t(fn, reps=10):
    for _ in range(3): fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize(); return a.elapsed_time(b) / reps
flops = 2.0 * M * K * N
want = (X[:512].double() @ W.double().T)
err = float((run_tc()[:512].double() - want).abs().max() / want.abs().max())
if "--profiler-range" in sys.argv:
    torch.cuda.synchronize(); torch.cuda.cudart().cudaProfilerStart(); run_gemm_only(); torch.cuda.synchronize(); torch.cuda.cudart().cudaProfilerStop()
else:
    for name, fn in (("bf16x3 tcgen05 GEMM only", run_gemm_only), ("bf16x3 incl. the hi/lo split of X and W", run_tc), ("fp32 SIMT", run_fp32)):
        ms = t(fn)
        print(f"{name:42s} {ms:7.3f} ms  logical {flops/ms/1e9:7.1f} TFLOP/s  (x3 bf16 MMAs: {3*flops/ms/1e9:7.1f})", flush=True)
    print(f"max rel err vs fp64 (512 rows): {err:.2e}")
