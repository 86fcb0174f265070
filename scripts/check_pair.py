"""CTA-pair (cta_group::2) GEMM vs the single-CTA kernel and fp64: python scripts/test_pair.py"""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fitgnn_b200 as fg
from fitgnn_b200 import ops
dev = torch.device("cuda:0")
def run(A_pl, W_pl, b, pair, split, N, K):
    fg._lib.set_tuning("gemm_pair", 1 if pair else 0)
    return ops.gemm_bias_act(A_pl, W_pl, b, ops.ACT_ELU, precision=ops.GEMM_BF16X3, N=N, K=K, split_out=split)
ok = True
for (M, K, N) in [(4096, 512, 512), (5000, 512, 512), (4224, 192, 384), (100003, 512, 512), (8192, 1024, 256)]:
    g = torch.Generator().manual_seed(M)
    A = torch.randn(M, K, generator=g); W = torch.randn(N, K, generator=g) / K ** 0.5; b = torch.randn(N, generator=g) * 0.1
    A_pl, W_pl = ops.split_bf16(A.to(dev)), ops.split_bf16(W.to(dev))
    bd = b.to(dev)
    for split in (False, True):
        y1 = run(A_pl, W_pl, bd, False, split, N, K)
        y2 = run(A_pl, W_pl, bd, True, split, N, K)
        torch.cuda.synchronize()
        if split:
            same = torch.equal(y1[0], y2[0]) and torch.equal(y1[1], y2[1])
            got = y2[0].float() + y2[1].float()
        else:
            same = torch.equal(y1, y2)
            got = y2
        want = torch.nn.functional.elu(A[:2000].double() @ W.double().T + b.double())
        err = float((got[:2000].cpu().double() - want).abs().max() / want.abs().max())
        print(f"M={M} K={K} N={N} split={split}: identical to 1-CTA kernel: {same}, rel err vs fp64 {err:.2e}", flush=True)
        ok = ok and same and err < 1e-4
# timing at the benchmark shape
M, K, N = 2449029, 512, 512
A_pl = (torch.randn(M, K, device=dev).to(torch.bfloat16), torch.zeros(M, K, device=dev, dtype=torch.bfloat16))
W = torch.randn(N, K) / K ** 0.5
W_pl = ops.split_bf16(W.to(dev)); bd = torch.zeros(N, device=dev)
out = (torch.empty(M, N, dtype=torch.bfloat16, device=dev), torch.empty(M, N, dtype=torch.bfloat16, device=dev))
for pair in (0, 1, 0, 1):
    fg._lib.set_tuning("gemm_pair", int(pair))
    for _ in range(2):
        ops.gemm_bias_act(A_pl, W_pl, bd, ops.ACT_ELU, precision=ops.GEMM_BF16X3, N=N, K=K, split_out=True, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        ops.gemm_bias_act(A_pl, W_pl, bd, ops.ACT_ELU, precision=ops.GEMM_BF16X3, N=N, K=K, split_out=True, out=out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"pair={pair}: {ms:.3f} ms  = {3 * 2 * M * K * N / ms / 1e9:.0f} bf16 TFLOP/s", flush=True)
print("ALL OK" if ok else "MISMATCH")
