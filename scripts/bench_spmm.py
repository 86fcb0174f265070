"""Times the two SpMM launches of the products workload for the library named by FITGNN_B200_LIB."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fitgnn_b200 as fg
n, e, F, C = 2449029, 61859140, 100, 47
dev = torch.device("cuda:0")
ei, part, cw, k = fg.synth.planted_partition(n, e, 0.5, seed=0, device=dev)
part = fg.synth.relabel_partition_reference_order(part)
pack = fg.build_pack(ei, part, k, "none"); del ei
X = torch.rand(n, 104, device=dev); H = torch.rand(n, 512, device=dev)
def t(fn, reps=10):
    for _ in range(3): fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize(); return a.elapsed_time(b) / reps
o1 = (torch.empty(n, 104, dtype=torch.bfloat16, device=dev), torch.empty(n, 104, dtype=torch.bfloat16, device=dev))
o2 = (torch.empty(n, 512, dtype=torch.bfloat16, device=dev), torch.empty(n, 512, dtype=torch.bfloat16, device=dev))
o3 = torch.empty(n, 512, device=dev)
t0 = t(lambda: fg.ops.spmm_symnorm(pack.rowptr, pack.col, pack.dinv, X, 104, pack.gid, out=o1, split=True))
t1 = t(lambda: fg.ops.spmm_symnorm(pack.rowptr, pack.col, pack.dinv, H, 512, out_rows=pack.core_rows, out=o2, split=True))
t2 = t(lambda: fg.ops.spmm_symnorm(pack.rowptr, pack.col, pack.dinv, H, 512, out=o3))
print(f"{os.environ.get('FITGNN_B200_LIB','default'):40s} spmm0 {t0:.3f} ms   spmm1(split,core rows) {t1:.3f} ms   spmm 512 fp32 all rows {t2:.3f} ms")
