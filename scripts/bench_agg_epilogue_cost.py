"""What does the epilogue of the fused layer-1 transform (+ layer-2 aggregation) cost?  Products-shaped aligned pack, fp16-plane
output: with / without ELU, with / without the exchange, with / without the TMA stores.  python scripts/bench_agg_epilogue_cost.py"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fitgnn_b200 as fg
from fitgnn_b200 import ops
from fitgnn_b200._lib import set_tuning
dev = torch.device("cuda:0")
n, F, H, C = 2449029, 100, 512, 47
ei, part, cw, k = fg.synth.planted_partition(n, 61859140, 0.5, seed=0, device=dev)
pack = fg.build_pack(ei, part, k, "none")
del ei
sd = fg.synth.init_state_dict(F, H, C, seed=0)
fwd = fg.PackedForward(pack, sd, precision="fp16")
ap = fwd.apack
X = fwd.pack_features(fg.synth.features(n, F, seed=0, device=dev))
fwd(X, packed=True); torch.cuda.synchronize()
A = fwd._planes0
K = ops.pad4(F) + 1
def t(fn, reps=10):
    for _ in range(3): fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize(); return a.elapsed_time(b) / reps
for label, act, desc, dbg in [("ELU + exchange (the forward's call)", ops.ACT_ELU, ap.agg_desc, 0), ("exchange only", ops.ACT_NONE, ap.agg_desc, 0),
                              ("ELU only (plain transform)", ops.ACT_ELU, None, 0), ("nothing (convert + store)", ops.ACT_NONE, None, 0),
                              ("ELU + exchange, NO TMA stores", ops.ACT_ELU, ap.agg_desc, 1), ("nothing, NO TMA stores", ops.ACT_NONE, None, 1)]:
    set_tuning("gemm_debug", dbg)
    ms = t(lambda: ops.gcn_transform_aggregate_f16(A, fwd.W0_f16 if fwd.f16_layer0 else fwd.W0_fold, None, act, desc, ap.dinv if desc is not None else None, K=K, N=H,
                                                   defer_row_scale=desc is not None))
    print(f"{label:40s} {ms:.3f} ms", flush=True)
set_tuning("gemm_debug", 0)
