"""Pack-build / projection timing on the products-shaped graph (one-time builders)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import fitgnn_b200 as fg
import bench

args = bench.parse()
dev = torch.device("cuda:0")
n, F, C, ei, part, cw, k, X, sd = bench.generate(args, dev)
torch.cuda.synchronize()
for i in range(3):
    t0 = time.perf_counter()
    pack = fg.build_pack(ei, part, k, args.mode)
    torch.cuda.synchronize()
    print(f"build_pack #{i}: {(time.perf_counter() - t0) * 1e3:.2f} ms", flush=True)
    del pack
for i in range(2):
    t0 = time.perf_counter()
    r = fg.ops.project_adj(ei, part, k)
    torch.cuda.synchronize()
    print(f"project_adj #{i}: {(time.perf_counter() - t0) * 1e3:.2f} ms", flush=True)
    del r
