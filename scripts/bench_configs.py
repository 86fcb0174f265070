"""Measures the parity-test configurations of BASELINE.json (configs 1-4: Cora / PubMed / Coauthor-Physics / ZINC shaped
synthetic inputs) on one GPU: pack build, whole-pack forward (eager and CUDA-graph replay), the per-query path, the
restated reference CPU path on a sample, and the max logit error against the oracle on that sample.
Writes gpurun_out/configs_r2.md and .json (per-kernel rooflines included).  Not the headline bench (that is bench.py, config 5)."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fitgnn_b200 as fg  # noqa: E402
from oracle import fitgnn_oracle as fo  # noqa: E402

DEV = torch.device("cuda:0")
_pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
_d = json.load(open(_pk)) if os.path.exists(_pk) else {}
HBM_PEAK, TC_PEAK = _d.get("hbm_gbs", 6650.0), _d.get("bf16_tflops_sustained", 1400.0)


def timeit(fn, reps=20, warm=3):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))


def node_config(name, mode, sample_batches=8):
    n, e_und, F, C, ratio = fg.synth.SHAPES[name]
    ei = fg.synth.powerlaw_graph(n, e_und, seed=0)
    t0 = time.perf_counter()
    partition, comps, C_list = fg.synth.neighborhood_partition(ei, n, ratio, seed=0)
    t_part = time.perf_counter() - t0
    eid = torch.tensor(ei, device=DEV)
    X = fg.synth.features(n, F, seed=0, kind="bow" if F > 256 else "dense", device=DEV)
    sd = fg.synth.init_state_dict(F, 512, C, seed=0)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    proj = fg.coarsen.project(eid, X, partition) if mode == "cluster" else None
    pack = fg.build_pack(eid, torch.tensor(partition.part), partition.k, mode,
                         proj["ac_rowptr"] if proj else None, proj["ac_col"] if proj else None)
    torch.cuda.synchronize(); build_ms = (time.perf_counter() - t0) * 1e3
    Xg = torch.cat([X, proj["Xc"]], 0) if mode == "cluster" else X
    fwd = fg.PackedForward(pack, sd, precision="bf16x3")
    eager_ms = timeit(lambda: fwd(Xg))
    run = fwd.capture(Xg)
    graph_ms = timeit(lambda: run(Xg))
    out = run(Xg).clone()
    # per-kernel rooflines of the eager forward (CUDA events per launch; algorithmic bytes / flops as in bench.py)
    fwd.enable_profile(True)
    for _ in range(5):
        fwd(Xg)
    torch.cuda.synchronize()
    kernels = {}
    for kname, r in fwd.profile_summary().items():
        gbs = r["bytes"] / (r["ms"] * 1e-3) / 1e9
        tfs = r["flops"] / (r["ms"] * 1e-3) / 1e12 if r["flops"] else 0.0
        tensor = kname.startswith("gemm") and 3 * tfs / TC_PEAK > gbs / HBM_PEAK
        kernels[kname] = dict(ms=round(r["ms"], 4), GBps=round(gbs, 1), TFLOPs=round(tfs, 1), bound="tensor" if tensor else "hbm",
                              frac=round(3 * tfs / TC_PEAK if tensor else gbs / HBM_PEAK, 3))
    fwd.enable_profile(False)
    q = torch.randperm(n, generator=torch.Generator().manual_seed(0))[:100].to(DEV)
    pq_ms = timeit(lambda: fg.infer.per_query(sd, pack, Xg, q, precision="bf16x3"), reps=5, warm=1)
    row = dict(config=name, mode=mode, nodes=n, subgraphs=pack.n_sub, rows=pack.n_rows, nnz=pack.nnz, F=F,
               partition_cpu_s=round(t_part, 2), pack_build_ms=round(build_ms, 2), forward_eager_ms=round(eager_ms, 4),
               forward_graph_ms=round(graph_ms, 4), nodes_per_s=n / (graph_ms * 1e-3),
               per_query_100_ms=round(pq_ms, 3), hub_rows=int(fwd.hubs_all[1]), kernels=kernels)
    # oracle on a sample of consecutive batches (none / extra): CPU time + parity
    if mode != "cluster":
        n_sub = min(pack.n_sub, sample_batches * 128)
        cos = [None if Cm is None else dict(part=fo.partition_of(Cm)[0], CX=None, adj=None) for Cm in C_list]
        subs = fo.build_subgraphs(ei, X.cpu().numpy(), np.zeros(n, dtype=np.int64), comps, cos, mode, only=set(range(n_sub)))
        subs = subs[:n_sub]
        sel = []
        for s in subs:
            m = np.zeros(s["x"].shape[0], dtype=bool)
            m[np.searchsorted(s["orig_idx"], s["core"])] = True
            sel.append(m)
        sdc = {k: v.cpu() for k, v in sd.items()}
        torch.set_num_threads(os.cpu_count())
        fo.node_infer_batched(sdc, subs[:128], sel[:128])
        t0 = time.perf_counter()
        want = fo.node_infer_batched(sdc, subs, sel, "node_cls", 128).numpy()
        t_cpu = time.perf_counter() - t0
        n_core = int(sum(m.sum() for m in sel))
        got = out[:n_core].cpu().numpy()  # core rows in pack order = subgraph order
        row.update(cpu_nodes_per_s=n_core / t_cpu, cpu_sample=f"{n_sub} subgraphs / {n_core} nodes, {os.cpu_count()} cores",
                   max_rel_err=float(np.abs(got - want).max() / max(1.0, np.abs(want).max())))
    return row


def zinc_config(n_graphs=12000, ratio=0.3, sample_graphs=256):
    graphs = fg.synth.molecule_graphs(n_graphs, seed=0)
    t0 = time.perf_counter()
    xs, eis, parts, graph_of_sub, off, sub_off = [], [], [], [], 0, 0
    per_graph = []
    for gi, (x, ei, y) in enumerate(graphs):
        n = x.shape[0]
        partition, comps, C_list = fg.synth.neighborhood_partition(ei, n, ratio, seed=gi)
        xs.append(x); eis.append(ei + off); parts.append(partition.part + sub_off)
        graph_of_sub.extend([gi] * partition.k)
        if gi < sample_graphs:
            per_graph.append((x, ei, comps, C_list))
        off += n; sub_off += partition.k
    t_part = time.perf_counter() - t0
    X = torch.tensor(np.concatenate(xs)).float().to(DEV)
    eid = torch.tensor(np.concatenate(eis, 1), device=DEV)
    part = torch.tensor(np.concatenate(parts))
    sd = fg.synth.init_state_dict(1, 512, 1, seed=0)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    pack = fg.build_pack(eid, part, sub_off, "extra")
    torch.cuda.synchronize(); build_ms = (time.perf_counter() - t0) * 1e3
    gos = torch.tensor(graph_of_sub)
    ms = timeit(lambda: fg.infer.graph_level_Gs(sd, pack, X, gos, "graph_reg", precision="bf16x3"), reps=10)
    pred = fg.infer.graph_level_Gs(sd, pack, X, gos, "graph_reg", precision="bf16x3").cpu().numpy()
    # oracle: the reference's double loop (network.py:189-204) on the first sample_graphs graphs
    set_gs, bt = [], []
    for gi, (x, ei, comps, C_list) in enumerate(per_graph):
        cos = [None if Cm is None else dict(part=fo.partition_of(Cm)[0], CX=None, adj=None) for Cm in C_list]
        subs = fo.build_subgraphs(ei, x.astype(np.float32), np.zeros(x.shape[0], dtype=np.int64), comps, cos, "extra")
        set_gs.append([dict(x=torch.tensor(s["x"]), edge_index=torch.tensor(s["edge_index"]), mask=torch.tensor(s["mask"]))
                       for s in subs])
        bt.extend([gi] * x.shape[0])
    sdc = {k: v.cpu() for k, v in sd.items()}
    t0 = time.perf_counter()
    with torch.no_grad():
        want = fo.graph_gs_forward(sdc, set_gs, torch.tensor(bt), "graph_reg").numpy()
    t_cpu = time.perf_counter() - t0
    err = float(np.abs(pred[:sample_graphs] - want).max() / max(1.0, np.abs(want).max()))
    return dict(config="zinc", mode="extra", graphs=n_graphs, nodes=off, subgraphs=pack.n_sub, rows=pack.n_rows,
                nnz=pack.nnz, F=1, partition_cpu_s=round(t_part, 2), pack_build_ms=round(build_ms, 2),
                forward_eager_ms=round(ms, 4), graphs_per_s=n_graphs / (ms * 1e-3), nodes_per_s=off / (ms * 1e-3),
                cpu_graphs_per_s=sample_graphs / t_cpu, cpu_sample=f"{sample_graphs} graphs, {os.cpu_count()} cores",
                max_rel_err=err)


if __name__ == "__main__":
    rows = []
    for name, mode in (("cora", "none"), ("cora", "extra"), ("pubmed", "extra"), ("physics", "extra"), ("physics", "cluster")):
        r = node_config(name, mode, sample_batches=1 if name == "physics" else 8)
        print(json.dumps(r), flush=True)
        rows.append(r)
    r = zinc_config()
    print(json.dumps(r), flush=True)
    rows.append(r)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "configs_r2.json"), "w"), indent=1)
    cols = ["config", "mode", "nodes", "subgraphs", "rows", "nnz", "pack_build_ms", "forward_eager_ms", "forward_graph_ms",
            "nodes_per_s", "per_query_100_ms", "cpu_nodes_per_s", "max_rel_err"]
    with open(os.path.join(ROOT, "gpurun_out", "configs_r2.md"), "w") as f:
        f.write("| " + " | ".join(cols) + " |\n|" + "---|" * len(cols) + "\n")
        for r in rows:
            f.write("| " + " | ".join(f"{r.get(c, ''):.4g}" if isinstance(r.get(c), float) else str(r.get(c, "")) for c in cols) + " |\n")
        f.write("\nPer-kernel rooflines (eager forward, CUDA events per launch; frac = of the measured copy peak / sustained bf16 peak x3):\n\n")
        for r in rows:
            if "kernels" in r:
                f.write(f"* {r['config']} / {r['mode']} (hub rows: {r.get('hub_rows')}): " + "; ".join(
                    f"{k} {v['ms']} ms {v['GBps']} GB/s {v['TFLOPs']} TF ({v['bound']} {v['frac']})" for k, v in r["kernels"].items()) + "\n")
