"""Generates tests/golden/*.npz by running the UNMODIFIED reference code from /root/reference
(graph_coarsening.coarsen, utils.coarsening_classification / coarsening_regression /
load_data_classification / load_graph_data / colater, network.Classify_node / Regress_graph_gs) behind
oracle/ref_shims.py.  Run in the authoring container only:

    python tests/golden/make_golden.py

The fixtures pin the oracle (oracle/fitgnn_oracle.py) to the reference's own outputs; /root/reference does
not exist on the GPU box, so nothing else reads it.
"""
import argparse
import os
import sys
import tempfile

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_shims  # noqa: E402

ref_shims.install()
_cwd = os.getcwd()
os.chdir(tempfile.mkdtemp())  # run.py / utils.py create ./results etc. on import
import utils as ref_utils  # noqa: E402  (reference utils.py)
import network as ref_network  # noqa: E402  (reference network.py)
os.chdir(_cwd)
from oracle.ref_shims import Data  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def synth_graph(seed, n_main, extra_components, avg_deg=4.0):
    """Seeded undirected graph: a preferential-attachment style main component plus small ones
    (sizes in extra_components: exercises the single-node and <= 10-node paths of utils.py)."""
    rng = np.random.default_rng(seed)
    edges = set()
    targets = [0]
    for v in range(1, n_main):
        m = max(1, int(rng.poisson(avg_deg / 2)))
        for u in rng.choice(targets, size=min(m, len(targets)), replace=True):
            if u != v:
                edges.add((min(u, v), max(u, v)))
        targets.extend([v] * m)
        targets.append(int(rng.integers(0, v + 1)))
    off = n_main
    for size in extra_components:
        for i in range(1, size):
            j = int(rng.integers(0, i))
            edges.add((off + j, off + i))
        off += size
    n = off
    und = np.array(sorted(edges), dtype=np.int64)
    perm = rng.permutation(n)  # shuffle node ids so components are interleaved
    und = perm[und]
    ei = np.concatenate([und, und[:, ::-1]], 0)
    ei = ei[rng.permutation(len(ei))].T  # unsorted COO, both directions
    return n, np.ascontiguousarray(ei)


def record_coarsen(records):
    """Wrap the reference coarsen() as imported into utils' namespace so each call's outputs are captured
    (eigsh start vectors are random, so a second call need not give the same partition)."""
    orig = ref_utils.coarsen

    def wrapped(H, r=0.5, method="variation_neighborhoods", **kw):
        C, Gc, maps = orig(H, r=r, method=method, **kw)
        records.append(dict(orig_idx=np.asarray(H.info["orig_idx"]), C=C.copy(), W=Gc.W.copy(),
                            maps=[dict(m) for m in maps]))
        return C, Gc, maps

    ref_utils.coarsen = wrapped
    return orig


def pack_sparse(prefix, M, out):
    M = M.tocoo()
    out[prefix + "_row"], out[prefix + "_col"], out[prefix + "_val"] = M.row, M.col, M.data
    out[prefix + "_shape"] = np.array(M.shape)


def save_subgraphs(prefix, subgraph_list, out, with_split=None):
    out[prefix + "_count"] = np.array(len(subgraph_list))
    sizes, esizes = [], []
    xs, eis, ys, masks, oidx, aext, aext_sizes, oidx_sizes, mk, mv, msz = [], [], [], [], [], [], [], [], [], [], []
    for M in subgraph_list:
        sizes.append(M.x.shape[0]); esizes.append(M.edge_index.shape[1])
        xs.append(M.x.cpu().numpy()); eis.append(M.edge_index.cpu().numpy()); ys.append(M.y.cpu().numpy())
        masks.append(M.mask.cpu().numpy())
        o = M.orig_idx.cpu().numpy(); oidx.append(o); oidx_sizes.append(len(o))
        a = M.actual_ext
        a = a.cpu().numpy() if torch.is_tensor(a) else np.asarray(a)
        a = a.reshape(-1); aext.append(a.astype(np.int64)); aext_sizes.append(len(a))
        mk.append(np.array(list(M.map_dict.keys()), dtype=np.int64))
        mv.append(np.array(list(M.map_dict.values()), dtype=np.int64)); msz.append(len(M.map_dict))
    cat = lambda L, ax=0: np.concatenate(L, ax) if L else np.zeros(0)
    out[prefix + "_sizes"] = np.array(sizes); out[prefix + "_esizes"] = np.array(esizes)
    out[prefix + "_x"] = cat(xs).astype(np.float32); out[prefix + "_ei"] = cat(eis, 1)
    out[prefix + "_y"] = cat(ys); out[prefix + "_mask"] = cat(masks)
    out[prefix + "_orig_idx"] = cat(oidx); out[prefix + "_orig_sizes"] = np.array(oidx_sizes)
    out[prefix + "_actual_ext"] = cat(aext); out[prefix + "_ext_sizes"] = np.array(aext_sizes)
    out[prefix + "_map_k"] = cat(mk); out[prefix + "_map_v"] = cat(mv); out[prefix + "_map_sizes"] = np.array(msz)
    if with_split is not None:
        out[prefix + "_train"] = cat([g.train_mask.numpy() for g in with_split])
        out[prefix + "_val"] = cat([g.val_mask.numpy() for g in with_split])
        out[prefix + "_test"] = cat([g.test_mask.numpy() for g in with_split])


def node_case(name, seed, n_main, extra_components, F, C, ratio, hidden):
    n, ei = synth_graph(seed, n_main, extra_components)
    rng = np.random.default_rng(seed + 1)
    x = rng.random((n, F), dtype=np.float32)
    x /= x.sum(1, keepdims=True)  # --normalize_features main.py:38
    y = rng.integers(0, C, n)
    perm = rng.permutation(n)
    train = np.zeros(n, bool); val = np.zeros(n, bool); test = np.zeros(n, bool)
    train[perm[: n // 5]] = True; val[perm[n // 5: n // 2]] = True; test[perm[n // 2:]] = True
    out = dict(n=np.array(n), edge_index=ei, x=x, y=y, train_mask=train, val_mask=val, test_mask=test,
               ratio=np.array(ratio), n_classes=np.array(C), hidden=np.array(hidden))
    for mode in ("none", "extra", "cluster"):
        args = argparse.Namespace(task="node_cls", extra_node=mode == "extra", cluster_node=mode == "cluster",
                                  num_classes=C, num_features=F, hidden=hidden, num_layers1=2, layer_name="GCNConv")
        data = Data(x=torch.tensor(x), edge_index=torch.tensor(ei), y=torch.tensor(y),
                    train_mask=torch.tensor(train), val_mask=torch.tensor(val), test_mask=torch.tensor(test))
        records = []
        orig = record_coarsen(records)
        try:
            torch.manual_seed(seed); np.random.seed(seed)
            _, candidate, C_list, Gc_list, subgraph_list = ref_utils.coarsening_classification(
                args, data, 1 - ratio, "variation_neighborhoods")  # r = 1 - ratio: main.py:278
        finally:
            ref_utils.coarsen = orig
        out[f"{mode}_ncomp"] = np.array(len(candidate))
        out[f"{mode}_comp_sizes"] = np.array([len(c.info["orig_idx"]) for c in candidate])
        out[f"{mode}_comp_nodes"] = np.concatenate([np.asarray(c.info["orig_idx"]) for c in candidate])
        out[f"{mode}_nrec"] = np.array(len(records))
        for i, r in enumerate(records):
            pack_sparse(f"{mode}_C{i}", r["C"], out)
            pack_sparse(f"{mode}_W{i}", r["W"], out)
            # composed mapping dicts == subgraph_mapping(mapping_dict_list) utils.py:113-121
            comp_map = ref_utils.subgraph_mapping(r["maps"])
            out[f"{mode}_map{i}"] = np.array([comp_map[j] for j in range(len(r["orig_idx"]))])
            out[f"{mode}_CX{i}"] = np.asarray(r["C"].dot(x[r["orig_idx"]]))  # utils.py:161 (float64)
        (n_classes, cf, ctl, ctm, cvl, cvm, cedge, graphs) = ref_utils.load_data_classification(
            args, data, candidate, C_list, Gc_list, "fixed", subgraph_list)
        save_subgraphs(f"{mode}_sub", subgraph_list, out, with_split=graphs)
        out[f"{mode}_gc_x"] = cf.numpy(); out[f"{mode}_gc_train_y"] = ctl.numpy(); out[f"{mode}_gc_train_m"] = ctm.numpy()
        out[f"{mode}_gc_val_y"] = cvl.numpy(); out[f"{mode}_gc_val_m"] = cvm.numpy(); out[f"{mode}_gc_edge"] = cedge.numpy()
        # model forward through the reference's network.py (GCNConv = oracle restatement, see ref_shims)
        torch.manual_seed(seed + 7)
        model = ref_network.Classify_node(args)
        with torch.no_grad():
            for p in model.parameters():
                if p.dim() == 1:
                    p.uniform_(-0.1, 0.1)
        model.eval()
        if mode == "none":
            for k, v in model.state_dict().items():
                out["sd_" + k] = v.numpy()
        else:
            model.load_state_dict({k[3:]: torch.tensor(v) for k, v in out.items() if k.startswith("sd_")})
        outs = []
        loader = ref_shims.DataLoader(graphs, batch_size=128, shuffle=False)  # run.py:336
        with torch.no_grad():
            for batch in loader:  # run.py:59-77 (test branch)
                if True in batch.test_mask:
                    o = model(batch.x, batch.edge_index)
                    outs.append(o[batch.test_mask].numpy())
        out[f"{mode}_test_out"] = np.concatenate(outs, 0)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, "n =", n, {m: int(out[f"{m}_sub_count"]) for m in ("none", "extra", "cluster")})


def graph_case(name, seed, n_graphs, F, ratio, hidden):
    """graph_reg on ZINC-shaped small graphs (config 4): main.py:370-381 + Regress_graph_gs."""
    rng = np.random.default_rng(seed)
    out = dict(n_graphs=np.array(n_graphs), ratio=np.array(ratio), hidden=np.array(hidden))
    args = argparse.Namespace(task="graph_reg", extra_node=True, cluster_node=False, num_classes=1, num_features=F,
                              hidden=hidden, num_layers1=2, layer_name="GCNConv")
    data_list = []
    kept = 0
    for g in range(n_graphs):
        n, ei = synth_graph(seed * 1000 + g, int(rng.integers(12, 30)), [1] if g % 3 == 0 else [])
        x = rng.integers(0, 21, (n, F)).astype(np.int64)  # ZINC atom types (int64 -> .float() network.py:195)
        y = rng.normal(size=(1,)).astype(np.float32)
        graph = Data(x=torch.tensor(x), edge_index=torch.tensor(ei), y=torch.tensor(y))
        records = []
        orig = record_coarsen(records)
        try:
            _, candidate, subgraph_list, CLIST, GcLIST = ref_utils.coarsening_regression(
                args, graph, 1 - ratio, "variation_neighborhoods")
        finally:
            ref_utils.coarsen = orig
        Gc = ref_utils.load_graph_data(graph, CLIST, GcLIST, candidate)
        p = f"g{kept}"
        out[p + "_n"] = np.array(n); out[p + "_ei"] = ei; out[p + "_x"] = x; out[p + "_y"] = y
        out[p + "_comp_sizes"] = np.array([len(c.info["orig_idx"]) for c in candidate])
        out[p + "_comp_nodes"] = np.concatenate([np.asarray(c.info["orig_idx"]) for c in candidate])
        out[p + "_nrec"] = np.array(len(records))
        for i, r in enumerate(records):
            pack_sparse(f"{p}_C{i}", r["C"], out)
            pack_sparse(f"{p}_W{i}", r["W"], out)
        save_subgraphs(p + "_sub", subgraph_list, out)
        out[p + "_gc_x"] = Gc.x.numpy(); out[p + "_gc_edge"] = Gc.edge_index.numpy()
        data_list.append([graph, Gc, subgraph_list])
        kept += 1
    out["n_kept"] = np.array(kept)
    torch.manual_seed(seed + 7)
    model = ref_network.Regress_graph_gs(args)
    with torch.no_grad():
        for p_ in model.parameters():
            if p_.dim() == 1:
                p_.uniform_(-0.1, 0.1)
    model.eval()
    for k, v in model.state_dict().items():
        out["sd_" + k] = v.numpy()
    GC_batch, GS, Y, batch_tensor = ref_utils.colater()(data_list)  # utils.py:893-908
    with torch.no_grad():
        pred = model(GS, batch_tensor)  # network.py:189-204
        model_gc = ref_network.Regress_graph_gc(args)
        model_gc.load_state_dict(model.state_dict())
        model_gc.eval()
        GC_batch.x = GC_batch.x.float()
        pred_gc = model_gc(GC_batch)  # network.py:158-166
    out["batch_tensor"] = batch_tensor.numpy(); out["pred_gs"] = pred.numpy(); out["pred_gc"] = pred_gc.numpy()
    out["gc_batch"] = GC_batch.batch.numpy()
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, "graphs =", kept, "pred", pred.shape)


if __name__ == "__main__":
    node_case("node_small", seed=3, n_main=260, extra_components=[1, 1, 2, 5, 9, 12, 14], F=12, C=4, ratio=0.3,
              hidden=32)
    node_case("node_mid", seed=5, n_main=900, extra_components=[1, 3, 11], F=20, C=5, ratio=0.5, hidden=64)
    graph_case("graph_small", seed=11, n_graphs=9, F=1, ratio=0.3, hidden=32)
