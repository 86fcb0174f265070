"""Generates tests/golden/node_reg_small.npz: NODE REGRESSION, by running the UNMODIFIED reference code from
/root/reference behind oracle/ref_shims.py — `coarsening_regression` with task 'node_reg' in all three modes
(utils.py:376-605, the regression twin of coarsening_classification; main.py:311), `load_data_regression`
(utils.py:780-808: random splits + per-subgraph masks), the batch loop of run.py:59-77 and `Regress_node`
(network.py:37-64).  Same layout as node_small.npz so tests/golden_io.py reads both.

    python tests/golden/make_golden_node_reg.py          # authoring container only
"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import make_golden as mg  # noqa: E402  (installs the shims, imports the reference's utils / network)
from oracle import ref_shims  # noqa: E402

ref_utils, ref_network, Data = mg.ref_utils, mg.ref_network, mg.Data


def record_components(store):
    """Capture Graph.extract_components() so the candidate order (sorted by size, descending, stable — utils.py:379)
    can be stored: the node_reg branch does not return `candidate`."""
    orig = ref_shims.Graph.extract_components

    def wrapped(self):
        comps = orig(self)
        store.append([np.asarray(c.info["orig_idx"]) for c in comps])
        return comps

    ref_shims.Graph.extract_components = wrapped
    return orig


def main(name="node_reg_small", seed=9, n_main=240, extra_components=(1, 1, 3, 6, 11, 13), F=10, ratio=0.35, hidden=32):
    n, ei = mg.synth_graph(seed, n_main, list(extra_components))
    rng = np.random.default_rng(seed + 1)
    x = rng.random((n, F), dtype=np.float32)
    y = rng.normal(size=n).astype(np.float32)
    out = dict(n=np.array(n), edge_index=ei, x=x, y=y, ratio=np.array(ratio), hidden=np.array(hidden))
    for mode in ("none", "extra", "cluster"):
        args = argparse.Namespace(task="node_reg", extra_node=mode == "extra", cluster_node=mode == "cluster",
                                  num_classes=1, num_features=F, hidden=hidden, num_layers1=2, layer_name="GCNConv",
                                  train_ratio=0.3, val_ratio=0.2)
        data = Data(x=torch.tensor(x), edge_index=torch.tensor(ei), y=torch.tensor(y))
        records, comp_store = [], []
        orig = mg.record_coarsen(records)
        orig_ec = record_components(comp_store)
        try:
            torch.manual_seed(seed); np.random.seed(seed)
            _, subgraph_list = ref_utils.coarsening_regression(args, data, 1 - ratio, "variation_neighborhoods")
        finally:
            ref_utils.coarsen = orig
            ref_shims.Graph.extract_components = orig_ec
        candidate = sorted(comp_store[0], key=len, reverse=True)  # utils.py:379 (stable)
        out[f"{mode}_ncomp"] = np.array(len(candidate))
        out[f"{mode}_comp_sizes"] = np.array([len(c) for c in candidate])
        out[f"{mode}_comp_nodes"] = np.concatenate(candidate)
        out[f"{mode}_nrec"] = np.array(len(records))
        for i, r in enumerate(records):
            mg.pack_sparse(f"{mode}_C{i}", r["C"], out)
            mg.pack_sparse(f"{mode}_W{i}", r["W"], out)
            comp_map = ref_utils.subgraph_mapping(r["maps"])
            out[f"{mode}_map{i}"] = np.array([comp_map[j] for j in range(len(r["orig_idx"]))])
            out[f"{mode}_CX{i}"] = np.asarray(r["C"].dot(x[r["orig_idx"]]))
        torch.manual_seed(seed + 3)  # splits_regression draws torch.randperm (utils.py:651)
        graphs = ref_utils.load_data_regression(args, data, subgraph_list)
        if mode == "none":
            out["train_mask"] = data.train_mask.numpy(); out["val_mask"] = data.val_mask.numpy()
            out["test_mask"] = data.test_mask.numpy()
        else:
            assert np.array_equal(out["test_mask"], data.test_mask.numpy())  # same seed -> same split in every mode
        mg.save_subgraphs(f"{mode}_sub", subgraph_list, out, with_split=graphs)
        torch.manual_seed(seed + 7)
        model = ref_network.Regress_node(args)
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
    np.savez_compressed(os.path.join(mg.OUT, name + ".npz"), **out)
    print(name, "n =", n, {m: int(out[f"{m}_sub_count"]) for m in ("none", "extra", "cluster")},
          "test_out", out["cluster_test_out"].shape)


if __name__ == "__main__":
    main()
