"""Generates tests/golden/graph_cls_small.npz: graph CLASSIFICATION on small graphs, by running the UNMODIFIED reference
code from /root/reference behind oracle/ref_shims.py — `coarsening_classification` with task 'graph_cls' per graph
(main.py:336-347), `load_graph_data`, `colater` (utils.py:893-908), and the models `Classify_graph_gs`
(network.py:97-135: max pool + softmax) and `Classify_graph_gc` (network.py:66-95).  Complements graph_small.npz
(graph regression, extra_node); this case uses cluster_node.

    python tests/golden/make_golden_graph_cls.py          # authoring container only
"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import make_golden as mg  # noqa: E402  (installs the shims, imports the reference's utils / network)

ref_utils, ref_network, Data = mg.ref_utils, mg.ref_network, mg.Data


def main(name="graph_cls_small", seed=23, n_graphs=8, F=5, C=3, ratio=0.4, hidden=16):
    rng = np.random.default_rng(seed)
    out = dict(n_graphs=np.array(n_graphs), ratio=np.array(ratio), hidden=np.array(hidden), n_classes=np.array(C))
    args = argparse.Namespace(task="graph_cls", extra_node=False, cluster_node=True, num_classes=C, num_features=F,
                              hidden=hidden, num_layers1=2, layer_name="GCNConv")
    data_list = []
    for g in range(n_graphs):
        n, ei = mg.synth_graph(seed * 1000 + g, int(rng.integers(14, 32)), [1] if g % 4 == 1 else [])
        x = rng.random((n, F)).astype(np.float32)
        y = np.array([int(rng.integers(0, C))], dtype=np.int64)
        graph = Data(x=torch.tensor(x), edge_index=torch.tensor(ei), y=torch.tensor(y))
        _, candidate, subgraph_list, CLIST, GcLIST = ref_utils.coarsening_classification(
            args, graph, 1 - ratio, "variation_neighborhoods")
        Gc = ref_utils.load_graph_data(graph, CLIST, GcLIST, candidate)
        p = f"g{g}"
        out[p + "_n"] = np.array(n); out[p + "_ei"] = ei; out[p + "_x"] = x; out[p + "_y"] = y
        mg.save_subgraphs(p + "_sub", subgraph_list, out)
        out[p + "_gc_x"] = Gc.x.numpy(); out[p + "_gc_edge"] = Gc.edge_index.numpy()
        data_list.append([graph, Gc, subgraph_list])
    out["n_kept"] = np.array(n_graphs)
    torch.manual_seed(seed + 7)
    model = ref_network.Classify_graph_gs(args)
    with torch.no_grad():
        for p_ in model.parameters():
            if p_.dim() == 1:
                p_.uniform_(-0.1, 0.1)
    model.eval()
    for k, v in model.state_dict().items():
        out["sd_" + k] = v.numpy()
    GC_batch, GS, Y, batch_tensor = ref_utils.colater()(data_list)
    with torch.no_grad():
        pred = model(GS, batch_tensor)  # network.py:118-135
        model_gc = ref_network.Classify_graph_gc(args)
        model_gc.load_state_dict(model.state_dict())
        model_gc.eval()
        GC_batch.x = GC_batch.x.float()
        pred_gc = model_gc(GC_batch)  # network.py:87-95
    out["batch_tensor"] = batch_tensor.numpy(); out["pred_gs"] = pred.numpy(); out["pred_gc"] = pred_gc.numpy()
    out["gc_batch"] = GC_batch.batch.numpy()
    np.savez_compressed(os.path.join(mg.OUT, name + ".npz"), **out)
    print(name, "graphs =", n_graphs, "pred_gs", tuple(pred.shape), "pred_gc", tuple(pred_gc.shape),
          "rows sum to 1:", bool(torch.allclose(pred.sum(1), torch.ones(n_graphs))))


if __name__ == "__main__":
    main()
