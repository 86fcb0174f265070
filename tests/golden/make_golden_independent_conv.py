"""Generates tests/golden/independent_conv.npz: the reference's model classes (network.py) run on the fixture subgraph lists
with a GCNConv whose ARITHMETIC comes from the reference tree itself instead of from the oracle: the layer is assembled from
`normalize_adj` (Baselines/GCOND/models/mycheby.py:393-414) and the forward of `GraphConvolution`
(Baselines/GCOND/models/gcn.py:15-52), both cut out of the unmodified files with `ast`.  The other fixtures run network.py
on oracle/ref_shims.py's GCNConv, which calls the oracle's own restatement — a structural pin only.  Here nothing of the
oracle is on the path, so agreement of the two (tests/test_oracle_golden.py) pins the oracle's GCNConv arithmetic on every
subgraph of the fixtures (simple graphs: no duplicate edges, no self loops).

    python tests/golden/make_golden_independent_conv.py          # authoring container only
"""
import argparse
import ast
import math
import os
import sys
import types

import numpy as np
import scipy.sparse as sp
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import make_golden as mg  # noqa: E402
from oracle import ref_shims  # noqa: E402
from tests import golden_io as gio  # noqa: E402

REF = "/root/reference/Baselines/GCOND/models"


def cut(path, name, namespace):
    node = next(n for n in ast.parse(open(path).read()).body
                if isinstance(n, (ast.FunctionDef, ast.ClassDef)) and n.name == name)
    exec(compile(ast.Module(body=[node], type_ignores=[]), path, "exec"), namespace)
    return namespace[name]


normalize_adj = cut(os.path.join(REF, "mycheby.py"), "normalize_adj", {"sp": sp, "np": np})
GraphConvolution = cut(os.path.join(REF, "gcn.py"), "GraphConvolution",
                       {"torch": torch, "math": math, "Module": torch.nn.Module, "Parameter": torch.nn.Parameter,
                        "torch_sparse": types.SimpleNamespace(SparseTensor=type("SparseTensor", (), {}))})


class TreeGCNConv(torch.nn.Module):
    """PyG GCNConv's parameter surface (lin.weight [out, in], bias) in front of the reference tree's own operator."""

    def __init__(self, in_channels, out_channels, **kw):
        super().__init__()
        self.lin = torch.nn.Linear(in_channels, out_channels, bias=False)
        self.bias = torch.nn.Parameter(torch.zeros(out_channels))
        self.inner = GraphConvolution(in_channels, out_channels)  # its own weight / bias are replaced per call

    def forward(self, x, edge_index):
        n = x.shape[0]
        ei = edge_index.numpy()
        assert (ei[0] != ei[1]).all() and len({(int(a), int(b)) for a, b in ei.T}) == ei.shape[1], "not a simple graph"
        A = sp.csr_matrix((np.ones(ei.shape[1]), (ei[1], ei[0])), shape=(n, n))  # row = target, as PyG aggregates
        A_norm = torch.tensor(np.asarray(normalize_adj(A).todense()), dtype=torch.float32).to_sparse()
        self.inner.weight = torch.nn.Parameter(self.lin.weight.t().contiguous())  # the layer stores W as [in, out]
        self.inner.bias = self.bias
        return self.inner(x, A_norm)

    def state_dict(self, *a, **k):  # only lin.weight / bias are checkpoint keys
        sd = super().state_dict(*a, **k)
        return {key: v for key, v in sd.items() if "inner" not in key}


def main():
    out = {}
    tg_nn = sys.modules["torch_geometric.nn"]
    shim_conv = tg_nn.GCNConv
    tg_nn.GCNConv = TreeGCNConv  # network.py resolves the layer with getattr(pyg_nn, args.layer_name) at construction
    try:
        for case, Model, task in (("node_small", mg.ref_network.Classify_node, "node_cls"),
                                  ("node_reg_small", mg.ref_network.Regress_node, "node_reg")):
            d = gio.load(case)
            sd = gio.state_dict(d)
            for mode in ("none", "extra", "cluster"):
                args = argparse.Namespace(task=task, num_classes=int(d["n_classes"]) if "n_classes" in d.files else 1,
                                          num_features=d["x"].shape[1], hidden=int(d["hidden"]), num_layers1=2,
                                          layer_name="GCNConv")
                model = Model(args)
                model.load_state_dict(sd, strict=False)
                model.eval()
                outs = []
                with torch.no_grad():
                    for s in gio.subgraphs(d, mode + "_sub"):  # one subgraph at a time: dense normalisation per call
                        o = model(torch.tensor(s["x"]), torch.tensor(s["edge_index"]))
                        outs.append(o[torch.tensor(s["test_mask"])].numpy())
                out[f"{case}_{mode}_test_out"] = np.concatenate(outs, 0)
    finally:
        tg_nn.GCNConv = shim_conv
    np.savez_compressed(os.path.join(mg.OUT, "independent_conv.npz"), **out)
    d = gio.load("node_small")
    print("independent_conv:", {k: v.shape for k, v in out.items()},
          "max |diff| vs shim-based fixture:", float(np.abs(out["node_small_extra_test_out"] - d["extra_test_out"]).max()))


if __name__ == "__main__":
    main()
