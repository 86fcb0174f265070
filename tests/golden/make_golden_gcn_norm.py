"""Generates tests/golden/gcn_norm_gcond.npz: the symmetric GCN normalisation and a dense GCN layer as implemented
INSIDE THE REFERENCE TREE, executed unmodified from /root/reference:

  * `normalize_adj`        Baselines/GCOND/models/mycheby.py:393-414   A' = (D+I)^-1/2 (A+I) (D+I)^-1/2  (scipy)
  * `GraphConvolution`     Baselines/GCOND/models/gcn.py:15-52         out = A'·(X·W) + b                 (torch CPU)

Both modules import packages that are not installed (deeprobust, torch_sparse, torch_geometric), so the two
definitions are cut out of the files with `ast` and executed on their own — no reference source is copied into this
repo.  The hot path's GCNConv (network.py:31) lives in torch_geometric, which the reference does not vendor; on SIMPLE
graphs (undirected, no self loops, no duplicate edges) PyG's gcn_norm is exactly this normalisation, which is what the
fixture pins.  PyG's handling of duplicates / existing self loops stays a restatement (oracle/fitgnn_oracle.py header).

    python tests/golden/make_golden_gcn_norm.py          # authoring container only
"""
import ast
import math
import os
import types

import numpy as np
import scipy.sparse as sp
import torch

REF = "/root/reference/Baselines/GCOND/models"
OUT = os.path.dirname(os.path.abspath(__file__))


def cut(path, name, namespace):
    tree = ast.parse(open(path).read())
    node = next(n for n in tree.body if isinstance(n, (ast.FunctionDef, ast.ClassDef)) and n.name == name)
    exec(compile(ast.Module(body=[node], type_ignores=[]), path, "exec"), namespace)
    return namespace[name]


def main():
    normalize_adj = cut(os.path.join(REF, "mycheby.py"), "normalize_adj", {"sp": sp, "np": np})
    ns = {"torch": torch, "math": math, "Module": torch.nn.Module, "Parameter": torch.nn.Parameter,
          "torch_sparse": types.SimpleNamespace(SparseTensor=type("SparseTensor", (), {}))}
    GraphConvolution = cut(os.path.join(REF, "gcn.py"), "GraphConvolution", ns)

    rng = np.random.default_rng(17)
    n, F, H = 97, 11, 8
    und = set()
    while len(und) < 260:
        a, b = (int(v) for v in rng.integers(0, n - 3, 2))  # the last three nodes stay isolated
        if a != b:
            und.add((min(a, b), max(a, b)))
    und = np.array(sorted(und), dtype=np.int64)
    ei = np.concatenate([und, und[:, ::-1]], 0)
    ei = np.ascontiguousarray(ei[rng.permutation(len(ei))].T)  # unsorted COO, both directions
    A = sp.csr_matrix((np.ones(ei.shape[1]), (ei[1], ei[0])), shape=(n, n))
    A_norm = np.asarray(normalize_adj(A).todense(), dtype=np.float64)

    torch.manual_seed(3)
    layer = GraphConvolution(F, H)
    X = torch.rand(n, F)
    with torch.no_grad():
        out = layer(X, torch.tensor(A_norm, dtype=torch.float32).to_sparse())
    np.savez_compressed(os.path.join(OUT, "gcn_norm_gcond.npz"), n=n, edge_index=ei, A_norm=A_norm, X=X.numpy(),
                        weight_in_out=layer.weight.detach().numpy(), bias=layer.bias.detach().numpy(), out=out.numpy())
    print("gcn_norm_gcond: n =", n, "E =", ei.shape[1], "nnz(A') =", int((A_norm != 0).sum()), "out", tuple(out.shape))


if __name__ == "__main__":
    main()
