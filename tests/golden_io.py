"""Readers for tests/golden/*.npz (written by tests/golden/make_golden.py from the reference's own code)."""
import os

import numpy as np
import scipy.sparse as sp

HERE = os.path.dirname(os.path.abspath(__file__))


def load(name):
    return np.load(os.path.join(HERE, "golden", name + ".npz"))


def sparse(d, prefix):
    return sp.coo_matrix((d[prefix + "_val"], (d[prefix + "_row"], d[prefix + "_col"])),
                         shape=tuple(d[prefix + "_shape"]))


def split(arr, sizes, axis=0):
    idx = np.cumsum(sizes)[:-1]
    return np.split(arr, idx, axis=axis)


def components(d, prefix):
    return split(d[prefix + "_comp_nodes"], d[prefix + "_comp_sizes"])


def records(d, prefix):
    """Per component with > 1 node, in candidate order: C (csc), W = Gc.W (csr), map (comp node -> supernode), CX."""
    out = []
    for i in range(int(d[prefix + "_nrec"])):
        rec = dict(C=sparse(d, f"{prefix}_C{i}").tocsc(), W=sparse(d, f"{prefix}_W{i}").tocsr())
        if f"{prefix}_map{i}" in d.files:
            rec["map"] = d[f"{prefix}_map{i}"]
            rec["CX"] = d[f"{prefix}_CX{i}"]
        out.append(rec)
    return out


def subgraphs(d, prefix):
    """The reference's subgraph_list as dicts (x, edge_index, y, mask, orig_idx, actual_ext, map_dict)."""
    n = int(d[prefix + "_count"])
    sizes, esizes = d[prefix + "_sizes"], d[prefix + "_esizes"]
    xs = split(d[prefix + "_x"], sizes)
    eis = split(d[prefix + "_ei"], esizes, axis=1)
    ys = split(d[prefix + "_y"], sizes)
    masks = split(d[prefix + "_mask"], sizes)
    oidx = split(d[prefix + "_orig_idx"], d[prefix + "_orig_sizes"])
    aext = split(d[prefix + "_actual_ext"], d[prefix + "_ext_sizes"])
    mk = split(d[prefix + "_map_k"], d[prefix + "_map_sizes"])
    mv = split(d[prefix + "_map_v"], d[prefix + "_map_sizes"])
    out = []
    for i in range(n):
        s = dict(x=xs[i], edge_index=eis[i].astype(np.int64), y=ys[i], mask=masks[i].astype(bool),
                 orig_idx=oidx[i].astype(np.int64), actual_ext=aext[i].astype(np.int64),
                 map_dict={int(a): int(b) for a, b in zip(mk[i], mv[i])})
        out.append(s)
    if prefix + "_train" in d.files:
        for s, tr, va, te in zip(out, split(d[prefix + "_train"], sizes), split(d[prefix + "_val"], sizes),
                                 split(d[prefix + "_test"], sizes)):
            s["train_mask"], s["val_mask"], s["test_mask"] = tr.astype(bool), va.astype(bool), te.astype(bool)
    return out


def state_dict(d):
    import torch
    return {k[3:]: torch.tensor(d[k]) for k in d.files if k.startswith("sd_")}


def coarsenings_for_oracle(d, prefix, comps):
    """The oracle builder's per-component inputs, taken from the reference's recorded coarsen() outputs."""
    from oracle import fitgnn_oracle as fo
    recs = records(d, prefix)
    out, j = [], 0
    for comp in comps:
        if len(comp) > 1:
            r = recs[j]; j += 1
            part, _ = fo.partition_of(r["C"])
            out.append(dict(part=part, CX=r["CX"], adj=(r["W"] > 0).tocsr(), C=r["C"], W=r["W"]))
        else:
            out.append(None)
    return out
