"""ORACLE — TEST INFRASTRUCTURE ONLY.  CPU restatement (numpy / scipy / torch-CPU) of FIT-GNN's hot path.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may
import this module; the product package `fitgnn_b200` never does.

Every function cites the reference lines (under /root/reference) it follows.

PARITY PINNING
  * coarsening matrix C, partition, coarsened adjacency Ac, C·X, the subgraph lists (none / extra_node /
    cluster_node), per-subgraph split masks and the Gc assembly are pinned against the reference's OWN
    code, executed unmodified from /root/reference behind `oracle/ref_shims.py`
    (fixtures: tests/golden/*.npz, generators: tests/golden/make_golden*.py, index: tests/golden/README.md);
    so are the drivers of run.py (node_infer_Gs_GD / _MB, graph_infer_Gs, node_train_Gs_GD, node_train_Gc).
  * The GCNConv arithmetic lives in torch_geometric, which is not installed and not vendored by the
    reference (requirements.txt:2, unpinned; code needs PyG >= 2.1).  `gcn_norm` / `gcn_conv` restate the
    published PyG algorithm.  They are pinned on SIMPLE graphs (undirected, no self loops, no duplicates,
    isolated nodes included) against the reference tree's own implementation of the same operator —
    `normalize_adj` (Baselines/GCOND/models/mycheby.py:393-414) and the dense `GraphConvolution` layer
    (Baselines/GCOND/models/gcn.py:15-52), executed unmodified (tests/golden/make_golden_gcn_norm.py ->
    tests/golden/gcn_norm_gcond.npz), on a seeded graph and, through network.py's own model classes, on every
    subgraph of the node fixtures in all three modes (tests/golden/make_golden_independent_conv.py ->
    independent_conv.npz: no oracle code on that path) — and checked against the hand-computed known-answer
    vector of SURVEY.md §8c.  PyG's treatment of duplicate edges (counted twice) and of pre-existing self loops
    (replaced by one) has no executable counterpart in the reference tree: "parity unpinned" for those two
    rules only.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp
import torch
import torch.nn.functional as F

# ============================================================================================
# a1  GCNConv  (torch_geometric.nn.GCNConv; call sites network.py:31,60,90,126,161,197)
# ============================================================================================


def gcn_norm(edge_index: torch.Tensor, num_nodes: int, dtype=torch.float32):
    """PyG gcn_norm(add_self_loops=True, improved=False, flow='source_to_target'):
    add_remaining_self_loops (existing self loops are dropped and exactly one weight-1 loop per node
    is appended), deg = scatter_add(w, col), w = deg^-1/2[row] * w * deg^-1/2[col], inf -> 0.
    Duplicate non-loop edges are kept and therefore count twice."""
    row, col = edge_index[0], edge_index[1]
    keep = row != col
    loop = torch.arange(num_nodes, dtype=row.dtype)
    row = torch.cat([row[keep], loop])
    col = torch.cat([col[keep], loop])
    w = torch.ones(row.numel(), dtype=dtype)
    deg = torch.zeros(num_nodes, dtype=dtype).scatter_add_(0, col, w)
    dinv = deg.pow(-0.5)
    dinv[dinv == float("inf")] = 0
    return row, col, dinv[row] * w * dinv[col]


def gcn_conv_torch(x, edge_index, weight, bias):
    """GCNConv.forward: x' = lin(x) (no bias); out[col] += w * x'[row]; out += bias.  fp32 torch CPU."""
    n = x.shape[0]
    row, col, w = gcn_norm(edge_index, n, x.dtype)
    xw = F.linear(x, weight)
    out = torch.zeros(n, weight.shape[0], dtype=x.dtype).index_add_(0, col, w.view(-1, 1) * xw[row])
    if bias is not None:
        out = out + bias
    return out


def normalized_adjacency_dense(edge_index, n):
    """fp64 dense restatement  Â = D^-1/2 (A + I) D^-1/2  with A[col,row] = multiplicity (target rows)."""
    A = np.zeros((n, n), dtype=np.float64)
    r, c = np.asarray(edge_index[0]), np.asarray(edge_index[1])
    for s, t in zip(r, c):
        if s != t:
            A[t, s] += 1.0
    A += np.eye(n)
    d = A.sum(axis=1)
    dinv = np.where(d > 0, d ** -0.5, 0.0)
    return dinv[:, None] * A * dinv[None, :]


def gcn_conv_fp64(x, edge_index, weight, bias):
    A = normalized_adjacency_dense(edge_index, x.shape[0])
    out = A @ (np.asarray(x, np.float64) @ np.asarray(weight, np.float64).T)
    return out + (0 if bias is None else np.asarray(bias, np.float64))


# ============================================================================================
# a2-a4  model forwards (network.py)
# ============================================================================================


def _conv_stack(sd, x, edge_index, num_layers):
    # network.py:30-33 (same body :59-62, :89-92, :125-128, :160-163, :196-199); eval: dropout = identity
    for i in range(num_layers):
        x = gcn_conv_torch(x, edge_index, sd[f"conv.{i}.lin.weight"], sd[f"conv.{i}.bias"])
        x = F.elu(x)
    return x


def num_layers_of(sd):
    return len({k.split(".")[1] for k in sd if k.startswith("conv.")})


def classify_node(sd, x, edge_index):
    """Classify_node.forward network.py:29-35 (eval mode)."""
    x = _conv_stack(sd, x, edge_index, num_layers_of(sd))
    x = F.linear(x, sd["lt1.weight"], sd["lt1.bias"])
    return F.log_softmax(x, dim=1)


def regress_node(sd, x, edge_index):
    """Regress_node.forward network.py:58-64."""
    x = _conv_stack(sd, x, edge_index, num_layers_of(sd))
    return F.linear(x, sd["lt1.weight"], sd["lt1.bias"])


def _pool(x, batch, kind, size):
    if kind == "max":  # global_max_pool network.py:93,131
        out = torch.full((size, x.shape[1]), float("-inf"), dtype=x.dtype)
        out = out.scatter_reduce(0, batch.view(-1, 1).expand_as(x), x, reduce="amax", include_self=True)
        return torch.where(torch.isinf(out), torch.zeros_like(out), out)
    s = torch.zeros((size, x.shape[1]), dtype=x.dtype).index_add_(0, batch, x)  # global_mean_pool :164,:202
    cnt = torch.zeros(size, dtype=x.dtype).index_add_(0, batch, torch.ones(batch.numel(), dtype=x.dtype))
    return s / cnt.clamp(min=1).view(-1, 1)


def graph_gs_forward(sd, set_gs, batch_tensor, task):
    """Classify_graph_gs.forward network.py:118-135 / Regress_graph_gs.forward :189-204:
    per graph, per subgraph conv stack -> x[mask] -> cat -> pool over batch_tensor -> lt1 -> softmax/identity.
    set_gs: list (graphs) of list (subgraphs) of dicts with x, edge_index, mask."""
    L = num_layers_of(sd)
    rows = []
    for gs in set_gs:
        for g in gs:
            x = _conv_stack(sd, g["x"].float(), g["edge_index"], L)
            rows.append(x[g["mask"]])
    X_main = torch.cat(rows, 0) if rows else torch.zeros(0, sd["lt1.weight"].shape[1])
    bt = batch_tensor.to(torch.int64)
    size = int(bt.max()) + 1 if bt.numel() else 0
    x = _pool(X_main, bt, "max" if task == "graph_cls" else "mean", size)
    x = F.linear(x, sd["lt1.weight"], sd["lt1.bias"])
    return F.softmax(x, dim=1) if task == "graph_cls" else x


def graph_gc_forward(sd, x, edge_index, batch, task):
    """Classify_graph_gc.forward network.py:87-95 / Regress_graph_gc.forward :158-166."""
    x = _conv_stack(sd, x, edge_index, num_layers_of(sd))
    size = int(batch.max()) + 1
    x = _pool(x, batch, "max" if task == "graph_cls" else "mean", size)
    x = F.linear(x, sd["lt1.weight"], sd["lt1.bias"])
    return F.softmax(x, dim=1) if task == "graph_cls" else x


def init_state_dict(num_features, hidden, num_classes, num_layers=2, seed=0, bias_scale=0.1):
    """Seeded parameters with the reference's state_dict keys.  PyG initialises lin.weight
    glorot-uniform and bias zero; tests use non-zero biases so the bias paths are exercised."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    dims = [num_features] + [hidden] * num_layers
    for i in range(num_layers):
        a = (6.0 / (dims[i] + dims[i + 1])) ** 0.5
        sd[f"conv.{i}.lin.weight"] = (torch.rand(dims[i + 1], dims[i], generator=g) * 2 - 1) * a
        sd[f"conv.{i}.bias"] = (torch.rand(dims[i + 1], generator=g) * 2 - 1) * bias_scale
    a = (1.0 / hidden) ** 0.5
    sd["lt1.weight"] = (torch.rand(num_classes, hidden, generator=g) * 2 - 1) * a
    sd["lt1.bias"] = (torch.rand(num_classes, generator=g) * 2 - 1) * a
    return sd


# ============================================================================================
# a10 / a11  projection math (graph_coarsening/coarsening_utils.py, graph_utils.py)
# ============================================================================================


def get_coarsening_matrix(N, partitioning):
    """coarsening_utils.py:212-254: C = I with row subgraph[0] set to 1/sqrt(nc) on the members and the
    other members' rows deleted -> CSC [N - sum(nc-1), N]; surviving rows keep their relative order."""
    C = sp.eye(N, format="lil")
    rows_to_delete = []
    for subgraph in partitioning:
        nc = len(subgraph)
        C[subgraph[0], subgraph] = 1 / np.sqrt(nc)
        rows_to_delete.extend(subgraph[1:])
    keep = np.setdiff1d(np.arange(N), np.asarray(rows_to_delete, dtype=np.int64))
    return sp.csc_matrix(C.tocsr()[keep, :])


def coarsen_matrix(W, C):
    """coarsening_utils.py:201-205: Pinv = (C·diag(1/colsum(C)))^T ; Wc = Pinv^T · W · Pinv."""
    D = sp.diags(np.array(1 / np.sum(C, 0))[0])
    Pinv = (C.dot(D)).T
    return (Pinv.T).dot(W.dot(Pinv))


def zero_diag(A):
    """graph_utils.py:79-87."""
    return A - sp.dia_matrix((A.diagonal()[np.newaxis, :], [0]), shape=A.shape)


def coarsen_from_levels(W, levels):
    """The structural part of coarsen() coarsening_utils.py:130-139 for given per-level contraction lists:
    C = iC·C, Wc = zero_diag(coarsen_matrix(W, iC)), Wc = (Wc + Wc^T)/2.  Returns (C, Wc)."""
    N = W.shape[0]
    C = sp.eye(N, format="csc")
    Wc = sp.csr_matrix(W, dtype=np.float64)
    for partitioning in levels:
        iC = get_coarsening_matrix(Wc.shape[0], partitioning)
        C = iC.dot(C)
        Wc = zero_diag(coarsen_matrix(Wc, iC))
        Wc = (Wc + Wc.T) / 2
    return sp.csc_matrix(C), sp.csr_matrix(Wc)


def partition_of(C):
    """comp node -> supernode: the row of the unique non-zero in each column of C, which equals
    subgraph_mapping(mapping_dict_list) utils.py:113-121 (asserted against the reference in the goldens)."""
    C = sp.csc_matrix(C)
    assert np.all(np.diff(C.indptr) == 1), "C must have exactly one non-zero per column"
    return C.indices.astype(np.int64), C.data.astype(np.float64)


def project_features(C, X):
    """utils.py:161,738,827: torch.FloatTensor(C.dot(H_features)) — float64 accumulate, cast to fp32."""
    return np.asarray(C.dot(np.asarray(X, dtype=np.float32).astype(np.float64))).astype(np.float32)


def project_adj_pattern(edge_index, part, k):
    """Pattern (+ integer multiplicities) of zero_diag(P_bin·A·P_bin^T) in row-major sorted COO — what
    Gc.W.tocoo() yields at utils.py:745-746 (values there are float64 within 1 ulp of these integers)."""
    a = part[np.asarray(edge_index[0])]
    b = part[np.asarray(edge_index[1])]
    keep = a != b
    M = sp.coo_matrix((np.ones(int(keep.sum()), dtype=np.int64), (a[keep], b[keep])), shape=(k, k)).tocsr()
    M.sum_duplicates()
    M.sort_indices()
    coo = M.tocoo()
    return coo.row.astype(np.int64), coo.col.astype(np.int64), coo.data.astype(np.int64)


# ============================================================================================
# a8 / a9  subgraph builder (utils.py:143-374; regression twin :376-605)
# ============================================================================================


class _Adj:
    """neighbour() utils.py:52-56 and nodes_2_neighbours() utils.py:58-62 without the O(E) rescans:
    out-neighbour lists in original edge order."""

    def __init__(self, edge_index, n):
        src = np.asarray(edge_index[0])
        dst = np.asarray(edge_index[1])
        order = np.argsort(src, kind="stable")
        self.dst = dst[order]
        self.ptr = np.zeros(n + 1, dtype=np.int64)
        np.add.at(self.ptr, src + 1, 1)
        self.ptr = np.cumsum(self.ptr)

    def neighbour(self, node):
        return self.dst[self.ptr[node]:self.ptr[node + 1]]

    def nodes_2_neighbours(self, nodes):
        if len(nodes) == 0:
            return np.zeros(0, dtype=np.int64)
        return np.unique(np.concatenate([self.neighbour(v) for v in nodes]))


def induced_subgraph(edge_index, subset, num_nodes):
    """Data.subgraph(value) utils.py:248: edges with both ends in `subset`, original order, relabelled to the
    position in `subset` (which the reference has sorted ascending, utils.py:243)."""
    node_mask = np.zeros(num_nodes, dtype=bool)
    node_mask[subset] = True
    src, dst = np.asarray(edge_index[0]), np.asarray(edge_index[1])
    em = node_mask[src] & node_mask[dst]
    relabel = np.zeros(num_nodes, dtype=np.int64)
    relabel[subset] = np.arange(len(subset))
    return np.stack([relabel[src[em]], relabel[dst[em]]])


def extract_components(edge_index, n):
    """pygsp Graph.extract_components (copy at utils.py:73-104): components in order of their smallest
    unvisited node, each as a sorted node list; then sorted by size descending (stable) utils.py:146."""
    A = sp.coo_matrix((np.ones(edge_index.shape[1]), (np.asarray(edge_index[0]), np.asarray(edge_index[1]))),
                      shape=(n, n)).tocsr()
    ncomp, labels = sp.csgraph.connected_components(A, directed=False)
    first = np.full(ncomp, n, dtype=np.int64)
    np.minimum.at(first, labels, np.arange(n))
    comps = [np.nonzero(labels == c)[0] for c in np.argsort(first, kind="stable")]
    return sorted(comps, key=lambda c: len(c), reverse=True)


def build_subgraphs(edge_index, x, y, comps, coarsenings, mode, only=None):
    """coarsening_classification utils.py:143-374 (node_cls branch :186-267; the other task branch :269-350
    and coarsening_regression :417-584 are the same body).

    comps:       list of sorted node-id arrays (candidate[i].info['orig_idx'])
    coarsenings: per component with > 1 node: dict(part=comp node -> supernode (utils.py:182),
                 CX = C.dot(H_feature) (utils.py:161), adj = Gc.A as scipy sparse (utils.py:160)); None for
                 single-node components (utils.py:352-368)
    mode:        'none' | 'extra' | 'cluster'   (args.extra_node / args.cluster_node after arg_correction
                 main.py:117-121)
    Returns list of dicts: x, edge_index, y, mask, orig_idx, actual_ext, map_dict, n_real, cluster_ids
    (cluster_ids = global subgraph index of every cluster node, which the reference leaves implicit).
    only: optional set of subgraph_list indices to build (the others become None) — for sampling big graphs."""
    edge_index = np.asarray(edge_index)
    n = x.shape[0]
    adjl = _Adj(edge_index, n)
    x = np.asarray(x)
    y = np.asarray(y)
    out = []
    for comp, co in zip(comps, coarsenings):
        comp = np.asarray(comp, dtype=np.int64)
        base = len(out)  # index of this component's first subgraph in subgraph_list
        if len(comp) > 1:
            node_2_comp_node = {int(v): i for i, v in enumerate(comp)}  # orig_to_new_map utils.py:106-111
            part = co["part"]
            # metanode_to_node_mapping_new utils.py:123-130: dict in first-seen order over comp nodes
            meta_node_2_node = {}
            for comp_node in range(len(comp)):
                meta_node_2_node.setdefault(int(part[comp_node]), []).append(int(comp[comp_node]))
            meta_order = {m: i for i, m in enumerate(meta_node_2_node.keys())}
            for key, value in meta_node_2_node.items():
                if only is not None and len(out) not in only:
                    out.append(None)
                    continue
                value = np.sort(np.asarray(value, dtype=np.int64))
                num_nodes = len(value)
                actual_ext = np.zeros(0, dtype=np.int64)
                new_edges, new_features, cluster_ids = [], [], []
                if mode == "cluster":  # utils.py:190-234
                    node_2_subgraph_node = {int(v): i for i, v in enumerate(value)}
                    meta_node_2_new_node = {}
                    for node in value:
                        N_node = adjl.neighbour(int(node))
                        Nt_node = N_node[~np.isin(N_node, value)]
                        connected = np.unique([part[node_2_comp_node[int(v)]] for v in Nt_node]) \
                            if len(Nt_node) else np.zeros(0, dtype=np.int64)  # neighbor_2_cluster utils.py:64-71
                        for cluster in connected:
                            cluster = int(cluster)
                            if cluster not in meta_node_2_new_node:
                                meta_node_2_new_node[cluster] = num_nodes
                                new_features.append(co["CX"][cluster])
                                cluster_ids.append(base + meta_order[cluster])
                                num_nodes += 1
                            a, b = node_2_subgraph_node[int(node)], meta_node_2_new_node[cluster]
                            new_edges.append((a, b))
                            new_edges.append((b, a))
                    keys = list(meta_node_2_new_node.keys())
                    adj = co["adj"]
                    for i in range(len(keys) - 1):  # utils.py:224-232
                        for j in range(i + 1, len(keys)):
                            if adj[keys[i], keys[j]] or adj[keys[j], keys[i]]:
                                a, b = meta_node_2_new_node[keys[i]], meta_node_2_new_node[keys[j]]
                                new_edges.append((a, b))
                                new_edges.append((b, a))
                    actual_ext = np.arange(len(value), num_nodes, dtype=np.int64)  # local ids, utils.py:201-206
                    nodes = value
                elif mode == "extra":  # utils.py:235-239
                    extra = adjl.nodes_2_neighbours(value)
                    actual_ext = extra[~np.isin(extra, value)]
                    nodes = np.sort(np.concatenate([value, actual_ext]))  # utils.py:239,243
                else:
                    nodes = value
                ei = induced_subgraph(edge_index, nodes, n)
                xs = x[nodes]
                ys = y[nodes]
                map_dict = {int(v): i for i, v in enumerate(nodes)}  # utils.py:245-247
                if mode == "cluster":  # utils.py:251-259
                    if new_features:
                        xs = np.concatenate([xs, np.asarray(new_features, dtype=np.float32)], 0)
                        ei = np.concatenate([ei, np.asarray(new_edges, dtype=np.int64).T], 1)
                        ys = np.concatenate([ys, np.zeros((len(new_features),) + ys.shape[1:], dtype=ys.dtype)])
                    for new_node in actual_ext:
                        map_dict[int(new_node)] = int(new_node)
                    mask = np.array([True] * len(value) + [False] * len(actual_ext))  # utils.py:262-263
                elif mode == "extra":
                    # utils.py:260-261 — positional over the RE-SORTED list: does not mark the core nodes
                    mask = np.array([True] * (len(nodes) - len(actual_ext)) + [False] * len(actual_ext))
                else:
                    mask = np.ones(len(value), dtype=bool)  # utils.py:264-265
                out.append(dict(x=xs.astype(np.float32), edge_index=ei, y=ys, mask=mask, orig_idx=nodes,
                                actual_ext=actual_ext, map_dict=map_dict, n_real=len(nodes), core=value,
                                cluster_ids=np.asarray(cluster_ids, dtype=np.int64)))
        else:  # utils.py:352-368
            nodes = comp
            if only is not None and len(out) not in only:
                out.append(None)
                continue
            out.append(dict(x=x[nodes].astype(np.float32), edge_index=induced_subgraph(edge_index, nodes, n),
                            y=y[nodes], mask=np.array([True]), orig_idx=nodes, actual_ext=np.zeros(0, dtype=np.int64),
                            map_dict={int(nodes[0]): 0}, n_real=1, core=nodes,
                            cluster_ids=np.zeros(0, dtype=np.int64)))
    return out


def partition_vector(subgraphs, n):
    """part[v] = index in subgraph_list of the subgraph whose cluster owns v (the packed builder's input)."""
    part = np.full(n, -1, dtype=np.int64)
    for i, s in enumerate(subgraphs):
        part[s["core"]] = i
    assert (part >= 0).all()
    return part


def expected_pack(subgraphs, n, mode):
    """The packed block-diagonal CSR the CUDA builder must reproduce bit-exactly from `subgraphs`:
    rows in subgraph order, CSR row = target with the PyG-normalisation edge multiset (self loops of the
    input dropped, one loop per row added, duplicates kept), columns ascending."""
    rowptr, col, gid, sub_ptr, core_rows, is_core, mask = [0], [], [], [0], [], [], []
    base = 0
    for i, s in enumerate(subgraphs):
        ns = s["x"].shape[0]
        src, dst = s["edge_index"]
        keep = src != dst
        src = np.concatenate([src[keep], np.arange(ns)])
        dst = np.concatenate([dst[keep], np.arange(ns)])
        order = np.lexsort((src, dst))
        src, dst = src[order], dst[order]
        cnt = np.bincount(dst, minlength=ns)
        col.append(src + base)
        rowptr.extend((rowptr[-1] + np.cumsum(cnt)).tolist())
        g = np.concatenate([s["orig_idx"], n + s["cluster_ids"]]) if mode == "cluster" else s["orig_idx"]
        gid.append(g)
        core = np.zeros(ns, dtype=bool)
        core[np.searchsorted(s["orig_idx"], s["core"])] = True
        is_core.append(core)
        mask.append(s["mask"])
        core_rows.append(base + np.nonzero(core)[0])
        base += ns
        sub_ptr.append(base)
    rowptr = np.asarray(rowptr, dtype=np.int64)
    deg = np.diff(rowptr).astype(np.float32)
    return dict(rowptr=rowptr, col=np.concatenate(col), dinv=(1.0 / np.sqrt(deg)).astype(np.float32),
                gid=np.concatenate(gid), sub_ptr=np.asarray(sub_ptr), core_rows=np.concatenate(core_rows),
                is_core=np.concatenate(is_core), mask=np.concatenate(mask))


# ============================================================================================
# a13  per-subgraph split masks (utils.py:683-703; regression twin :788-808)
# ============================================================================================


def subgraph_split_masks(subgraphs, train_mask, val_mask, test_mask, mode):
    """Map the global masks through map_dict; extra nodes (global ids in actual_ext, extra mode) and cluster
    nodes (local ids in actual_ext, cluster mode) are forced to False."""
    # Literal restatement, including the key collision of utils.py:258-259: cluster nodes are entered into
    # map_dict under their LOCAL id, overwriting a real node whose GLOBAL id happens to equal it; that real
    # node then never receives its masks.
    out = []
    for s in subgraphs:
        ns = s["x"].shape[0]
        tr, va, te = (np.zeros(ns, dtype=bool) for _ in range(3))
        ext = set(int(v) for v in s["actual_ext"])
        for node, new_node in s["map_dict"].items():
            if train_mask[node]:
                tr[new_node] = True
            if val_mask[node]:
                va[new_node] = True
            if test_mask[node]:
                te[new_node] = True
            if mode == "extra" and node in ext:  # utils.py:695-698 (actual_ext holds global ids)
                tr[new_node] = va[new_node] = te[new_node] = False
            if mode == "cluster" and new_node in ext:  # utils.py:699-702 (actual_ext holds local ids)
                tr[new_node] = va[new_node] = te[new_node] = False
        out.append((tr, va, te))
    return out


# ============================================================================================
# a12  Gc assembly (utils.py:705-778 load_data_classification; :811-852 load_graph_data)
# ============================================================================================


def assemble_gc_classification(comps, Cs, Ws_coarse, Ws_comp, features, labels, train_mask, val_mask, n_classes):
    """utils.py:705-772.  Cs / Ws_coarse are indexed by the candidate index (components > 10 nodes come first
    because of the size-descending sort); Ws_comp[i] is the component's own adjacency (H.W).
    Returns coarsen_features, train_labels, train_mask, val_labels, val_mask, coarsen_edge."""
    feats, tl, tm, vl, vm, rows, cols = [], [], [], [], [], [], []
    node_off = 0
    started = False
    for number, keep in enumerate(comps):
        keep = np.asarray(keep)
        Hf, Hl = features[keep], labels[keep]
        Htm, Hvm = train_mask[keep], val_mask[keep]
        if len(keep) > 10 and Htm.sum() + Hvm.sum() > 0:
            C = Cs[number]
            onehot = np.eye(n_classes)[Hl]
            trl = onehot.copy(); trl[~Htm] = 0
            vall = onehot.copy(); vall[~Hvm] = 0
            ctl, cvl = C.dot(trl), C.dot(vall)

            def pure(cl):  # utils.py:726-730: any label mass, and not mixed
                m = np.asarray(cl.sum(axis=1)).ravel().astype(bool)
                mix = (cl > 0).sum(axis=1)
                m[np.asarray(mix).ravel() > 1] = False
                return m

            feats.append(project_features(C, Hf))
            tl.append(np.argmax(ctl.astype(np.float32), axis=1)); tm.append(pure(ctl))
            vl.append(np.argmax(cvl.astype(np.float32), axis=1)); vm.append(pure(cvl))
            coo = sp.lil_matrix(Ws_coarse[number]).tocoo()
            rows.append(coo.row + node_off); cols.append(coo.col + node_off)
            node_off += C.shape[0]
            started = True
        elif Htm.sum() + Hvm.sum() > 0:
            if not started:
                raise Exception("The graph does not need coarsening.")  # utils.py:763
            feats.append(Hf); tl.append(Hl); tm.append(Htm); vl.append(Hl); vm.append(Hvm)
            coo = sp.lil_matrix(Ws_comp[number]).tocoo()
            rows.append(coo.row + node_off); cols.append(coo.col + node_off)
            node_off += len(keep)
    return (np.concatenate(feats).astype(np.float32), np.concatenate(tl).astype(np.int64), np.concatenate(tm),
            np.concatenate(vl).astype(np.int64), np.concatenate(vm),
            np.stack([np.concatenate(rows), np.concatenate(cols)]).astype(np.int64))


# ============================================================================================
# a14 / a5 / a6  batching and the inference drivers (run.py:336, :49-115; inference.py:672-688)
# ============================================================================================


def collate(subgraphs):
    """PyG Batch.from_data_list as used by G_DataLoader(graphs, batch_size=128, shuffle=False) run.py:336."""
    xs, eis, off = [], [], 0
    for s in subgraphs:
        xs.append(torch.as_tensor(s["x"]))
        eis.append(torch.as_tensor(s["edge_index"]) + off)
        off += s["x"].shape[0]
    return torch.cat(xs, 0), torch.cat(eis, 1)


def node_infer_batched(sd, subgraphs, sel_masks, task="node_cls", batch_size=128, no_grad=True):
    """node_infer_Gs_GD run.py:49-115: per 128-subgraph batch holding >= 1 selected node -> forward ->
    out[mask] concatenated in batch order.  Returns the [n_selected, C] outputs."""
    fwd = classify_node if task == "node_cls" else regress_node
    outs = []
    ctx = torch.no_grad() if no_grad else torch.enable_grad()
    with ctx:
        for b in range(0, len(subgraphs), batch_size):
            chunk = subgraphs[b:b + batch_size]
            m = np.concatenate(sel_masks[b:b + batch_size])
            if not m.any():  # run.py:62
                continue
            x, ei = collate(chunk)
            out = fwd(sd, x, ei)
            outs.append(out[torch.as_tensor(m)].detach())
    return torch.cat(outs, 0) if outs else torch.zeros(0, sd["lt1.weight"].shape[0])


def node_infer_per_query(sd, subgraphs, queries, task="node_cls"):
    """inference.py:672-688: one subgraph forward per query (subgraph index i, local node j) -> out[j]."""
    fwd = classify_node if task == "node_cls" else regress_node
    with torch.no_grad():
        return torch.stack([fwd(sd, torch.as_tensor(subgraphs[i]["x"]), torch.as_tensor(subgraphs[i]["edge_index"]))[j]
                            for i, j in queries])


def subgraphs_from_partition(edge_index, x, part, sub_ids):
    """Vectorised form of build_subgraphs(mode='none') for a SAMPLE of clusters of one big graph (bench.py's CPU
    baseline): subgraph i = nodes of cluster sub_ids[i] ascending + induced edges in original edge order,
    relabelled (utils.py:243-248).  Checked against build_subgraphs in tests/test_oracle_golden.py."""
    edge_index = np.asarray(edge_index)
    part = np.asarray(part)
    sub_ids = np.asarray(sub_ids)
    k = int(part.max()) + 1
    rank_of = np.full(k, -1, dtype=np.int64)
    rank_of[sub_ids] = np.arange(len(sub_ids))
    node_rank = rank_of[part]
    nodes = np.nonzero(node_rank >= 0)[0]
    order = np.lexsort((nodes, node_rank[nodes]))
    nodes = nodes[order]
    bounds = np.searchsorted(node_rank[nodes], np.arange(len(sub_ids) + 1))
    local = np.zeros(part.shape[0], dtype=np.int64)
    local[nodes] = np.arange(len(nodes)) - bounds[node_rank[nodes]]
    src, dst = edge_index
    em = (node_rank[src] >= 0) & (node_rank[src] == node_rank[dst])
    es, ed, er = src[em], dst[em], node_rank[src[em]]
    eorder = np.argsort(er, kind="stable")
    es, ed, er = es[eorder], ed[eorder], er[eorder]
    ebounds = np.searchsorted(er, np.arange(len(sub_ids) + 1))
    out = []
    x = np.asarray(x)
    for i in range(len(sub_ids)):
        nd = nodes[bounds[i]:bounds[i + 1]]
        a, b = ebounds[i], ebounds[i + 1]
        out.append(dict(x=x[nd].astype(np.float32), edge_index=np.stack([local[es[a:b]], local[ed[a:b]]]),
                        mask=np.ones(len(nd), dtype=bool), orig_idx=nd, core=nd, n_real=len(nd),
                        actual_ext=np.zeros(0, dtype=np.int64), cluster_ids=np.zeros(0, dtype=np.int64)))
    return out


# ============================================================================================
# Group-aligned layout of a pack (no reference counterpart: a pure re-layout; the checker for
# fitgnn_pack_align_* and for the aggregation fused into the transform epilogue)
# ============================================================================================


def aligned_layout(sub_ptr, group=32, policy="order", rowptr=None):
    """Placement of the subgraphs in the group-aligned layout (restates fitgnn_pack_align_plan).
    policy 'order': greedy in subgraph order; a group is closed with padding when the next subgraph does not fit.
    policy 'degree': order = (largest non-self row degree desc, size desc, index asc); when the next subgraph does not
    fit, subgraphs are taken from the END of that order while they fit, else the group is closed with padding.
    Returns (new_start[n_sub+1] indexed by subgraph, n_rows_aligned) or None when a subgraph has more than `group` rows."""
    sub_ptr = np.asarray(sub_ptr, dtype=np.int64)
    sizes = np.diff(sub_ptr)
    n_sub = sizes.size
    if n_sub and sizes.max() > group:
        return None
    if policy == "degree":
        deg = np.diff(np.asarray(rowptr, dtype=np.int64)) - 1
        maxdeg = np.array([min(max(int(deg[sub_ptr[s]:sub_ptr[s + 1]].max()) if sizes[s] else 0, 0), 63) for s in range(n_sub)],
                          dtype=np.int64)
        order = np.lexsort((np.arange(n_sub), -sizes, -maxdeg))
    else:
        order = np.arange(n_sub)
    new_start = np.zeros(n_sub + 1, dtype=np.int64)
    pos, head, tail = 0, 0, n_sub - 1
    while head <= tail:
        size = int(sizes[order[head]])
        rem = group - pos % group
        if size <= rem:
            new_start[order[head]] = pos
            pos += size
            head += 1
        elif policy == "degree" and head < tail and sizes[order[tail]] <= rem:
            new_start[order[tail]] = pos
            pos += int(sizes[order[tail]])
            tail -= 1
        else:
            pos += rem
    new_start[-1] = pos
    return new_start, int(pos)


def aligned_pack(pack, group=32, policy="order"):
    """Expected arrays of Pack.aligned(): `pack` is a dict of numpy arrays (rowptr, col, dinv, gid, sub_ptr, core_rows,
    is_core, mask).  Padding rows: empty CSR row, dinv 0, gid 0, flags 0, orig_row -1."""
    lay = aligned_layout(pack["sub_ptr"], group, policy, pack["rowptr"])
    if lay is None:
        return None
    new_start, n_al = lay
    sub_ptr = np.asarray(pack["sub_ptr"], dtype=np.int64)
    n_rows = int(sub_ptr[-1]) if sub_ptr.size else 0
    sizes = np.diff(sub_ptr)
    shift_of_row = np.repeat(new_start[:-1] - sub_ptr[:-1], sizes)
    new_of_old = np.arange(n_rows) + shift_of_row
    orig_row = np.full(n_al, -1, dtype=np.int64)
    orig_row[new_of_old] = np.arange(n_rows)
    rowptr = np.asarray(pack["rowptr"], dtype=np.int64)
    deg = np.zeros(n_al, dtype=np.int64)
    deg[new_of_old] = np.diff(rowptr)
    rowptr_a = np.concatenate([[0], np.cumsum(deg)])
    col = np.asarray(pack["col"], dtype=np.int64)
    col_a = np.zeros(col.size, dtype=np.int64)
    for r in range(n_rows):  # entries keep their order inside a row; rows land where the aligned row pointers say
        e0, e1 = rowptr[r], rowptr[r + 1]
        d0 = rowptr_a[new_of_old[r]]
        col_a[d0:d0 + (e1 - e0)] = col[e0:e1] + shift_of_row[r]
    out = dict(rowptr=rowptr_a, col=col_a, new_of_old=new_of_old, orig_row=orig_row, sub_ptr=new_start,
               core_rows=new_of_old[np.asarray(pack["core_rows"], dtype=np.int64)], n_rows=n_al)
    for name, dt in (("dinv", np.float32), ("gid", np.int64), ("is_core", np.uint8), ("mask", np.uint8)):
        a = np.zeros(n_al, dtype=dt)
        a[new_of_old] = pack[name]
        out[name] = a
    # aggregation descriptors: count of non-self entries + their row-in-group, in CSR order
    desc = np.zeros(n_al, dtype=np.uint64)
    truncated = False
    for r in range(n_rows):
        nr = int(new_of_old[r])
        lanes, self_seen = [], False
        for c in col[rowptr[r]:rowptr[r + 1]]:
            if c == r and not self_seen:
                self_seen = True
                continue
            lanes.append((int(c) + int(shift_of_row[r])) % group)
        if len(lanes) > 12:
            truncated, lanes = True, lanes[:12]
        d = len(lanes)
        for k_, ln in enumerate(lanes):
            d |= ln << (4 + 5 * k_)
        desc[nr] = d
    out["agg_desc"], out["agg_ok"] = desc, not truncated
    return out


def aggregate_dense(rowptr, col, dinv, H):
    """Â·H for a CSR with materialised self loops: out[r] = dinv[r] * sum_e dinv[col_e] * H[col_e]  (fp64)."""
    H = np.asarray(H, dtype=np.float64)
    out = np.zeros_like(H)
    dinv = np.asarray(dinv, dtype=np.float64)
    for r in range(len(rowptr) - 1):
        cs = col[rowptr[r]:rowptr[r + 1]]
        if len(cs):
            out[r] = dinv[r] * (dinv[cs, None] * H[cs]).sum(0)
    return out
