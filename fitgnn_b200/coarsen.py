"""Host glue between the reference's coarsening outputs and the device builders.

The coarsening ALGORITHM (graph_coarsening.coarsen: eigsh + greedy contraction) produces, once, the coarsening
matrices C; `coarsen_algo.py` implements it (candidate costs, spectral basis and level projections
on the device, the sequential contraction on the host).  Everything derived from C on the hot path runs on the device:
  * partition vector + C weights  <- C.indices / C.data  (subgraph_mapping utils.py:113-121)
  * Xc = C·X                      <- utils.py:161, :738, :827
  * Ac pattern                    <- coarsening_utils.py:138 / utils.py:745-746
  * Gc assembly                   <- utils.py:705-778 (node tasks: assemble_gc_classification),
                                     :811-852 (graph tasks: load_graph_data / load_graph_data_batch)
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import torch

from . import ops


@dataclass
class Partition:
    """Global partition in the reference's subgraph_list order."""
    part: np.ndarray      # int32 [N]: subgraph index of every node
    cweight: np.ndarray   # float64 [N]: C[part[v], v] (1.0 for un-coarsened single-node components)
    k: int
    comp_of_sub: np.ndarray   # int32 [k]: component (candidate index) each subgraph came from
    sub_offset: np.ndarray    # int64 [n_comp+1]: first subgraph of every component


def partition_from_components(comps, C_list, n_nodes) -> Partition:
    """comps: list of sorted node-id arrays in candidate order (size-descending, utils.py:146);
    C_list[i]: scipy CSC coarsening matrix of component i (exactly one non-zero per column), or None for a
    single-node component (utils.py:352-368).  Subgraphs are numbered component by component, supernodes
    ascending — the order coarsening_classification appends them to subgraph_list (utils.py:186, :267)."""
    part = np.full(n_nodes, -1, dtype=np.int64)
    cw = np.ones(n_nodes, dtype=np.float64)
    comp_of_sub, offs, base = [], [0], 0
    for i, (comp, C) in enumerate(zip(comps, C_list)):
        comp = np.asarray(comp, dtype=np.int64)
        if C is None:
            assert len(comp) == 1
            part[comp] = base
            kc = 1
        else:
            C = C.tocsc()
            if not np.all(np.diff(C.indptr) == 1):
                raise ValueError("coarsening matrix must have exactly one non-zero per column")
            part[comp] = base + C.indices
            cw[comp] = C.data
            kc = C.shape[0]
        comp_of_sub.extend([i] * kc)
        base += kc
        offs.append(base)
    if (part < 0).any():
        raise ValueError("components do not cover every node")
    return Partition(part.astype(np.int32), cw, base, np.asarray(comp_of_sub, dtype=np.int32),
                     np.asarray(offs, dtype=np.int64))


def project(edge_index: torch.Tensor, X: torch.Tensor, partition: Partition):
    """Device projection of one graph: returns dict(Xc [k,F] fp32, ac_row, ac_col (int64, row-major sorted),
    ac_cnt int32, ac_rowptr int32, members, member_ptr)."""
    dev = edge_index.device
    part = torch.as_tensor(partition.part, device=dev)
    cw = torch.as_tensor(partition.cweight, device=dev)
    members, member_ptr = ops.group_by_part(part, partition.k)
    Xc = ops.project_features(members, member_ptr, cw, X.contiguous())
    row, col, cnt, rowptr = ops.project_adj(edge_index, part, partition.k)
    return dict(Xc=Xc, ac_row=row, ac_col=col, ac_cnt=cnt, ac_rowptr=rowptr, members=members, member_ptr=member_ptr,
                part=part)


def assemble_gc_classification(proj, partition: Partition, comps, edge_index, X, y, train_mask, val_mask, n_classes):
    """load_data_classification utils.py:705-778 on top of the device projection: components with > 10 nodes
    holding a train/val node contribute their coarsened block (Xc rows, Ac pattern, projected labels and the
    "pure class" masks, utils.py:726-742); smaller ones are appended un-coarsened (utils.py:754-769); the
    rest are skipped.  Returns (coarsen_features, train_labels, train_mask, val_labels, val_mask,
    coarsen_edge) as device tensors, in the reference's row order."""
    dev = X.device
    part = proj["part"].long()
    k = partition.k
    y = y.to(dev).long().view(-1)
    trm, vam = train_mask.to(dev).bool(), val_mask.to(dev).bool()
    cw = torch.as_tensor(partition.cweight, device=dev)

    def proj_labels(mask):
        # C.dot(onehot(y) * mask): [k, n_classes] float64 accumulate (utils.py:714-727)
        idx = torch.nonzero(mask).view(-1)
        M = torch.zeros(k, n_classes, dtype=torch.float64, device=dev)
        M.index_put_((part[idx], y[idx]), cw[idx], accumulate=True)
        any_mass = M.sum(1).bool()                     # torch.BoolTensor(np.sum(C.dot(labels), axis=1))
        mixed = (M > 0).sum(1) > 1
        return torch.argmax(M.float(), dim=1), any_mass & ~mixed

    ctl, ctm = proj_labels(trm)
    cvl, cvm = proj_labels(vam)
    feats, tl, tm, vl, vm, rows, cols = [], [], [], [], [], [], []
    node_off = 0
    started = False
    ac_row, ac_col, ac_rowptr = proj["ac_row"], proj["ac_col"], proj["ac_rowptr"].long()
    ei = edge_index
    for i, comp in enumerate(comps):
        comp_t = torch.as_tensor(np.asarray(comp), device=dev, dtype=torch.long)
        has = bool((trm[comp_t].sum() + vam[comp_t].sum()) > 0)
        if len(comp) > 10 and has:
            s0, s1 = int(partition.sub_offset[i]), int(partition.sub_offset[i + 1])
            feats.append(proj["Xc"][s0:s1])
            tl.append(ctl[s0:s1]); tm.append(ctm[s0:s1]); vl.append(cvl[s0:s1]); vm.append(cvm[s0:s1])
            e0, e1 = int(ac_rowptr[s0]), int(ac_rowptr[s1])
            rows.append(ac_row[e0:e1] - s0 + node_off); cols.append(ac_col[e0:e1] - s0 + node_off)
            node_off += s1 - s0
            started = True
        elif has:
            if not started:
                raise Exception("The graph does not need coarsening.")  # utils.py:763
            feats.append(X[comp_t]); tl.append(y[comp_t]); tm.append(trm[comp_t]); vl.append(y[comp_t]); vm.append(vam[comp_t])
            # H.W.tocoo(): the component's own adjacency, row-major sorted, duplicates merged
            relabel = torch.full((X.shape[0],), -1, dtype=torch.long, device=dev)
            relabel[comp_t] = torch.arange(len(comp), device=dev)
            a, b = relabel[ei[0]], relabel[ei[1]]
            keep = (a >= 0) & (b >= 0)
            keys = torch.unique(a[keep] * len(comp) + b[keep])
            rows.append(keys // len(comp) + node_off); cols.append(keys % len(comp) + node_off)
            node_off += len(comp)
    return (torch.cat(feats), torch.cat(tl), torch.cat(tm), torch.cat(vl), torch.cat(vm),
            torch.stack([torch.cat(rows), torch.cat(cols)]))


def load_graph_data(edge_index: torch.Tensor, X: torch.Tensor, y, partition: Partition, comps=None):
    """`load_graph_data(data, C_LIST, GC_LIST, candidate)` /root/reference/utils.py:811-852 (called per graph at
    main.py:370-381): the coarsened graph of ONE graph for the graph-level *_gc models.  Components in candidate order
    (size-descending): a component with > 1 node contributes `C.dot(H_features)` and `GC.W.tocoo()` row/col shifted by
    the running supernode count; a single-node component contributes its own feature row (its H.W is empty).  With the
    partition numbered component by component (partition_from_components: a single-node component is its own cluster with
    C weight 1) that is exactly the device projection of the whole graph: x = Xc, edge_index = the row-major pattern of
    Ac.  Raises like the reference when the FIRST candidate is a single node (utils.py:841).
    Returns dict(x [k, F] fp32, edge_index [2, nnz(Ac)] int64, y)."""
    if comps is not None and len(comps) and len(comps[0]) <= 1:
        raise Exception("The graph does not need coarsening.")
    proj = project(edge_index, X, partition)
    return dict(x=proj["Xc"], edge_index=torch.stack([proj["ac_row"], proj["ac_col"]]), y=y)


def load_graph_data_batch(edge_indices, xs, partitions, device="cuda"):
    """load_graph_data for a whole dataset in ONE projection (main.py:370-381 loops it per graph; the results are then
    collated by colater / PyG Batch, utils.py:893-908): the graphs are concatenated into one block-diagonal graph (node
    and cluster ids offset), projected once on the device, and returned in the collated form the *_gc models take —
    dict(x [sum k_g, F], edge_index [2, sum nnz_g] (already offset), batch [sum k_g] graph of every supernode,
    ptr [n_graphs + 1] supernode offsets).  Graph g's own Gc is rows ptr[g]:ptr[g+1] and the edges between them."""
    dev = torch.device(device)
    n_off, k_off, eis, parts, cws, ptr = 0, 0, [], [], [], [0]
    for ei, x, p in zip(edge_indices, xs, partitions):
        eis.append(torch.as_tensor(ei).long() + n_off)
        parts.append(torch.as_tensor(p.part).long() + k_off)
        cws.append(torch.as_tensor(p.cweight, dtype=torch.float64))
        n_off += int(torch.as_tensor(x).shape[0])
        k_off += int(p.k)
        ptr.append(k_off)
    X = torch.cat([torch.as_tensor(x).float().view(torch.as_tensor(x).shape[0], -1) for x in xs]).to(dev).contiguous()
    ei = torch.cat(eis, 1).to(dev).contiguous()
    part = torch.cat(parts).to(device=dev, dtype=torch.int32).contiguous()
    cw = torch.cat(cws).to(dev)
    members, member_ptr = ops.group_by_part(part, k_off)
    Xc = ops.project_features(members, member_ptr, cw, X)
    row, col, cnt, _ = ops.project_adj(ei, part, k_off, want_rowptr=False)
    ptr_t = torch.tensor(ptr, device=dev)
    batch = torch.repeat_interleave(torch.arange(len(ptr) - 1, device=dev), ptr_t[1:] - ptr_t[:-1])
    return dict(x=Xc, edge_index=torch.stack([row, col]), batch=batch, ptr=ptr_t, cnt=cnt)
