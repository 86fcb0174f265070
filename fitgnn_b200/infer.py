"""Inference drivers over a pack — the callers of the hot path.

  node_infer_Gs     <- node_infer_Gs_GD /root/reference/run.py:49-115 (and the MB variant :117-175): the
                       reference loops over 128-subgraph batches, skips batches without a selected node, and
                       concatenates out[mask] in batch order — i.e. the selected rows in pack order.
  per_query         <- the per-sample loop /root/reference/inference.py:672-688 (node_reg :805-819): one
                       subgraph forward per queried node; here all queried subgraphs run as one small pack.
"""
from __future__ import annotations

import torch

from .engine import PackedForward
from .pack import Pack


def node_infer_Gs(state_dict, pack: Pack, X, select_mask=None, task="node_cls", precision="bf16x3",
                  reference_quirks=True):
    """Whole-pack inference.  Returns (out [n_selected, C], node_ids [n_selected]) with rows in the order the
    reference concatenates them.  select_mask: global bool mask (test_mask / val_mask) or None for every node."""
    head = "log_softmax" if task == "node_cls" else "identity"
    fwd = PackedForward(pack, state_dict, head=head, rows="core", precision=precision)
    out = fwd(X)
    ids = pack.core_gid
    if select_mask is not None:
        sel = pack.split_masks(select_mask, reference_quirks)[pack.core_rows.long()]
        out, ids = out[sel], ids[sel]
    return out, ids


def select_subgraphs(pack: Pack, sub_ids: torch.Tensor, return_rows: bool = False):
    """The pack restricted to the given subgraphs (kept in the given order); block-diagonal, so columns only
    need re-basing.  Index plumbing on the device, no host loop.  return_rows: also return the source-pack row of every
    row of the result (int64)."""
    dev = pack.device
    sub_ids = sub_ids.to(dev).long()
    sp = pack.sub_ptr.long()
    lens = sp[sub_ids + 1] - sp[sub_ids]
    new_sub_ptr = torch.zeros(sub_ids.numel() + 1, dtype=torch.long, device=dev)
    new_sub_ptr[1:] = torch.cumsum(lens, 0)
    n_rows = int(new_sub_ptr[-1])
    which = torch.repeat_interleave(torch.arange(sub_ids.numel(), device=dev), lens)
    rows = sp[sub_ids][which] + (torch.arange(n_rows, device=dev) - new_sub_ptr[:-1][which])
    shift = (new_sub_ptr[:-1] - sp[sub_ids])[which]  # new row = old row + shift
    rp = pack.rowptr.long()
    deg = rp[rows + 1] - rp[rows]
    new_rowptr = torch.zeros(n_rows + 1, dtype=torch.long, device=dev)
    new_rowptr[1:] = torch.cumsum(deg, 0)
    nnz = int(new_rowptr[-1])
    erow = torch.repeat_interleave(torch.arange(n_rows, device=dev), deg)
    eidx = rp[rows][erow] + (torch.arange(nnz, device=dev) - new_rowptr[:-1][erow])
    col = pack.col.long()[eidx] + shift[erow]
    is_core = pack.is_core[rows]
    core_rows = torch.nonzero(is_core).view(-1)
    sub = Pack(n_rows=n_rows, nnz=nnz, n_sub=sub_ids.numel(), n_core=core_rows.numel(), n_src=pack.n_src,
               n_nodes=pack.n_nodes, mode=pack.mode, rowptr=new_rowptr.to(torch.int32), col=col.to(torch.int32),
               dinv=pack.dinv[rows].contiguous(), gid=pack.gid[rows].contiguous(),
               sub_ptr=new_sub_ptr.to(torch.int32), core_rows=core_rows.to(torch.int32),
               is_core=is_core.contiguous(), mask=pack.mask[rows].contiguous(), part=pack.part)
    return (sub, rows) if return_rows else sub


def per_query(state_dict, pack: Pack, X, query_nodes: torch.Tensor, task="node_cls", precision="bf16x3"):
    """Outputs for the queried nodes, each computed on its own subgraph only (inference.py:672-688), all
    queried subgraphs batched into one small pack.  Returns [n_queries, C] in query order."""
    dev = pack.device
    q = query_nodes.to(dev).long()
    subs, inv = torch.unique(pack.part.long()[q], return_inverse=True)
    small = select_subgraphs(pack, subs)
    head = "log_softmax" if task == "node_cls" else "identity"
    out = PackedForward(small, state_dict, head=head, rows="core", precision=precision)(X)
    ids = small.core_gid.long()
    lookup = torch.full((pack.n_nodes,), -1, dtype=torch.long, device=dev)
    lookup[ids] = torch.arange(ids.numel(), device=dev)
    return out[lookup[q]]


def graph_level_Gs(state_dict, pack: Pack, X, graph_of_sub: torch.Tensor, task="graph_reg", precision="bf16x3"):
    """Graph-level models on subgraphs (Classify_graph_gs / Regress_graph_gs, /root/reference/network.py:118-135,
    :189-204) for a whole dataset at once: `pack` holds the subgraphs of ALL graphs (subgraphs numbered graph by graph,
    as main.py:370-381 builds them), `graph_of_sub[s]` is the graph each subgraph belongs to (non-decreasing).
    conv stack on every subgraph -> rows with M.mask -> global max / mean pool per graph -> lt1 -> softmax / identity.
    Returns [n_graphs, C]."""
    from . import ops
    dev = pack.device
    fwd = PackedForward(pack, state_dict, rows="mask", precision=precision, with_head=False)
    h = fwd(X)
    if isinstance(h, tuple):
        raise RuntimeError("graph_level_Gs needs the fp32 hidden state")
    rows = pack.mask_rows().long()
    sp = pack.sub_ptr.long()
    sub_of_row = torch.repeat_interleave(torch.arange(pack.n_sub, device=dev), sp[1:] - sp[:-1])
    g = graph_of_sub.to(dev).long()[sub_of_row[rows]]
    assert bool((g[1:] >= g[:-1]).all()), "subgraphs must be numbered graph by graph"
    n_graphs = int(graph_of_sub.max().item()) + 1
    seg_ptr = torch.zeros(n_graphs + 1, dtype=torch.int32, device=dev)
    seg_ptr[1:] = torch.cumsum(torch.bincount(g, minlength=n_graphs), 0).to(torch.int32)
    pooled = ops.segment_pool(h, None, seg_ptr, ops.POOL_MAX if task == "graph_cls" else ops.POOL_MEAN)
    w = state_dict["lt1.weight"].detach().to(dev).float().contiguous()
    b = state_dict["lt1.bias"].detach().to(dev).float().contiguous()
    return ops.gemm_bias_act(pooled, w, b, ops.ACT_NONE, ops.HEAD_SOFTMAX if task == "graph_cls" else ops.HEAD_IDENTITY)


def batch_of_rows(sub_ptr, rows, batch_size=128):
    """Index of the reference's DataLoader batch (`G_DataLoader(graphs, batch_size)`, run.py:336: consecutive subgraphs)
    every pack row in `rows` falls into, and the number of batches.  Pure index arithmetic (any device)."""
    sp = sub_ptr.long()
    n_sub = sp.numel() - 1
    sub = torch.searchsorted(sp, rows.long().to(sp.device), right=True) - 1
    return sub // batch_size, (n_sub + batch_size - 1) // batch_size


def node_metrics(out, labels, task="node_cls", loss_reduction="mean", batch_ids=None, n_batches=None):
    """(loss, acc) exactly as the reference's evaluation drivers report them, from the selected rows' outputs
    (log-probabilities [n, C] for node_cls, predictions [n] / [n, 1] for node_reg) and labels, both in the order
    `node_infer_Gs` returns them.

    batch_ids None -> `node_infer_Gs_GD` (/root/reference/run.py:99-115): node_cls = NLLLoss_numpy (utils.py:927-954) +
    accuracy; node_reg = L1Loss_numpy (utils.py:972-987) divided by the population std of the labels, acc = 0; with
    loss_reduction='sum' the summed loss is divided by the number of rows.
    batch_ids given (batch_of_rows) -> `node_infer_Gs_MB` (run.py:117-175): the loss is taken per DataLoader batch and the
    per-batch values are added up; 'mean' divides by the number of ALL batches `n_batches` (batches without a selected
    row included), 'sum' by the number of rows; node_reg divides by the UNBIASED std of the labels (torch.std).
    Host arithmetic on numpy (the reference does the same for GD)."""
    import numpy as np
    o = out.detach().cpu().numpy() if torch.is_tensor(out) else np.asarray(out)
    y = labels.detach().cpu().numpy() if torch.is_tensor(labels) else np.asarray(labels)
    if loss_reduction not in ("mean", "sum"):
        raise ValueError("Reduction must be 'mean' or 'sum'.")
    red = np.mean if loss_reduction == "mean" else np.sum
    if task == "node_cls":
        y = y.reshape(-1).astype(np.int64)
        if o.ndim != 2 or y.shape[0] != o.shape[0] or not np.all((y >= 0) & (y < o.shape[1])):
            raise ValueError("node_metrics: log-probabilities must be [n, C] and labels valid class ids of length n")
        per_row = -o[np.arange(o.shape[0]), y].astype(np.float64)
        acc = float(np.sum(np.argmax(o, axis=1) == y) / len(y))
    else:
        o, y = o.reshape(-1).astype(np.float64), y.reshape(-1).astype(np.float64)
        if o.shape != y.shape:
            raise ValueError("node_metrics: predictions and labels must have the same number of elements")
        per_row = np.abs(o - y)
        acc = 0
    if batch_ids is None:
        loss = float(red(per_row))
        if task != "node_cls":
            loss /= float(np.std(y))
        return (loss, acc) if loss_reduction == "mean" else (loss / len(per_row), acc)
    b = batch_ids.detach().cpu().numpy() if torch.is_tensor(batch_ids) else np.asarray(batch_ids)
    if b.shape[0] != per_row.shape[0] or n_batches is None:
        raise ValueError("node_metrics: batch_ids needs one entry per row and n_batches")
    total = float(sum(red(per_row[b == i]) for i in np.unique(b)))
    scale = 1.0 if task == "node_cls" else float(np.std(y, ddof=1))
    denom = n_batches if loss_reduction == "mean" else len(per_row)
    return total / (denom * scale), acc


def graph_metrics(pred, y, task="graph_reg", batch_size=128, reference_quirks=True):
    """(loss, acc) as the reference's graph-level evaluation driver `graph_infer_Gs` (/root/reference/run.py:306-328)
    reports them for predictions `pred` [n_graphs, C] (graph_level_Gs / the graph models) and labels `y` [n_graphs], with
    the graphs batched `batch_size` at a time in order (T_DataLoader + colater, run.py:577-580).  The loss is taken per
    batch and averaged over the batches.  graph_cls: CrossEntropyLoss applied to the model's softmax OUTPUT (run.py:583,
    network.py:131-135 — a log-softmax of probabilities).  graph_reg: L1Loss, divided by the unbiased
    std of the labels.

    reference_quirks=True reproduces three things the driver does: labels are cast to int64 first (`.type(torch.long)`,
    run.py:312 — regression targets are truncated), the L1 loss broadcasts predictions [B, 1] against labels [B]
    (mean over all B x B pairs), and the classification accuracy is that of the LAST batch only (run.py:323).  With
    False: untruncated labels, row-wise L1, accuracy over all graphs."""
    import numpy as np
    p = (pred.detach().cpu().numpy() if torch.is_tensor(pred) else np.asarray(pred)).astype(np.float64)
    t = (y.detach().cpu().numpy() if torch.is_tensor(y) else np.asarray(y)).reshape(-1)
    if p.ndim == 1:
        p = p[:, None]
    if p.shape[0] != t.shape[0] or batch_size < 1:
        raise ValueError("graph_metrics: one label per prediction row and batch_size >= 1")
    if reference_quirks or task == "graph_cls":
        t = t.astype(np.int64)  # numpy truncates toward zero like torch's float -> long cast
    n = p.shape[0]
    bounds = [(a, min(a + batch_size, n)) for a in range(0, n, batch_size)]
    total, acc = 0.0, 0
    for a, b in bounds:
        pb, tb = p[a:b], t[a:b]
        if task == "graph_cls":
            z = pb - pb.max(axis=1, keepdims=True)
            logp = z - np.log(np.exp(z).sum(axis=1, keepdims=True))
            total += float(np.mean(-logp[np.arange(b - a), tb]))
        elif reference_quirks:
            total += float(np.mean(np.abs(pb[:, :1] - tb[None, :].astype(np.float64))))
        else:
            total += float(np.mean(np.abs(pb[:, 0] - tb.astype(np.float64))))
    if task == "graph_cls":
        a, b = bounds[-1] if reference_quirks else (0, n)
        acc = float(np.sum(np.argmax(p[a:b], axis=1) == t[a:b]) / (b - a))
    else:
        total /= float(np.std(t.astype(np.float32 if reference_quirks else np.float64), ddof=1))
    return total / len(bounds), acc
