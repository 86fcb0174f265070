"""The packed block-diagonal CSR of all subgraphs Gs, built once on the device from the partition vector.

Replaces the list[Data] the reference builds in coarsening_classification / coarsening_regression
(/root/reference/utils.py:143-374, :376-605) and re-collates every epoch with G_DataLoader (run.py:336).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import torch

from . import ops
from ._lib import PackStruct, PlanStruct, check, lib, ptr, stream_ptr


@dataclass
class Pack:
    """Device arrays of one pack (see include/fitgnn.h `fitgnn_pack`).  Immutable after build."""
    n_rows: int
    nnz: int
    n_sub: int
    n_core: int
    n_src: int
    n_nodes: int          # N of the graph the pack was built from
    mode: str
    rowptr: torch.Tensor  # int32 [n_rows+1]
    col: torch.Tensor     # int32 [nnz]
    dinv: torch.Tensor    # fp32  [n_rows]
    gid: torch.Tensor     # int32 [n_rows]
    sub_ptr: torch.Tensor  # int32 [n_sub+1]
    core_rows: torch.Tensor  # int32 [n_core]
    is_core: torch.Tensor  # uint8 [n_rows]
    mask: torch.Tensor    # uint8 [n_rows]
    part: torch.Tensor | None = None  # int32 [N]
    # group-aligned packs only (Pack.aligned): see include/fitgnn.h fitgnn_pack_align_*
    orig_row: torch.Tensor | None = None    # int32 [n_rows] row of the source pack, -1 for padding rows
    new_of_old: torch.Tensor | None = None  # int32 [source n_rows] aligned row of every source row
    agg_desc: torch.Tensor | None = None    # int64 [n_rows] per-row neighbour lanes for the fused aggregation
    agg_ok: bool = False                    # every row has <= 12 non-self entries: agg_desc is complete

    def struct(self) -> PackStruct:
        return PackStruct(self.n_rows, self.nnz, self.n_sub, self.n_core, self.n_src,
                          self.rowptr.data_ptr(), self.col.data_ptr(), self.dinv.data_ptr(), self.gid.data_ptr(),
                          self.sub_ptr.data_ptr(), self.core_rows.data_ptr(), self.is_core.data_ptr(),
                          self.mask.data_ptr())

    @property
    def device(self):
        return self.rowptr.device

    @property
    def core_gid(self):
        """Global node id of every core row, in pack order (= the order outputs are returned in)."""
        return self.gid[self.core_rows.long()]

    def mask_rows(self):
        """Pack rows with M.mask == True, ascending (the rows graph-level models pool, network.py:129,200)."""
        return torch.nonzero(self.mask).view(-1).to(torch.int32)

    def split_masks(self, global_mask, reference_quirks=True):
        """Per-row mask = global mask mapped through map_dict with extra / cluster nodes forced to False
        (load_data_classification /root/reference/utils.py:683-703).  With reference_quirks the key
        collision of utils.py:258-259 is reproduced: in cluster mode a real node whose global id equals the
        local id of one of its subgraph's cluster nodes never receives its mask."""
        gm = global_mask.to(self.device).bool()
        rows = torch.zeros(self.n_rows, dtype=torch.bool, device=self.device)
        core = self.is_core.bool()
        gid = self.gid.long()
        rows[core] = gm[gid[core]]
        if reference_quirks and self.mode == "cluster":
            sub_of_row = torch.repeat_interleave(torch.arange(self.n_sub, device=self.device),
                                                 (self.sub_ptr[1:] - self.sub_ptr[:-1]).long())
            n_rows_s = (self.sub_ptr[1:] - self.sub_ptr[:-1]).long()
            n_core_s = torch.zeros(self.n_sub, dtype=torch.long, device=self.device).index_add_(
                0, sub_of_row, core.long())
            lo, hi = n_core_s[sub_of_row], n_rows_s[sub_of_row]
            rows &= ~(core & (gid >= lo) & (gid < hi))
        return rows

    # ---- on-disk form (the reference pickles its subgraph lists: main.py:131-172; a pack is ten flat arrays) -------
    _ARRAYS = ("rowptr", "col", "dinv", "gid", "sub_ptr", "core_rows", "is_core", "mask", "part")
    _SCALARS = ("n_rows", "nnz", "n_sub", "n_core", "n_src", "n_nodes", "mode")

    def save(self, path):
        blob = {k: getattr(self, k) for k in self._SCALARS}
        blob.update({k: (getattr(self, k).cpu() if getattr(self, k) is not None else None) for k in self._ARRAYS})
        blob["format"] = "fitgnn_b200.pack.v1"
        torch.save(blob, path)

    @staticmethod
    def load(path, device="cuda"):
        blob = torch.load(path, map_location="cpu", weights_only=True)
        if blob.get("format") != "fitgnn_b200.pack.v1":
            raise ValueError(f"{path} is not a fitgnn_b200 pack")
        kw = {k: blob[k] for k in Pack._SCALARS}
        kw.update({k: (blob[k].to(device) if blob[k] is not None else None) for k in Pack._ARRAYS})
        return Pack(**kw)

    def aligned(self, group=32, policy="degree"):
        """The same pack re-laid-out so that no subgraph straddles a multiple of `group` rows (padding rows are empty
        CSR rows with dinv = 0).  policy 'order' keeps the subgraph order; 'degree' groups subgraphs of similar row degree
        and fills group tails with low-degree subgraphs instead of padding (see include/fitgnn.h).  Returns None when a
        subgraph has more than `group` rows.  The result carries orig_row / new_of_old / agg_desc for
        fitgnn_gcn_transform_aggregate; its sub_ptr[s] is the first row of subgraph s (not monotone under 'degree')."""
        dev = self.device
        i32 = dict(dtype=torch.int32, device=dev)
        new_sub = torch.empty(self.n_sub + 1, **i32)
        pol = {"order": 0, "degree": 1}[policy]
        ws = ops._ws(lib().fitgnn_pack_align_workspace_bytes(self.n_sub, 0), dev)
        n_al, ok = C.c_int64(0), C.c_int(0)
        src = self.struct()
        check(lib().fitgnn_pack_align_plan(C.byref(src), group, pol, ptr(new_sub), C.byref(n_al), C.byref(ok), ptr(ws),
                                           ws.numel(), stream_ptr()))
        if not ok.value:
            return None
        n = n_al.value
        ws = ops._ws(lib().fitgnn_pack_align_workspace_bytes(self.n_sub, n), dev)
        a = Pack(n_rows=n, nnz=self.nnz, n_sub=self.n_sub, n_core=self.n_core, n_src=self.n_src, n_nodes=self.n_nodes,
                 mode=self.mode, rowptr=torch.empty(n + 1, **i32), col=torch.empty(self.nnz, **i32),
                 dinv=torch.empty(n, dtype=torch.float32, device=dev), gid=torch.empty(n, **i32),
                 sub_ptr=torch.empty(self.n_sub + 1, **i32), core_rows=torch.empty(self.n_core, **i32),
                 is_core=torch.empty(n, dtype=torch.uint8, device=dev), mask=torch.empty(n, dtype=torch.uint8, device=dev),
                 part=self.part, orig_row=torch.empty(n, **i32), new_of_old=torch.empty(self.n_rows, **i32),
                 agg_desc=torch.empty(n, dtype=torch.int64, device=dev))
        flags = C.c_int(0)
        dst = a.struct()
        check(lib().fitgnn_pack_align_fill(C.byref(src), ptr(new_sub), group, n, C.byref(dst), ptr(a.orig_row),
                                           ptr(a.new_of_old), ptr(a.agg_desc), C.byref(flags), ptr(ws), ws.numel(),
                                           stream_ptr()))
        if flags.value & 6:
            raise ValueError("Pack.aligned: the pack is not block-diagonal with one self loop per row")
        a.agg_ok = (flags.value & 1) == 0
        return a

    def nbytes(self):
        return sum(t.numel() * t.element_size() for t in
                   (self.rowptr, self.col, self.dinv, self.gid, self.sub_ptr, self.core_rows, self.is_core, self.mask))


def build_pack(edge_index: torch.Tensor, part: torch.Tensor, k: int, mode: str = "none",
               ac_rowptr: torch.Tensor | None = None, ac_col: torch.Tensor | None = None) -> Pack:
    """Build the pack on the device.  edge_index [2,E] int64 (the reference's COO), part[v] = index of v's
    subgraph in the reference's subgraph_list order.  mode 'cluster' needs the coarsened adjacency pattern
    as CSR (ops.project_adj gives it; computed here when not supplied)."""
    assert edge_index.is_cuda and edge_index.dtype == torch.int64
    ei = edge_index.contiguous()
    part = part.to(device=ei.device, dtype=torch.int32).contiguous()
    N, E = part.numel(), ei.shape[1]
    m = ops.MODES[mode]
    ac_nnz = 0
    if m == ops.MODE_CLUSTER:
        if ac_rowptr is None:
            _, c, _, ac_rowptr = ops.project_adj(ei, part, k)
            ac_col = c.to(torch.int32)
        ac_col = ac_col.to(torch.int32).contiguous()
        ac_nnz = ac_col.numel()
    ws = ops._ws(lib().fitgnn_pack_workspace_bytes(E, N, k, m, ac_nnz), ei.device)
    plan = PlanStruct()
    check(lib().fitgnn_pack_plan(ptr(ei), E, N, ptr(part), k, m, ptr(ac_rowptr), ptr(ac_col), ac_nnz, ptr(ws),
                                 ws.numel(), C.byref(plan), stream_ptr()))
    dev = ei.device
    i32 = dict(dtype=torch.int32, device=dev)
    p = Pack(n_rows=plan.n_rows, nnz=plan.nnz, n_sub=plan.n_sub, n_core=plan.n_core, n_src=plan.n_src, n_nodes=N,
             mode=mode,
             rowptr=torch.empty(plan.n_rows + 1, **i32), col=torch.empty(plan.nnz, **i32),
             dinv=torch.empty(plan.n_rows, dtype=torch.float32, device=dev), gid=torch.empty(plan.n_rows, **i32),
             sub_ptr=torch.empty(plan.n_sub + 1, **i32), core_rows=torch.empty(plan.n_core, **i32),
             is_core=torch.empty(plan.n_rows, dtype=torch.uint8, device=dev),
             mask=torch.empty(plan.n_rows, dtype=torch.uint8, device=dev), part=part)
    ws2 = ops._ws(plan.fill_ws_bytes, dev) if plan.fill_ws_bytes > 0 else None
    st = p.struct()
    check(lib().fitgnn_pack_fill(C.byref(plan), C.byref(st), ptr(ws), ws.numel(), ptr(ws2),
                                 ws2.numel() if ws2 is not None else 0, stream_ptr()))
    torch.cuda.current_stream().synchronize()  # ws / ws2 are released on return
    return p


def subgraph_row_counts(part: torch.Tensor, k: int, mode: str, ac_rowptr: torch.Tensor | None = None) -> torch.Tensor:
    """Rows every subgraph will have in the pack, known before it is built: its cluster's nodes, plus (cluster mode) one
    cluster node per adjacent cluster = the row length of Ac (utils.py:195-213).  int64 [k] on part's device."""
    rows = torch.bincount(part.long(), minlength=k)
    if mode == "cluster":
        rp = ac_rowptr.long()
        rows = rows + (rp[1:] - rp[:-1])
    elif mode != "none":
        raise NotImplementedError("subgraph_row_counts: mode 'extra' row counts need the frontier (build the pack)")
    return rows


def build_pack_range(edge_index: torch.Tensor, part: torch.Tensor, k: int, mode: str, sub_begin: int, sub_end: int,
                     ac_rowptr: torch.Tensor | None = None, ac_col: torch.Tensor | None = None) -> Pack:
    """The pack of subgraphs [sub_begin, sub_end) only — a shard of the reference's subgraph_list (the reference streams
    its list through the model 128 subgraphs at a time, run.py:336; at ogbn-products scale the cluster_node pack of ALL
    subgraphs has ~10^8 rows and more than 2^31 CSR entries, so it only exists as such shards).  Modes 'none' and
    'cluster': every edge of a subgraph derives from an out-edge of one of its core nodes (induced core edges,
    node<->cluster-node edges utils.py:214-222) or from Ac (utils.py:224-232), so the builder runs on the out-edges of
    the range's core nodes; the subgraphs outside the range come out as bare core rows and are dropped.
    Row / subgraph numbering is local to the shard; gid / part stay global."""
    from .infer import select_subgraphs
    if mode not in ("none", "cluster"):
        raise NotImplementedError("build_pack_range: mode 'extra' induces edges between non-core nodes; build the whole pack")
    part = part.to(device=edge_index.device, dtype=torch.int32).contiguous()
    if mode == "cluster" and ac_rowptr is None:  # Ac comes from ALL edges, not from the range's
        _, c, _, ac_rowptr = ops.project_adj(edge_index.contiguous(), part, k)
        ac_col = c.to(torch.int32)
    ps = part[edge_index[0]]
    sel = (ps >= sub_begin) & (ps < sub_end)
    ei_r = edge_index[:, sel].contiguous()
    del ps, sel
    full = build_pack(ei_r, part, k, mode, ac_rowptr, ac_col)
    del ei_r
    return select_subgraphs(full, torch.arange(sub_begin, sub_end, device=edge_index.device))


@dataclass
class PackStream:
    """A pack held as consecutive shards of at most `max_rows` rows each (whole subgraphs): `packs[i]` covers subgraphs
    [bounds[i], bounds[i+1]).  Outputs of a streamed forward are the shards' outputs concatenated = subgraph_list order."""
    packs: list
    bounds: list
    mode: str
    n_nodes: int
    n_src: int

    @property
    def n_rows(self):
        return sum(p.n_rows for p in self.packs)

    @property
    def nnz(self):
        return sum(p.nnz for p in self.packs)

    @property
    def n_sub(self):
        return self.bounds[-1]

    @property
    def n_core(self):
        return sum(p.n_core for p in self.packs)

    @property
    def device(self):
        return self.packs[0].device

    @property
    def core_gid(self):
        return torch.cat([p.core_gid for p in self.packs])

    def nbytes(self):
        return sum(p.nbytes() for p in self.packs)


def build_pack_stream(edge_index: torch.Tensor, part: torch.Tensor, k: int, mode: str = "none", max_rows: int = 1 << 22,
                      ac_rowptr: torch.Tensor | None = None, ac_col: torch.Tensor | None = None) -> PackStream:
    """Build the pack as shards of at most ~max_rows rows (runs of consecutive subgraphs; a single larger subgraph gets a
    shard of its own).  With one shard this is build_pack."""
    dev = edge_index.device
    part = part.to(device=dev, dtype=torch.int32).contiguous()
    if mode == "cluster" and ac_rowptr is None:
        _, c, _, ac_rowptr = ops.project_adj(edge_index.contiguous(), part, k)
        ac_col = c.to(torch.int32)
        del c
    if mode == "extra":
        p = build_pack(edge_index, part, k, mode)
        return PackStream([p], [0, k], mode, p.n_nodes, p.n_src)
    csum = torch.cumsum(subgraph_row_counts(part, k, mode, ac_rowptr), 0)
    if int(csum[-1]) <= max_rows:
        p = build_pack(edge_index, part, k, mode, ac_rowptr, ac_col)
        return PackStream([p], [0, k], mode, p.n_nodes, p.n_src)
    bounds = [0]
    while bounds[-1] < k:
        base = int(csum[bounds[-1] - 1]) if bounds[-1] > 0 else 0
        nxt = int(torch.searchsorted(csum, torch.tensor([base + max_rows], device=dev), right=True).item())
        bounds.append(min(k, max(nxt, bounds[-1] + 1)))
    packs = [build_pack_range(edge_index, part, k, mode, a, b, ac_rowptr, ac_col) for a, b in zip(bounds[:-1], bounds[1:])]
    return PackStream(packs, bounds, mode, packs[0].n_nodes, packs[0].n_src)


def _field(g, name, default=None):
    if isinstance(g, dict):
        return g.get(name, default)
    return getattr(g, name, default)


def pack_from_subgraph_list(graphs, device="cuda"):
    """Collate a reference-style `subgraph_list` into a pack — the data format on the caller's side of the hot path:
    the list `coarsening_classification` returns and `main.py:131-172` saves as `..._subgraph_list.pt`
    (/root/reference/utils.py:248-266), or the `new_graphs` list of `load_data_classification` (utils.py:683-703) that
    `G_DataLoader(graphs, 128)` re-collates every epoch (run.py:336; `Batch.from_data_list`).  Every element is a PyG
    `Data` (any object with attributes) or a dict holding `x [n_s, F]`, `edge_index [2, e_s]` (subgraph-local ids) and
    optionally `mask [n_s]` (M.mask; all True when absent) and `orig_idx [n_s]`.

    Returns (pack, X_packed, node_ids): rows in list order, `X_packed [n_rows, F]` fp32 on the device = the rows of the
    collated `batch.x` (feed it as `fwd(X_packed)`; `pack.gid` is the identity), `node_ids` = concatenated `orig_idx`
    (or None).  The CSR goes through the same device builder as the drop-in GCNConv (`fitgnn_csr_*`: self loops
    replaced by one, duplicates kept), so results equal running the reference model on every list element."""
    dev = torch.device(device)
    sizes, xs, eis, masks, ids = [], [], [], [], []
    off = 0
    for g in graphs:
        x = torch.as_tensor(_field(g, "x"))
        if x.dim() == 1:
            x = x.view(-1, 1)
        ei = torch.as_tensor(_field(g, "edge_index")).long().view(2, -1)
        n_s = x.shape[0]
        if ei.numel() and (int(ei.min()) < 0 or int(ei.max()) >= n_s):
            raise ValueError(f"subgraph {len(sizes)}: edge_index refers to a node outside [0, {n_s})")
        m = _field(g, "mask")
        masks.append(torch.ones(n_s, dtype=torch.bool) if m is None else torch.as_tensor(m).bool().view(-1).cpu())
        oi = _field(g, "orig_idx")
        if oi is not None:  # cluster_node mode: orig_idx covers the real nodes only (utils.py:249); cluster rows get -1
            oi = torch.as_tensor(oi).long().view(-1).cpu()[:n_s]
            oi = torch.cat([oi, torch.full((n_s - oi.numel(),), -1, dtype=torch.long)])
        ids.append(oi)
        sizes.append(n_s)
        xs.append(x.float().cpu())
        eis.append(ei.cpu() + off)
        off += n_s
    n_rows, n_sub = off, len(sizes)
    if n_rows >= 2 ** 31 - 1:
        raise ValueError("pack_from_subgraph_list: more than 2^31 rows")
    F = xs[0].shape[1] if xs else 0
    X = (torch.cat(xs, 0) if xs else torch.zeros(0, F)).to(dev).contiguous()
    ei = (torch.cat(eis, 1) if eis else torch.zeros(2, 0, dtype=torch.long)).to(dev).contiguous()
    mask = (torch.cat(masks) if masks else torch.zeros(0, dtype=torch.bool)).to(dev)
    rowptr, col, dinv = ops.csr_from_coo(ei, n_rows)
    i32 = dict(dtype=torch.int32, device=dev)
    sub_ptr = torch.zeros(n_sub + 1, **i32)
    if n_sub:
        sub_ptr[1:] = torch.cumsum(torch.tensor(sizes, dtype=torch.int64), 0).to(dev)
    core_rows = torch.nonzero(mask).view(-1).to(torch.int32)
    node_ids = None if (not ids or any(i is None for i in ids)) else torch.cat(ids).to(dev)
    pack = Pack(n_rows=n_rows, nnz=col.numel(), n_sub=n_sub, n_core=core_rows.numel(), n_src=n_rows, n_nodes=n_rows,
                mode="list", rowptr=rowptr, col=col, dinv=dinv, gid=torch.arange(n_rows, **i32), sub_ptr=sub_ptr,
                core_rows=core_rows, is_core=mask.to(torch.uint8), mask=mask.to(torch.uint8), part=None)
    return pack, X, node_ids
