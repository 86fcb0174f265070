"""Multi-GPU: subgraphs are independent (the pack is block-diagonal), so they are sharded across ranks in
size-balanced bins; weights and the de-duplicated feature table are replicated; the only collective on the
inference path is the all-gather of the core-node outputs.  The reference has no distributed
code at all — one process per GPU, torch.distributed (NCCL on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

import torch

from .infer import select_subgraphs
from .pack import Pack


def subgraph_costs(sub_rows: torch.Tensor, sub_nnz: torch.Tensor, hidden: int, in_features: int):
    """Per-subgraph cost model c_i = nnz_i*H (gathers) + rows_i*(H + F) (dense rows), arbitrary units."""
    return sub_nnz.double() * hidden + sub_rows.double() * (hidden + in_features)


def balanced_bins(costs: torch.Tensor, world: int) -> torch.Tensor:
    """Size-balanced assignment of items to `world` bins: sort by cost descending and deal in snake
    (boustrophedon) order — the vectorised form of LPT for many small items, with the handful of huge
    subgraphs landing in different bins first.  Returns bin[i] in [0, world)."""
    n = costs.numel()
    order = torch.argsort(costs, descending=True, stable=True)
    pos = torch.arange(n, device=costs.device)
    rnd, off = pos // world, pos % world
    snake = torch.where(rnd % 2 == 0, off, world - 1 - off)
    bins = torch.empty(n, dtype=torch.long, device=costs.device)
    bins[order] = snake
    return bins


def pack_subgraph_sizes(pack: Pack):
    sp = pack.sub_ptr.long()
    rows = sp[1:] - sp[:-1]
    rp = pack.rowptr.long()
    nnz = rp[sp[1:]] - rp[sp[:-1]]
    return rows, nnz


class ShardedPack:
    """Rank-local slice of a pack plus what the all-gather needs: for every rank the global node ids of its
    core rows (every rank derives all of them from the replicated full pack, so no metadata is exchanged)."""

    def __init__(self, pack: Pack, world: int, rank: int, hidden: int, in_features: int):
        rows, nnz = pack_subgraph_sizes(pack)
        self.bins = balanced_bins(subgraph_costs(rows, nnz, hidden, in_features), world)
        self.world, self.rank = world, rank
        self.sub_ids = [torch.nonzero(self.bins == r).view(-1) for r in range(world)]
        self.local = select_subgraphs(pack, self.sub_ids[rank]) if world > 1 else pack
        # core node ids per rank, in that rank's local pack order
        core_sub = torch.repeat_interleave(torch.arange(pack.n_sub, device=pack.device), rows)[pack.core_rows.long()]
        core_gid = pack.core_gid.long()
        rank_of_core = self.bins[core_sub]
        self.core_ids = []
        for r in range(world):
            # select_subgraphs keeps subgraphs in ascending id order -> core rows keep their relative order
            self.core_ids.append(core_gid[rank_of_core == r])
        self.counts = [int(c.numel()) for c in self.core_ids]
        self.max_count = max(self.counts) if self.counts else 0
        self.n_nodes = pack.n_nodes
        if world > 1:
            assert torch.equal(self.local.core_gid.long(), self.core_ids[rank])
        self.loads = [float(subgraph_costs(rows[s], nnz[s], hidden, in_features).sum()) for s in self.sub_ids]

    # ---- all-gather of the core-node outputs -------------------------------------------------------------
    def gather_buffer(self, C: int, device, dtype=torch.float32) -> torch.Tensor:
        """[world, max_count, C] buffer; rank r's logits live in slot r (rows beyond counts[r] are padding).  Pass
        `slot(buf)` as `out=` to the forward so the head kernel writes straight into the buffer (no staging copy)."""
        return torch.zeros(self.world, self.max_count, C, dtype=dtype, device=device)

    def slot(self, buf: torch.Tensor) -> torch.Tensor:
        return buf[self.rank, : self.counts[self.rank]]

    def all_gather_(self, buf: torch.Tensor, group=None) -> torch.Tensor:
        """In-place all-gather: after the call every rank holds every rank's slot."""
        if self.world > 1:
            import torch.distributed as dist
            dist.all_gather_into_tensor(buf.view(self.world * self.max_count, -1), buf[self.rank], group=group)
        return buf

    def node_index(self, device) -> torch.Tensor:
        """row_of_node[v] = flat row of node v in the gather buffer (buf.view(-1, C)[row_of_node] is node order)."""
        idx = torch.empty(self.n_nodes, dtype=torch.long, device=device)
        for r in range(self.world):
            idx[self.core_ids[r].to(device)] = r * self.max_count + torch.arange(self.counts[r], device=device)
        return idx

    def gather_outputs(self, local_out: torch.Tensor, group=None) -> torch.Tensor:
        """Convenience form: all-gather the per-rank core outputs and return [N, C] in global node order."""
        buf = self.gather_buffer(local_out.shape[1], local_out.device, local_out.dtype)
        self.slot(buf).copy_(local_out)
        self.all_gather_(buf, group)
        return buf.view(self.world * self.max_count, -1)[self.node_index(local_out.device)]
