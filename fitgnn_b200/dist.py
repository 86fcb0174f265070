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

    def gather_outputs(self, local_out: torch.Tensor, group=None) -> torch.Tensor:
        """all-gather(v) of the per-rank core outputs into [N, C] in global node order on every rank."""
        import torch.distributed as dist
        C = local_out.shape[1]
        full = torch.empty(self.n_nodes, C, dtype=local_out.dtype, device=local_out.device)
        if self.world == 1:
            full[self.core_ids[0]] = local_out
            return full
        send = torch.zeros(self.max_count, C, dtype=local_out.dtype, device=local_out.device)
        send[: local_out.shape[0]] = local_out
        recv = torch.empty(self.world, self.max_count, C, dtype=local_out.dtype, device=local_out.device)
        dist.all_gather_into_tensor(recv.view(-1, C), send, group=group)
        for r in range(self.world):
            full[self.core_ids[r]] = recv[r, : self.counts[r]]
        return full
