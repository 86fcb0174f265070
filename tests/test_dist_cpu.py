"""CPU (gloo, world_size 2): the sharding + all-gather plumbing of fitgnn_b200.dist, on a pack assembled from the
oracle's expected arrays (no GPU needed: ShardedPack is index plumbing over torch tensors)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import fitgnn_oracle as fo
from tests import golden_io as gio


def cpu_pack(mode="extra"):
    from fitgnn_b200.pack import Pack
    d = gio.load("node_mid")
    comps = gio.components(d, mode)
    cos = gio.coarsenings_for_oracle(d, mode, comps)
    n = int(d["n"])
    subs = fo.build_subgraphs(d["edge_index"], d["x"], d["y"], comps, cos, mode)
    w = fo.expected_pack(subs, n, mode)
    t32 = lambda a: torch.tensor(np.asarray(a), dtype=torch.int32)
    return Pack(n_rows=len(w["gid"]), nnz=len(w["col"]), n_sub=len(subs), n_core=len(w["core_rows"]), n_src=n, n_nodes=n,
                mode=mode, rowptr=t32(w["rowptr"]), col=t32(w["col"]), dinv=torch.tensor(w["dinv"]), gid=t32(w["gid"]),
                sub_ptr=t32(w["sub_ptr"]), core_rows=t32(w["core_rows"]),
                is_core=torch.tensor(w["is_core"].astype(np.uint8)), mask=torch.tensor(w["mask"].astype(np.uint8)),
                part=t32(fo.partition_vector(subs, n))), d


def test_balanced_bins_and_local_packs():
    from fitgnn_b200.dist import ShardedPack, balanced_bins
    costs = torch.tensor([100., 1, 1, 1, 50, 50, 2, 2, 2, 90])
    bins = balanced_bins(costs, 2)
    loads = [float(costs[bins == r].sum()) for r in range(2)]
    assert abs(loads[0] - loads[1]) <= 0.2 * sum(loads)
    pack, d = cpu_pack()
    for world, chunks in ((1, 1), (2, 1), (4, 1), (2, 3)):
        shards = [ShardedPack(pack, world, r, 512, 20, n_chunks=chunks) for r in range(world)]
        ids = torch.cat([lp.core_gid.long() for s in shards for lp in s.locals])
        assert sorted(ids.tolist()) == list(range(pack.n_nodes))  # every node is core on exactly one rank
        assert max(shards[0].loads) <= 1.25 * (sum(shards[0].loads) / world) + 1e5
        for s in shards:
          for lp in s.locals:
            # the local pack is a valid block-diagonal CSR: columns stay inside their subgraph
            rows_sub = torch.repeat_interleave(torch.arange(lp.n_sub), (lp.sub_ptr[1:] - lp.sub_ptr[:-1]).long())
            erow = torch.repeat_interleave(torch.arange(lp.n_rows), (lp.rowptr[1:] - lp.rowptr[:-1]).long())
            assert torch.equal(rows_sub[erow], rows_sub[lp.col.long()])


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from fitgnn_b200.dist import ShardedPack
    pack, d = cpu_pack()
    sp = ShardedPack(pack, world, rank, 512, 20, n_chunks=2)
    # stand-in for the forward: "logits" of a node = f(node id), in each chunk's local pack order
    outs = []
    for lp in sp.locals:
        ids = lp.core_gid.long()
        outs.append(torch.stack([ids.float(), ids.float() * 2 + 1], 1))
    full = sp.gather_outputs(outs)
    want = torch.stack([torch.arange(pack.n_nodes).float(), torch.arange(pack.n_nodes).float() * 2 + 1], 1)
    ret[rank] = bool(torch.equal(full, want))
    dist.destroy_process_group()


def test_all_gather_world2_gloo():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    assert ret[0] and ret[1]


@pytest.mark.parametrize("mode", ["none", "extra"])
def test_rank_local_feature_tables(mode):
    """ShardedPack(local_table=True): every rank keeps only the feature rows its own subgraphs reference and its packs
    index into that table; gathering X[table_ids] rank by rank reproduces exactly the rows the full pack would read."""
    from fitgnn_b200.dist import ShardedPack
    pack, d = cpu_pack(mode)
    X = torch.arange(pack.n_src, dtype=torch.float32)[:, None] * torch.tensor([1.0, -2.0])
    for world, chunks in ((2, 1), (3, 2)):
        plain = [ShardedPack(pack, world, r, 512, 20, n_chunks=chunks) for r in range(world)]
        local = [ShardedPack(pack, world, r, 512, 20, n_chunks=chunks, local_table=True) for r in range(world)]
        for a, b in zip(plain, local):
            assert torch.equal(b.table_ids, torch.unique(torch.cat([lp.gid.long() for lp in a.locals])))
            Xr = X[b.table_ids]
            for la, lb in zip(a.locals, b.locals):
                assert lb.n_src == b.table_ids.numel() and int(lb.gid.max()) < lb.n_src
                assert torch.equal(Xr[lb.gid.long()], X[la.gid.long()])  # same feature rows through the local table
                assert torch.equal(la.col, lb.col) and torch.equal(la.rowptr, lb.rowptr)
        if mode == "none":  # every node's features live on exactly one rank
            assert sorted(torch.cat([s.table_ids for s in local]).tolist()) == list(range(pack.n_nodes))


def _grad_worker(rank, world, port, ret):
    """Uneven shards (rank 1 may hold NO train rows): sum-loss backward + allreduce_gradients(local_count) on every rank
    must equal the gradient of ONE mean loss over all rows (node_train_Gs_GD, run.py:199-204)."""
    import torch.distributed as dist
    from fitgnn_b200.train import allreduce_gradients
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(0)
    X = torch.randn(50, 6, generator=g)
    y = torch.randint(0, 3, (50,), generator=g)
    ok = True
    for split in (35, 50, 1):  # rows [0, split) on rank 0, the rest on rank 1 (split = 50: rank 1 holds none)
        torch.manual_seed(1)
        model = torch.nn.Sequential(torch.nn.Linear(6, 8), torch.nn.ELU(), torch.nn.Linear(8, 3))
        ref = [p.detach().clone() for p in model.parameters()]
        # single-process reference gradient: one mean loss over all 50 rows
        loss = torch.nn.functional.nll_loss(torch.log_softmax(model(X), 1), y)
        want = torch.autograd.grad(loss, list(model.parameters()))
        lo, hi = (0, split) if rank == 0 else (split, 50)
        model.zero_grad()
        out = torch.log_softmax(model(X[lo:hi]), 1)
        n_local = hi - lo
        loss_sum = torch.nn.functional.nll_loss(out, y[lo:hi]) * n_local if n_local > 0 else out.sum() * 0.0
        loss_sum.backward()
        n = allreduce_gradients(model, world, local_count=n_local)
        ok = ok and n == 50.0
        for p, w, r in zip(model.parameters(), want, ref):
            ok = ok and bool(torch.isfinite(p.grad).all()) and torch.allclose(p.grad, w, rtol=1e-5, atol=1e-6)
    ret[rank] = ok
    dist.destroy_process_group()


def test_gradient_allreduce_world2_matches_single_process_mean_loss():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ret = mp.Manager().dict()
    mp.spawn(_grad_worker, args=(2, port, ret), nprocs=2, join=True)
    assert ret[0] and ret[1]
