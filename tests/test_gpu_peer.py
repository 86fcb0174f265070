"""GPU tests of the output exchange fused into the head kernel (fitgnn_gemm_head_rows_peers + CUDA-IPC peer buffers).
The two-rank test runs both ranks as separate processes on cuda:0 (IPC works across processes on one device), with
gloo for the handle exchange and the barrier; nothing spins on the device."""
import os

import numpy as np
import pytest
import torch

from oracle import fitgnn_oracle as fo

pytestmark = pytest.mark.gpu


def dev():
    return torch.device("cuda:0")


@pytest.fixture(scope="module")
def fg():
    import fitgnn_b200
    return fitgnn_b200


def test_head_rows_to_several_buffers(fg):
    """n_peers destination bases (here: three local buffers) all receive every mapped row; unmapped rows untouched."""
    g = torch.Generator().manual_seed(0)
    M, K, N = 700, 512, 47
    A = torch.randn(M, K, generator=g)
    W = torch.randn(N, K, generator=g) / K ** 0.5
    b = torch.randn(N, generator=g)
    keep = torch.rand(M, generator=g) > 0.2
    n_keep = int(keep.sum())
    row_map = torch.full((M,), -1, dtype=torch.int32)
    row_map[keep] = torch.randperm(n_keep, generator=g).to(torch.int32)
    ld = 48
    bufs = [fg.ops.PeerBuffer(4 * (n_keep + 5) * ld, dev()) for _ in range(3)]
    tens = [pb.tensor((n_keep + 5, ld)) for pb in bufs]
    for t in tens:
        t.fill_(3.0)
    fg.ops.gemm_head_rows_peers(fg.ops.split_bf16(A.to(dev())), fg.ops.split_bf16(W.to(dev())), b.to(dev()), fg.ops.ACT_NONE,
                                fg.ops.HEAD_LOG_SOFTMAX, row_map.to(dev()), [pb.ptr for pb in bufs], ld, K=K, N=N)
    z = torch.log_softmax(A.double() @ W.double().T + b.double(), 1)
    want = torch.empty(n_keep, N, dtype=torch.float64)
    want[row_map[keep].long()] = z[keep]
    for t in tens:
        got = t.cpu()
        assert (got[:n_keep, :N].double() - want).abs().max() < 1e-4 * max(1.0, want.abs().max())
        assert (got[n_keep:] == 3.0).all() and ((got[:n_keep, N:] == 3.0) | (got[:n_keep, N:] == 0.0)).all()
    assert torch.equal(tens[0], tens[1]) and torch.equal(tens[0], tens[2])
    del tens
    for pb in bufs:
        pb.close()


def _rank_main(rank, world, port, ret):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import fitgnn_b200 as fg
    from fitgnn_b200.dist import PeerGather, ShardedPack
    from tests.test_gpu_aligned import small_subgraph_graph
    d = torch.device("cuda:0")
    torch.cuda.set_device(0)
    n, F, H, Cc = 4000, 100, 512, 47
    ei, part, k = small_subgraph_graph(n, 1700, 7, max_size=12)
    pack = fg.build_pack(torch.tensor(ei, device=d), torch.tensor(part), k, "none")
    X = fg.synth.features(n, F, seed=1).to(d)
    sd = fo.init_state_dict(F, H, Cc, seed=2)
    shard = ShardedPack(pack, world, rank, H, F)
    f = fg.PackedForward(shard.local, sd, precision="bf16x3", fuse_aggregate=True)
    Cp = fg.ops.pad4(Cc)
    pgather = PeerGather(shard, Cp, d, n_buffers=2)
    for step in range(3):  # round-robin over the two buffers
        b = step % 2
        f(X, peer_ptrs=pgather.slot_ptrs(b, 0))
        torch.cuda.synchronize()
        dist.barrier()
        got = pgather.tensors[b].view(-1, Cp)[shard.node_index(d)][:, :Cc]
        want = fg.PackedForward(pack, sd, precision="bf16x3", fuse_aggregate=True)(X)
        full = torch.empty(n, Cc, device=d)
        full[pack.core_gid.long()] = want
        err = float((got - full).abs().max())
        ret[(rank, step)] = err
        dist.barrier()
    # overlapped exchange: head into the local slot, then the push kernel / the copy engines + barrier on the side stream
    for step in range(6):
        b = step % 2
        pgather.acquire(b)
        pgather.tensors[b].zero_()  # stale data must not survive: only this step's pushes can make the check pass
        torch.cuda.synchronize()
        dist.barrier()
        f(X, out=shard.slot(pgather.tensors[b], 0))
        pgather.exchange_async(b, engine="push" if step < 3 else "ce", push_ctas=3)
        pgather.wait(b)
        torch.cuda.synchronize()
        dist.barrier()
        got = pgather.tensors[b].view(-1, Cp)[shard.node_index(d)][:, :Cc]
        ret[(rank, 10 + step)] = float((got - full).abs().max())
        pgather.release(b)
        dist.barrier()
    # the fp16-hidden-state schedules store through the same peer-mapped head (fitgnn_gemm_f16_head_rows_peers): fp16x2, and
    # 'fp16' (every operand of the wide transforms an fp16 plane: the multi-GPU bench's default arithmetic)
    for slot, prec in ((20, "fp16x2"), (21, "fp16")):
        f16 = fg.PackedForward(shard.local, sd, precision=prec, fuse_aggregate=True)
        assert f16.f16_hidden and f16.f16_layer0 == (prec == "fp16")
        want16 = fg.PackedForward(pack, sd, precision=prec, fuse_aggregate=True)(X)
        full16 = torch.empty(n, Cc, device=d)
        full16[pack.core_gid.long()] = want16
        pgather.tensors[0].zero_()
        torch.cuda.synchronize()
        dist.barrier()
        f16(X, peer_ptrs=pgather.slot_ptrs(0, 0))
        torch.cuda.synchronize()
        dist.barrier()
        got = pgather.tensors[0].view(-1, Cp)[shard.node_index(d)][:, :Cc]
        ret[(rank, slot)] = float((got - full16).abs().max())
        dist.barrier()
    del got
    pgather.tensors = None
    dist.barrier()
    pgather.close()
    dist.destroy_process_group()


def test_two_ranks_exchange_through_peer_buffers(fg):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    port = 29600 + os.getpid() % 300
    procs = [ctx.Process(target=_rank_main, args=(r, 2, port, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    assert len(ret) == 22 and max(ret.values()) < 1e-5, dict(ret)
