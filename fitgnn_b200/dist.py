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
    """Rank-local slices of a pack plus what the all-gather needs.  Subgraphs are dealt into world * n_chunks
    size-balanced units; rank r owns units r*n_chunks .. (r+1)*n_chunks-1, each a small pack of its own, so the
    all-gather of chunk c can run (asynchronously, on NCCL's stream) behind the compute of chunk c+1.  Every rank
    derives every unit's node ids from the replicated full pack, so no metadata is exchanged."""

    def __init__(self, pack: Pack, world: int, rank: int, hidden: int, in_features: int, n_chunks: int = 1,
                 local_table: bool = False):
        """local_table: instead of replicating the whole de-duplicated feature table, every rank keeps only the rows
        its own subgraphs reference — `table_ids` (ascending global feature-row ids; exactly the rank's own nodes in mode
        'none') — and its packs' `gid` index into that table: X_rank = X[table_ids].  A rank's host->device traffic
        and feature memory then shrink by ~1/world and nothing about X has to be exchanged between ranks."""
        rows, nnz = pack_subgraph_sizes(pack)
        self.world, self.rank, self.n_chunks = world, rank, n_chunks
        units = world * n_chunks
        self.bins = balanced_bins(subgraph_costs(rows, nnz, hidden, in_features), units)
        self.sub_ids = [torch.nonzero(self.bins == u).view(-1) for u in range(units)]
        if units == 1:
            self.locals = [pack]
        else:
            self.locals = [select_subgraphs(pack, self.sub_ids[rank * n_chunks + c]) for c in range(n_chunks)]
        self.table_ids = None
        if local_table:
            import dataclasses
            ids = torch.unique(torch.cat([lp.gid.long() for lp in self.locals]))  # sorted ascending
            self.table_ids = ids
            self.locals = [dataclasses.replace(lp, gid=torch.searchsorted(ids, lp.gid.long()).to(torch.int32).contiguous(),
                                               n_src=int(ids.numel())) for lp in self.locals]
        self.local = self.locals[0]
        # core node ids per unit, in that unit's pack order (select_subgraphs keeps ascending subgraph order)
        core_sub = torch.repeat_interleave(torch.arange(pack.n_sub, device=pack.device), rows)[pack.core_rows.long()]
        core_gid = pack.core_gid.long()
        unit_of_core = self.bins[core_sub]
        self.core_ids = [core_gid[unit_of_core == u] for u in range(units)]
        self.counts = [int(c.numel()) for c in self.core_ids]
        self.max_count = max(self.counts) if self.counts else 0
        self.n_nodes = pack.n_nodes
        if units > 1:
            for c in range(n_chunks):
                lg = self.locals[c].core_gid.long()
                lg = self.table_ids[lg] if self.table_ids is not None else lg
                assert torch.equal(lg, self.core_ids[rank * n_chunks + c])
        costs = subgraph_costs(rows, nnz, hidden, in_features)
        self.loads = [float(sum(costs[self.sub_ids[r * n_chunks + c]].sum() for c in range(n_chunks)))
                      for r in range(world)]

    # ---- all-gather of the core-node outputs -------------------------------------------------------------
    def gather_buffer(self, C: int, device, dtype=torch.float32) -> torch.Tensor:
        """[n_chunks, world, max_count, C]; unit (r, c) writes rows [0, counts) of buf[c, r] (the rest is padding).
        Pass `slot(buf, c)` as `out=` to the forward of chunk c so the head kernel writes straight into the buffer."""
        return torch.zeros(self.n_chunks, self.world, self.max_count, C, dtype=dtype, device=device)

    def slot(self, buf: torch.Tensor, chunk: int = 0) -> torch.Tensor:
        return buf[chunk, self.rank, : self.counts[self.rank * self.n_chunks + chunk]]

    def all_gather_(self, buf: torch.Tensor, chunk: int = 0, group=None, async_op: bool = False):
        """In-place all-gather of one chunk: afterwards every rank holds every rank's slot of that chunk.  With
        async_op the NCCL work handle is returned: the collective runs on NCCL's stream behind whatever the caller
        enqueues next, and `handle.wait()` makes the current stream wait for it."""
        if self.world == 1:
            return None
        import torch.distributed as dist
        b = buf[chunk]
        return dist.all_gather_into_tensor(b.view(self.world * self.max_count, -1), b[self.rank], group=group,
                                           async_op=async_op)

    def node_index(self, device) -> torch.Tensor:
        """row_of_node[v] = flat row of node v in the gather buffer (buf.view(-1, C)[row_of_node] is node order)."""
        idx = torch.empty(self.n_nodes, dtype=torch.long, device=device)
        for r in range(self.world):
            for c in range(self.n_chunks):
                u = r * self.n_chunks + c
                base = (c * self.world + r) * self.max_count
                idx[self.core_ids[u].to(device)] = base + torch.arange(self.counts[u], device=device)
        return idx

    def gather_outputs(self, local_outs, group=None) -> torch.Tensor:
        """Convenience form: all-gather per-chunk core outputs (a tensor when n_chunks == 1, else a list) and
        return [N, C] in global node order."""
        outs = [local_outs] if torch.is_tensor(local_outs) else list(local_outs)
        buf = self.gather_buffer(outs[0].shape[1], outs[0].device, outs[0].dtype)
        for c, o in enumerate(outs):
            self.slot(buf, c).copy_(o)
            self.all_gather_(buf, c, group)
        return buf.view(-1, outs[0].shape[1])[self.node_index(outs[0].device)]


class PeerGather:
    """Gather buffers every rank can write: the head kernel of rank r stores its rows into slot (chunk, r) of EVERY
    rank's buffer over NVLink (ops.gemm_head_rows_peers), so the gather needs no collective — only `barrier()`,
    a one-element all-reduce on the current stream that orders every rank's stores before any rank's reads.
    `n_buffers` >= 2 buffers are used round-robin: a rank may start writing buffer b for step i+2 only after the
    barrier of step i+1, which every rank enqueues after its (same-stream) reads of step i's buffer b."""

    def __init__(self, shard: ShardedPack, C: int, device, n_buffers: int = 2, group=None, backend: str = "ipc"):
        """backend 'ipc': cudaMalloc + CUDA IPC handles (fitgnn_peer_*).  backend 'symm': torch symmetric memory (CUDA
        VMM allocations mapped on every rank) which additionally yields an NVLS multicast address: ONE store to it is
        replicated by the NVSwitch into every rank's buffer, so a rank's egress is its own slot once instead of
        once per peer (`slot_ptrs(..., multicast=True)`)."""
        import torch.distributed as dist
        from . import ops
        self.shard, self.C, self.group, self.backend = shard, C, group, backend
        self.shape = (shard.n_chunks, shard.world, shard.max_count, C)
        numel = shard.n_chunks * shard.world * max(shard.max_count, 1) * C
        nbytes = 4 * numel
        self.bufs, self.mc_bases = [], [0] * n_buffers
        if backend == "symm":
            import torch.distributed._symmetric_memory as symm_mem
            stride = (nbytes + 4095) // 4096 * 4096  # every buffer starts on a 4 KB boundary
            self._symm = symm_mem.empty(n_buffers * stride // 4, dtype=torch.float32, device=torch.device(device))
            self._symm.zero_()
            self._hdl = symm_mem.rendezvous(self._symm, dist.group.WORLD if group is None else group)
            self.tensors = [self._symm[i * stride // 4: i * stride // 4 + numel].view(self.shape) for i in range(n_buffers)]
            ptrs = list(self._hdl.buffer_ptrs)
            self.bases = [[int(ptrs[r]) + i * stride for r in range(shard.world)] for i in range(n_buffers)]
            mc = int(getattr(self._hdl, "multicast_ptr", 0) or 0)
            if mc:
                self.mc_bases = [mc + i * stride for i in range(n_buffers)]
        else:
            self.bufs = [ops.PeerBuffer(nbytes, device) for _ in range(n_buffers)]
            self.tensors = [b.tensor(self.shape) for b in self.bufs]
            handles = [None] * shard.world
            dist.all_gather_object(handles, [b.handle for b in self.bufs], group=group)
            # base address of buffer i on rank r, as seen from this process
            self.bases = [[(self.bufs[i].ptr if r == shard.rank else self.bufs[i].open_peer(r, handles[r][i]))
                           for r in range(shard.world)] for i in range(n_buffers)]
        self._flag = torch.zeros(1, device=device)
        self.step = 0
        self._peer_slots = {}
        self.xstream = torch.cuda.Stream(device=device)
        self.pstreams = [torch.cuda.Stream(device=device) for _ in range(max(shard.world - 1, 0))]
        self._ready = [torch.cuda.Event() for _ in range(n_buffers)]
        self._done = [torch.cuda.Event() for _ in range(n_buffers)]
        self._consumed = [None] * n_buffers  # see release()
        for e in self._done:
            e.record(torch.cuda.current_stream())

    @property
    def has_multicast(self):
        return all(b != 0 for b in self.mc_bases)

    def slot_ptrs(self, i: int, chunk: int = 0, multicast: bool = False):
        """Addresses of slot (chunk, my rank) in buffer i of every rank (own rank first).
        multicast=True is refused: the head kernel writes with ordinary st / cp.async.bulk stores, and PTX defines accesses
        to a multimem (NVLS multicast) address only for multimem.ld_reduce / multimem.st / multimem.red — the round-1 'mc'
        exchange worked empirically but relied on undefined behaviour, and it did not beat the peer stores anyway (an
        all-gather is ingress-bound: every rank still receives all the other slots; profiles/r1_multi_gpu.md)."""
        s = self.shard
        off = 4 * ((chunk * s.world + s.rank) * s.max_count) * self.C
        if multicast:
            raise NotImplementedError("multicast stores need multimem.st in the head kernel; use the p2p / ce exchange")
        order = [s.rank] + [r for r in range(s.world) if r != s.rank]
        return [self.bases[i][r] + off for r in order]

    def barrier(self):
        import torch.distributed as dist
        dist.all_reduce(self._flag, group=self.group)

    # ---- copy-engine exchange, overlapped with the next step's compute ------------------------------------------
    def _peer_slot(self, i: int, r: int, chunk: int):
        """slot (chunk, my rank) of buffer i on rank r as a tensor in this process (peer-mapped address)"""
        from . import ops
        key = (i, r, chunk)
        if key not in self._peer_slots:
            s = self.shard
            cnt = s.counts[s.rank * s.n_chunks + chunk]
            off = 4 * ((chunk * s.world + s.rank) * s.max_count) * self.C
            self._peer_slots[key] = ops.raw_tensor(self.bases[i][r] + off, (cnt, self.C), self._flag.device, owner=self)
        return self._peer_slots[key]

    def exchange_async(self, i: int, engine: str = "ce", push_ctas: int = 16, ce_peers: int = 0):
        """After this rank's head kernels wrote slot (c, rank) of its OWN buffer i on the current stream: push the slots to
        every peer on a side stream, then the barrier, all behind whatever the current stream does next (the following step's
        compute).  `wait(i)` makes the current stream wait for buffer i to be complete on every rank; `acquire(i)` must
        precede the next write of it.
        engine 'ce' (default): per-peer cudaMemcpyAsync on one stream per peer (copy engines, no SM time; ~310 GB/s in the
        8-GPU all-to-all pattern).
        engine 'push': ONE small kernel (fitgnn_peer_push, `push_ctas` CTAs launched as clusters of two = whole TPCs) streams the
        slot through shared memory and bulk-stores it to all peers: ~45 GB/s of egress per CTA, ~490 GB/s at 24.  Every SM
        it occupies would hold back one CTA of the persistent 148-CTA GEMM kernels for the whole push (2 GPUs: 3.45-4.70 ms
        against 3.14 ms), so the caller sets the tuning switch `sm_reserve` (= push_ctas) while the push overlaps compute:
        the GEMM grids then leave those SMs free.  Measured (profiles/r2_multi_gpu.md): 8 GPUs 1.28 (ce) -> 0.82 ms per step,
        4 GPUs 1.61 (p2p) -> 1.27 ms.
        `ce_peers` > 0 with engine 'push' (hybrid): the peers at rank distance 1..ce_peers are served by the copy engines (one
        stream each) and only the rest by the push kernel — the two paths have separate limits (copy engines ~300 GB/s, SM egress
        ~20 GB/s per CTA), so together they come closer to the link rate than either alone."""
        s = self.shard
        cur = torch.cuda.current_stream()
        self._ready[i].record(cur)
        if engine == "push" and s.world > 1:
            import ctypes as C
            from . import ops
            from ._lib import check, lib
            ps = self.pstreams[0]
            with torch.cuda.stream(ps):
                ps.wait_event(self._ready[i])
                for c in range(s.n_chunks):
                    src = s.slot(self.tensors[i], c)
                    nbytes = src.numel() * 4
                    if nbytes == 0:
                        continue
                    # peers in order of rank distance: every rank hands the same distances to the copy engines, so every
                    # destination receives ce_peers copy-engine streams and world - 1 - ce_peers push streams
                    order = [(s.rank + d) % s.world for d in range(1, s.world)]
                    dsts = [self._peer_slot(i, r, c).data_ptr() for r in order[ce_peers:]]
                    if dsts:
                        arr = (C.c_void_p * len(dsts))(*[C.c_void_p(d) for d in dsts])
                        check(lib().fitgnn_peer_push(ops.ptr(src), arr, len(dsts), nbytes, push_ctas,
                                                     C.c_void_p(ps.cuda_stream)))
            order = [(s.rank + d) % s.world for d in range(1, s.world)]
            for k, r in enumerate(order[:ce_peers]):
                cs = self.pstreams[1 + k]
                with torch.cuda.stream(cs):
                    cs.wait_event(self._ready[i])
                    for c in range(s.n_chunks):
                        self._peer_slot(i, r, c).copy_(s.slot(self.tensors[i], c), non_blocking=True)
        else:
            # one side stream per peer: a single copy stream reaches ~200 GB/s, several copy engines together ~310 GB/s
            k = 0
            for r in range(s.world):
                if r == s.rank:
                    continue
                ps = self.pstreams[k]
                k += 1
                with torch.cuda.stream(ps):
                    ps.wait_event(self._ready[i])
                    for c in range(s.n_chunks):
                        self._peer_slot(i, r, c).copy_(s.slot(self.tensors[i], c), non_blocking=True)
        with torch.cuda.stream(self.xstream):
            for ps in self.pstreams:
                self.xstream.wait_stream(ps)
            # buffer-reuse ordering: the peers overwrite THIS rank's copy of the other buffers two steps from now, and what
            # orders their pushes after this rank's reads is this barrier — so it must follow every read released so far
            for ev in self._consumed:
                if ev is not None:
                    self.xstream.wait_event(ev)
            self.barrier()
            self._done[i].record(self.xstream)

    def acquire(self, i: int):
        """Before the head kernels overwrite this rank's slots of buffer i: the pushes that read them are finished."""
        torch.cuda.current_stream().wait_event(self._done[i])

    def wait(self, i: int):
        torch.cuda.current_stream().wait_event(self._done[i])

    def release(self, i: int):
        """Call on the consumer's stream after its LAST read of buffer i (as gathered by step s): records the event the
        next barrier waits on, so no peer can push step s+2's rows into this rank's buffer i while it is still being read.
        Protocol per step: acquire(b) -> head kernels -> exchange_async(b) ... wait(b) -> reads -> release(b).  A consumer
        that does not call release() must finish reading buffer i (stream-ordered) before exchange_async of the next step."""
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream())
        self._consumed[i] = ev

    def close(self):
        self._peer_slots = {}
        self.tensors = None
        for b in self.bufs:
            b.close()
