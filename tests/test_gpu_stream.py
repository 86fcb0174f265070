"""GPU: sharded packs + the streamed / hybrid forward (fitgnn_b200.stream), and the config-shaped parity cases of
BASELINE.json that round 1 left to a script: configs[1] (PubMed-shaped, extra_node, r = 0.5, incl. the Gc projection),
configs[3] (ZINC-shaped graph regression through infer.graph_level_Gs), products-small (planted partition, bf16x3, M >= 4096:
CTA-pair GEMM + fused aggregation + row-mapped head against the oracle) and node regression on the reference's fixture.
Tolerance: 1e-3 relative (north_star), integer structure bit-exact."""
import numpy as np
import pytest
import torch

from oracle import fitgnn_oracle as fo
from tests import golden_io as gio
from tests.test_gpu_parity import RTOL, assert_close, dev, fg  # noqa: F401  (fg is a fixture)

pytestmark = pytest.mark.gpu

PACK_ARRAYS = ("rowptr", "col", "dinv", "gid", "sub_ptr", "core_rows", "is_core", "mask")


def planted(fg, n, e, ratio=0.5, seed=0, **kw):
    ei, part, cw, k = fg.synth.planted_partition(n, e, ratio, seed=seed, device=dev(), **kw)
    return ei, fg.synth.relabel_partition_reference_order(part), cw, k


def oracle_none(ei, X, part, k, sd, sub_ids=None):
    """Reference path for mode 'none' on (a sample of) the clusters: subgraphs + 128-subgraph batches (run.py:49-115)."""
    ids = np.arange(k) if sub_ids is None else np.asarray(sub_ids)
    subs = fo.subgraphs_from_partition(ei.cpu().numpy(), X.cpu().numpy(), part.cpu().numpy(), ids)
    sel = [np.ones(s["x"].shape[0], dtype=bool) for s in subs]
    return fo.node_infer_batched({k_: v.cpu() for k_, v in sd.items()}, subs, sel, "node_cls", 128).numpy()


# ------------------------------------------------------------------------------------------ sharded packs
@pytest.mark.parametrize("mode", ["none", "cluster"])
def test_pack_stream_shards_equal_the_whole_pack(fg, mode):
    ei, part, cw, k = planted(fg, 20000, 300000)
    whole = fg.build_pack(ei, part, k, mode)
    stream = fg.build_pack_stream(ei, part, k, mode, max_rows=max(2000, whole.n_rows // 5))
    assert len(stream.packs) >= 4 and stream.bounds[0] == 0 and stream.bounds[-1] == k
    assert stream.n_rows == whole.n_rows and stream.nnz == whole.nnz and stream.n_core == whole.n_core
    for shard, a, b in zip(stream.packs, stream.bounds[:-1], stream.bounds[1:]):
        want = fg.infer.select_subgraphs(whole, torch.arange(a, b, device=dev()))
        for name in PACK_ARRAYS:
            assert torch.equal(getattr(shard, name), getattr(want, name)), (name, a, b)
        assert shard.n_src == whole.n_src
    assert torch.equal(stream.core_gid, whole.core_gid)


def cluster_table(fg, ei, part, cw, k, X):
    members, member_ptr = fg.ops.group_by_part(part, k)
    return torch.cat([X, fg.ops.project_features(members, member_ptr, cw, X)], 0)


@pytest.mark.parametrize("mode,precision", [("none", "bf16x3"), ("cluster", "bf16x3"), ("cluster", "fp32")])
def test_streamed_forward_equals_whole_pack_forward(fg, mode, precision):
    n, F, C = 12000, 100, 47
    ei, part, cw, k = planted(fg, n, 150000, seed=1)
    X = fg.synth.features(n, F, seed=1, device=dev())
    Xg = cluster_table(fg, ei, part, cw, k, X) if mode == "cluster" else X
    sd = fg.synth.init_state_dict(F, 512, C, seed=1)
    whole = fg.build_pack(ei, part, k, mode)
    want = fg.PackedForward(whole, sd, precision=precision, fuse_aggregate=False)(Xg)
    stream = fg.build_pack_stream(ei, part, k, mode, max_rows=whole.n_rows // 4)
    fwd = fg.StreamedForward(stream, sd, precision=precision, hybrid=False)
    assert len(fwd.parts) >= 4
    got = fwd(Xg)
    assert got.shape == want.shape
    if mode == "none":
        assert torch.equal(got, want[:, :C])  # same kernels, same per-row arithmetic -> bit-identical
    else:  # dense packs aggregate on the tensor cores: the summation order depends on a subgraph's offset inside its block
        assert float((got - want[:, :C]).abs().max()) <= 1e-5 * float(want.abs().max())


def test_hybrid_heavy_tailed_pack_matches_oracle_and_classic(fg):
    """Power-law subgraph sizes (max 500): subgraphs <= 32 rows take the fused group-aligned schedule, the rest the classic
    one, both writing one output through row maps; against the oracle (all subgraphs) and the classic whole-pack forward."""
    n, F, C = 40000, 100, 47
    ei, part, cw, k = planted(fg, n, 600000, seed=2, sizes="powerlaw")
    sizes = torch.bincount(part.long())
    assert int(sizes.max()) > 200 and int((sizes <= 32).sum()) > 0
    X = fg.synth.features(n, F, seed=2, device=dev())
    sd = fg.synth.init_state_dict(F, 512, C, seed=2)
    pack = fg.build_pack(ei, part, k, "none")
    fwd = fg.StreamedForward(pack, sd)
    assert sorted(fwd.kinds) == ["classic", "fused"]
    got = fwd(X)
    classic = fg.PackedForward(pack, sd, fuse_aggregate=False)(X)
    assert float((got - classic[:, :C]).abs().max()) <= 1e-4 * float(classic.abs().max())
    assert_close(got.cpu().numpy(), oracle_none(ei, X, part, k, sd))
    # a forced single fused part refuses the oversized subgraphs loudly
    with pytest.raises(ValueError):
        fg.PackedForward(pack, sd, fuse_aggregate=True)


def test_products_small_fused_schedule_matches_oracle(fg):
    """The headline schedule (group-aligned pack, spmm0 -> transform + fused aggregation -> CTA-pair transform -> row-mapped
    head, bf16x3) at M >= 4096 rows, end to end against the oracle — the path bench.py times on the products workload."""
    n, F, C = 30000, 100, 47
    ei, part, cw, k = planted(fg, n, 750000, seed=3)
    X = fg.synth.features(n, F, seed=3, device=dev())
    sd = fg.synth.init_state_dict(F, 512, C, seed=3)
    pack = fg.build_pack(ei, part, k, "none")
    fwd = fg.PackedForward(pack, sd, precision="bf16x3", fuse_aggregate=True)
    assert fwd.apack is not None and fwd.apack.n_rows >= 4096
    want = oracle_none(ei, X, part, k, sd)
    assert_close(fwd(X).cpu().numpy(), want)
    assert_close(fwd(fwd.pack_features(X), packed=True).cpu().numpy(), want)


def test_cluster_mode_stream_matches_oracle_on_a_sample(fg):
    """cluster_node augmentation through the sharded pack + streamed forward against the reference's subgraph builder
    (utils.py:190-233, 251-259 restated in the oracle) on a sample of subgraphs."""
    import scipy.sparse as sp_
    n, F, C = 6000, 100, 47
    ei, part, cw, k = planted(fg, n, 60000, seed=4)
    X = fg.synth.features(n, F, seed=4, device=dev())
    sd = fg.synth.init_state_dict(F, 512, C, seed=4)
    row, col, cnt, ac_rowptr = fg.ops.project_adj(ei, part, k)
    Xg = cluster_table(fg, ei, part, cw, k, X)
    stream = fg.build_pack_stream(ei, part, k, "cluster", max_rows=30000, ac_rowptr=ac_rowptr, ac_col=col.to(torch.int32))
    assert len(stream.packs) >= 3
    out = fg.StreamedForward(stream, sd)(Xg).cpu().numpy()
    # oracle on a sample: the planted graph is handled as one component whose coarsening is `part`
    rng = np.random.default_rng(0)
    ids = np.sort(rng.choice(k, 60, replace=False))
    adj = sp_.csr_matrix((np.ones(row.numel(), dtype=bool), (row.cpu().numpy(), col.cpu().numpy())), shape=(k, k))
    co = dict(part=part.cpu().numpy().astype(np.int64), CX=Xg[n:].cpu().numpy(), adj=adj)
    subs = fo.build_subgraphs(ei.cpu().numpy(), X.cpu().numpy(), np.zeros(n, dtype=np.int64), [np.arange(n)], [co], "cluster",
                              only=set(ids.tolist()))
    sel = []
    for i in ids:
        m = np.zeros(subs[i]["x"].shape[0], dtype=bool)
        m[: len(subs[i]["core"])] = True
        sel.append(m)
    want = fo.node_infer_batched({k_: v.cpu() for k_, v in sd.items()}, [subs[i] for i in ids], sel, "node_cls", 128).numpy()
    # output rows of the sampled subgraphs: core rows are in subgraph order, nodes ascending inside a subgraph
    core_sub = part.long()[stream.core_gid.long()].cpu().numpy()
    assert (np.diff(core_sub) >= 0).all()
    rows = np.nonzero(np.isin(core_sub, ids))[0]
    assert_close(out[rows], want)


# ------------------------------------------------------------------------------------------ config-shaped parity cases
def test_pubmed_shaped_config_extra_with_gc(fg):
    """configs[1]: PubMed-shaped synthetic (19,717 nodes, 44,324 undirected edges, 500 features, 3 classes),
    Gc_train_2_Gs_infer with extra_node, r = 0.5: the Gs pack + logits and the Gc projection (Xc, Ac) vs the oracle."""
    n, e_und, F, C, ratio = fg.synth.SHAPES["pubmed"]
    ei = fg.synth.powerlaw_graph(n, e_und, seed=0)
    partition, comps, C_list = fg.synth.neighborhood_partition(ei, n, ratio, seed=0)
    X = fg.synth.features(n, F, seed=0, kind="bow")
    cos = [None if Cm is None else dict(part=fo.partition_of(Cm)[0], CX=None, adj=None) for Cm in C_list]
    eid = torch.tensor(ei, device=dev())
    # Gc: Xc bit-exact vs scipy (fp64 accumulate -> fp32), Ac pattern + counts bit-exact, per component
    proj = fg.coarsen.project(eid, X.to(dev()), partition)
    xc = proj["Xc"].cpu().numpy()
    ac_rp = proj["ac_rowptr"].cpu().numpy()
    for i in range(min(3, len(comps))):  # the giant component and the next ones
        comp, Cm = comps[i], C_list[i]
        if Cm is None:
            continue
        s0, s1 = int(partition.sub_offset[i]), int(partition.sub_offset[i + 1])
        assert np.array_equal(xc[s0:s1], fo.project_features(Cm, X.numpy()[comp]).astype(np.float32))
        relabel = np.full(n, -1); relabel[comp] = np.arange(len(comp))
        em = relabel[ei[0]] >= 0
        r, c, v = fo.project_adj_pattern(relabel[ei[:, em]], fo.partition_of(Cm)[0], Cm.shape[0])
        e0, e1 = int(ac_rp[s0]), int(ac_rp[s1])
        assert np.array_equal(proj["ac_row"][e0:e1].cpu().numpy() - s0, r)
        assert np.array_equal(proj["ac_col"][e0:e1].cpu().numpy() - s0, c)
        assert np.array_equal(proj["ac_cnt"][e0:e1].cpu().numpy(), np.rint(v).astype(np.int32))
    # Gs: a sample of subgraphs (the oracle builder rescans all edges per subgraph) bit-exact + logits
    pack = fg.build_pack(eid, torch.tensor(partition.part), partition.k, "extra")
    assert sorted(pack.core_gid.cpu().tolist()) == list(range(n))
    rng = np.random.default_rng(1)
    sizes = (pack.sub_ptr[1:] - pack.sub_ptr[:-1]).cpu().numpy()
    ids = np.unique(np.concatenate([np.argsort(-sizes)[:3], rng.choice(pack.n_sub, 300, replace=False)]))
    subs = fo.build_subgraphs(ei, X.numpy(), np.zeros(n, dtype=np.int64), comps, cos, "extra", only=set(ids.tolist()))
    small = fg.infer.select_subgraphs(pack, torch.tensor(ids))
    want = fo.expected_pack([subs[i] for i in ids], n, "extra")
    for name in ("rowptr", "col", "gid", "sub_ptr", "core_rows", "is_core", "mask"):
        assert np.array_equal(getattr(small, name).cpu().numpy().astype(np.int64), np.asarray(want[name]).astype(np.int64)), name
    sd = fo.init_state_dict(F, 512, C, seed=2)
    out, node_ids = fg.infer.node_infer_Gs(sd, pack, X.to(dev()))
    sel = []
    for i in ids:
        m = np.zeros(subs[i]["x"].shape[0], dtype=bool)
        m[np.searchsorted(subs[i]["orig_idx"], subs[i]["core"])] = True
        sel.append(m)
    want_out = fo.node_infer_batched(sd, [subs[i] for i in ids], sel, "node_cls", 128).numpy()
    core_sub = torch.tensor(partition.part).long()[node_ids.cpu().long()].numpy()
    rows = np.nonzero(np.isin(core_sub, ids))[0]
    assert_close(out.cpu().numpy()[rows], want_out)


def test_zinc_shaped_config_graph_regression(fg):
    """configs[3]: ZINC-shaped batched small-graph regression with coarsening per graph (main.py:370-381): every
    subgraph of every graph in ONE pack, infer.graph_level_Gs (conv stack -> x[mask] -> mean pool per graph -> lt1)
    against the reference's double loop (network.py:189-204) restated in the oracle."""
    n_graphs, ratio = 400, 0.3
    graphs = fg.synth.molecule_graphs(n_graphs, seed=0)
    xs, eis, parts, graph_of_sub, off, sub_off, set_gs, bt = [], [], [], [], 0, 0, [], []
    for gi, (x, ei, y) in enumerate(graphs):
        n = x.shape[0]
        partition, comps, C_list = fg.synth.neighborhood_partition(ei, n, ratio, seed=gi)
        xs.append(x); eis.append(ei + off); parts.append(partition.part + sub_off)
        graph_of_sub.extend([gi] * partition.k)
        cos = [None if Cm is None else dict(part=fo.partition_of(Cm)[0], CX=None, adj=None) for Cm in C_list]
        subs = fo.build_subgraphs(ei, x.astype(np.float32), np.zeros(n, dtype=np.int64), comps, cos, "extra")
        set_gs.append([dict(x=torch.tensor(s["x"]), edge_index=torch.tensor(s["edge_index"]), mask=torch.tensor(s["mask"]))
                       for s in subs])
        bt.extend([gi] * n)
        off += n; sub_off += partition.k
    X = torch.tensor(np.concatenate(xs)).float().to(dev())
    pack = fg.build_pack(torch.tensor(np.concatenate(eis, 1), device=dev()), torch.tensor(np.concatenate(parts)), sub_off, "extra")
    for task, C in (("graph_reg", 1), ("graph_cls", 4)):
        sd = fg.synth.init_state_dict(1, 512, C, seed=7)
        with torch.no_grad():
            want = fo.graph_gs_forward(sd, set_gs, torch.tensor(bt), task).numpy()
        for precision in ("bf16x3", "fp32"):
            got = fg.infer.graph_level_Gs(sd, pack, X, torch.tensor(graph_of_sub), task, precision=precision)
            assert_close(got.cpu().numpy(), want)


@pytest.mark.parametrize("mode", ["none", "extra", "cluster"])
def test_node_regression_fixture_on_cuda(fg, mode):
    """Regress_node on the reference's own node-regression fixture (node_reg_small.npz: coarsening_regression(task='node_reg'),
    load_data_regression, Regress_node over the run.py:59-77 batch loop executed unmodified): packed whole-pack inference on
    the GPU against the REFERENCE's outputs of the test rows, default (tensor-core) and exact-fp32 arithmetic."""
    from tests.test_gpu_parity import global_features, golden_partition
    d = gio.load("node_reg_small")
    comps, cos, partition = golden_partition(fg, d, mode)
    pack = fg.build_pack(torch.tensor(d["edge_index"], device=dev()), torch.tensor(partition.part), partition.k, mode)
    X = global_features(d, mode, cos, comps, partition).to(dev())
    sd = gio.state_dict(d)
    assert sd["lt1.weight"].shape[0] == 1
    for precision in ("bf16x3", "fp32"):
        out, ids = fg.infer.node_infer_Gs(sd, pack, X, torch.tensor(d["test_mask"]), task="node_reg", precision=precision)
        assert_close(out.cpu().numpy(), d[f"{mode}_test_out"])


@pytest.mark.parametrize("case", ["graph_small"])  # the fixture that records the per-graph coarsening matrices C
def test_load_graph_data_matches_reference(fg, case):
    """a12, graph tasks: `load_graph_data` (utils.py:811-852) — per graph and batched over the whole fixture in ONE device
    projection — against the Gc graphs the reference built (g{g}_gc_x bit-exact, g{g}_gc_edge in the same order), and the
    *_gc model on the batched result against the reference's predictions."""
    import argparse
    from oracle.ref_shims import Data
    d = gio.load(case)
    n_g = int(d["n_kept"])
    eis, xs, partitions = [], [], []
    for g in range(n_g):
        pre = f"g{g}"
        comps = gio.components(d, pre)
        recs = iter(gio.records(d, pre))
        n = int(d[pre + "_n"])
        partition = fg.coarsen.partition_from_components(comps, [next(recs)["C"] if len(c) > 1 else None for c in comps], n)
        ei, x = d[pre + "_ei"], d[pre + "_x"].reshape(n, -1)
        one = fg.coarsen.load_graph_data(torch.tensor(ei, device=dev()), torch.tensor(x).float().to(dev()), None, partition, comps)
        assert np.array_equal(one["x"].cpu().numpy(), d[pre + "_gc_x"].reshape(partition.k, -1).astype(np.float32))
        assert np.array_equal(one["edge_index"].cpu().numpy(), d[pre + "_gc_edge"])
        eis.append(ei); xs.append(x); partitions.append(partition)
    gc = fg.coarsen.load_graph_data_batch(eis, xs, partitions, dev())
    off = 0
    for g in range(n_g):
        a, b = int(gc["ptr"][g]), int(gc["ptr"][g + 1])
        assert np.array_equal(gc["x"][a:b].cpu().numpy(), d[f"g{g}_gc_x"].reshape(b - a, -1).astype(np.float32))
        m = (gc["edge_index"][0] >= a) & (gc["edge_index"][0] < b)
        assert np.array_equal(gc["edge_index"][:, m].cpu().numpy() - a, d[f"g{g}_gc_edge"])
    assert np.array_equal(gc["batch"].cpu().numpy(), d["gc_batch"])
    sd = gio.state_dict(d)
    cls = case == "graph_cls_small"
    args = argparse.Namespace(num_layers1=2, num_features=gc["x"].shape[1], hidden=int(d["hidden"]),
                              num_classes=int(d["n_classes"]) if cls else 1, layer_name="GCNConv")
    model = (fg.Classify_graph_gc if cls else fg.Regress_graph_gc)(args)
    model.load_state_dict(sd)
    pred = model.to(dev()).eval()(Data(x=gc["x"], edge_index=gc["edge_index"], batch=gc["batch"]))
    assert_close(pred.detach().cpu().numpy(), d["pred_gc"])
    # a graph whose largest component is a single node "does not need coarsening" (utils.py:841)
    with pytest.raises(Exception, match="does not need coarsening"):
        fg.coarsen.load_graph_data(torch.zeros(2, 0, dtype=torch.long, device=dev()), torch.ones(2, 1, device=dev()), None,
                                   fg.coarsen.partition_from_components([np.array([0]), np.array([1])], [None, None], 2),
                                   [np.array([0]), np.array([1])])


@pytest.mark.parametrize("mode,F,H,C,head", [("none", 100, 512, 47, "log_softmax"), ("extra", 100, 512, 47, "log_softmax"),
                                             ("cluster", 30, 64, 5, "softmax"), ("extra", 600, 128, 1, "identity"),
                                             ("none", 7, 32, 3, "log_softmax")])
@pytest.mark.parametrize("precision", ["bf16x3", "fp32"])
def test_one_call_c_forward_equals_the_engine(fg, mode, F, H, C, head, precision):
    """fitgnn_gcn_forward (the whole schedule behind one C entry point, no host synchronisation) against the Python engine's
    classic schedule (bit-identical: same kernels in the same order) and the oracle; aggregate-first and transform-first
    (F > H), all modes, hub rows included."""
    n = 5000
    ei = fg.synth.powerlaw_graph(n, 20000, seed=F)  # power-law degrees: cluster mode gets hub rows (deg >= 256)
    partition, comps, C_list = fg.synth.neighborhood_partition(ei, n, 0.3, seed=F)
    eid = torch.tensor(ei, device=dev())
    X = fg.synth.features(n, F, seed=F, device=dev())
    pack = fg.build_pack(eid, torch.tensor(partition.part), partition.k, mode)
    Xg = X
    if mode == "cluster":
        Xg = torch.cat([X, fg.coarsen.project(eid, X, partition)["Xc"]], 0)
    sd = {k_: v.to(dev()) for k_, v in fg.synth.init_state_dict(F, H, C, seed=F).items()}  # on the device: no copies in capture
    prec = fg.ops.GEMM_BF16X3 if precision == "bf16x3" else fg.ops.GEMM_FP32
    hd = {"identity": fg.ops.HEAD_IDENTITY, "log_softmax": fg.ops.HEAD_LOG_SOFTMAX, "softmax": fg.ops.HEAD_SOFTMAX}[head]
    got = fg.ops.gcn_forward(pack, Xg, sd, hd, prec)
    want = fg.PackedForward(pack, sd, head=head, precision=precision, fuse_aggregate=False)(Xg)
    assert torch.equal(got, want[:, :C])
    # 3 layers + the oracle on a sample of subgraphs (mode none: every row is a core row)
    if mode == "none":
        sd3 = fg.synth.init_state_dict(F, H, C, num_layers=3, seed=1)
        got3 = fg.ops.gcn_forward(pack, Xg, sd3, hd, prec)
        if head != "log_softmax":
            return
        ids = np.arange(min(partition.k, 400))
        subs = fo.subgraphs_from_partition(ei, X.cpu().numpy(), partition.part, ids)
        sel = [np.ones(s["x"].shape[0], dtype=bool) for s in subs]
        ref = fo.node_infer_batched({k_: v.cpu() for k_, v in sd3.items()}, subs, sel, "node_cls", 128).numpy()
        assert_close(got3[: ref.shape[0]].cpu().numpy(), ref)
    # a CUDA graph of the call replays to the same result (no host synchronisation inside)
    g = torch.cuda.CUDAGraph()
    out = torch.empty(pack.n_core, (C + 3) // 4 * 4, device=dev())
    s_ = torch.cuda.Stream()
    s_.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s_):
        fg.ops.gcn_forward(pack, Xg, sd, hd, prec, out=out)
    torch.cuda.current_stream().wait_stream(s_)
    with torch.cuda.graph(g):
        fg.ops.gcn_forward(pack, Xg, sd, hd, prec, out=out)
    out.zero_()
    g.replay()
    torch.cuda.synchronize()
    assert torch.equal(out[:, :C], got)


@pytest.mark.parametrize("width", [4, 100, 104, 128, 132, 512, 1540])
@pytest.mark.parametrize("split", [False, True])
def test_blocked_spmm_is_bit_identical_to_generic_spmm(fg, width, split):
    """fitgnn_spmm_symnorm_blocked (sources staged per block of whole subgraphs) against fitgnn_spmm_symnorm: heavy-tailed
    subgraph sizes, so small blocks (staged), blocks of several subgraphs and blocks beyond the staging capacity (read
    from global memory) all occur; with and without the gid indirection, bias + ELU, fp32 and bf16 hi/lo outputs."""
    if split and width % 8 != 0:
        pytest.skip("bf16 planes need a pitch that is a multiple of 8")
    n = 30000
    ei, part, cw, k = planted(fg, n, 300000, seed=5, sizes="powerlaw")
    pack = fg.build_pack(ei, part, k, "none")
    g = torch.Generator(device="cuda").manual_seed(width)
    X = torch.randn(n, width, generator=g, device=dev())
    bias = torch.randn(width, generator=g, device=dev())
    for window in (16, 64):
        blk = fg.ops.row_blocks(pack.sub_ptr, pack.n_rows, window)
        sizes = (blk[1:] - blk[:-1])
        assert int(sizes.max()) > 300 and int((sizes == 0).sum()) > 0 and int(sizes.sum()) == pack.n_rows
        for src in (None, pack.gid):
            Xin = X if src is not None else X[pack.gid.long()].contiguous()
            for b, act in ((None, fg.ops.ACT_NONE), (bias, fg.ops.ACT_ELU)):
                a = fg.ops.spmm_symnorm(pack.rowptr, pack.col, pack.dinv, Xin, width, src, b, act, split=split)
                for order in (None, fg.ops.block_row_order(pack.rowptr, blk)):
                    c = fg.ops.spmm_symnorm_blocked(pack.rowptr, pack.col, pack.dinv, Xin, blk, width, src, b, act, split=split,
                                                    row_order=order)
                    if split:
                        assert torch.equal(a[0], c[0]) and torch.equal(a[1], c[1])
                    else:
                        assert torch.equal(a, c)


def test_blocked_spmm_cluster_pack_and_engine_switch(fg):
    """cluster_node pack (dense rows: the case the staging is for): blocked == generic, and the engine picks it by itself."""
    n, F = 8000, 100
    ei, part, cw, k = planted(fg, n, 120000, seed=6)
    X = fg.synth.features(n, F, seed=6, device=dev())
    Xg = cluster_table(fg, ei, part, cw, k, X)
    pack = fg.build_pack(ei, part, k, "cluster")
    assert pack.nnz > 8 * pack.n_rows
    blk = fg.ops.row_blocks(pack.sub_ptr, pack.n_rows)
    a = fg.ops.spmm_symnorm(pack.rowptr, pack.col, pack.dinv, Xg, F, pack.gid, split=False)
    c = fg.ops.spmm_symnorm_blocked(pack.rowptr, pack.col, pack.dinv, Xg, blk, F, pack.gid, split=False)
    assert torch.equal(a, c)
    for width, src, Xin in ((F, pack.gid, Xg), (512, None, torch.randn(pack.n_rows, 512, device=dev())), (36, None, torch.randn(pack.n_rows, 36, device=dev()))):
        ref = fg.ops.spmm_symnorm(pack.rowptr, pack.col, pack.dinv, Xin, width, src)
        bias = torch.randn(width, device=dev())
        for split in (False, True):
            if split and width % 8:
                continue
            m = fg.ops.spmm_symnorm_mma(pack.rowptr, pack.col, pack.dinv, Xin, blk, width, src, split=split)
            m = (m[0].float() + m[1].float()) if split else m
            assert float((m - ref).abs().max()) <= 2e-5 * float(ref.abs().max()), (width, split)  # bf16 hi/lo sources: 2^-17
        refb = fg.ops.spmm_symnorm(pack.rowptr, pack.col, pack.dinv, Xin, width, src, bias, fg.ops.ACT_ELU)
        mb = fg.ops.spmm_symnorm_mma(pack.rowptr, pack.col, pack.dinv, Xin, blk, width, src, bias, fg.ops.ACT_ELU)
        assert float((mb - refb).abs().max()) <= 2e-5 * float(refb.abs().max())
    # blocks with more than 128 rows (tiled M pieces) and duplicate edges (counts > 1)
    big = fg.ops.row_blocks(pack.sub_ptr, pack.n_rows, window=512)
    assert int((big[1:] - big[:-1]).max()) > 300
    ref = fg.ops.spmm_symnorm(pack.rowptr, pack.col, pack.dinv, Xg, F, pack.gid)
    m = fg.ops.spmm_symnorm_mma(pack.rowptr, pack.col, pack.dinv, Xg, big, F, pack.gid)
    assert float((m - ref).abs().max()) <= 2e-5 * float(ref.abs().max())
    sd = fg.synth.init_state_dict(F, 512, 47, seed=6)
    on = fg.PackedForward(pack, sd)
    lds = fg.PackedForward(pack, sd, dense_spmm="lds")
    off = fg.PackedForward(pack, sd, blocked_spmm=False)
    assert on._blk is not None and off._blk is None
    assert float((lds(Xg) - off(Xg)).abs().max()) <= 1e-5 * float(off(Xg).abs().max())
    # (not bit-identical in general: the generic path hands rows with >= 256 entries to the hub kernel, which sums them
    # in a different order)
    a_, b_ = on(Xg), off(Xg)
    assert float((a_ - b_).abs().max()) <= 1e-5 * float(b_.abs().max())


def test_pack_from_reference_cache_runs_the_cached_lists(fg, tmp_path):
    """A cache directory in the reference's layout (main.py:131-172) -> pack -> forward: node task against the reference's own
    outputs of the same subgraph list, graph task through infer.graph_level_Gs against the reference's predictions."""
    from oracle import ref_shims
    ref_shims.install()
    from fitgnn_b200 import cache as fc
    d = gio.load("node_small")
    subs = [ref_shims.Data(x=torch.tensor(s["x"]), edge_index=torch.tensor(s["edge_index"]), mask=torch.tensor(s["mask"]),
                           orig_idx=torch.tensor(s["orig_idx"]), test_mask=torch.tensor(s["test_mask"]))
            for s in gio.subgraphs(d, "cluster_sub")]
    fc.save_reference_cache(str(tmp_path), "cora", "variation_neighborhoods", 0.3, "node_reg", subs, cluster_node=True)
    cache = fc.load_reference_cache(str(tmp_path), "cora", "variation_neighborhoods", 0.3, cluster_node=True)
    pack, X, node_ids = fc.pack_from_reference_cache(cache, dev())
    out = fg.PackedForward(pack, gio.state_dict(d), rows="all")(X)
    test_rows = torch.cat([s.test_mask for s in subs])
    assert_close(out[test_rows.to(dev())].cpu().numpy(), d["cluster_test_out"])
    g = gio.load("graph_small")
    n_g = int(g["n_kept"])
    glist = [[ref_shims.Data(x=torch.tensor(s["x"]), edge_index=torch.tensor(s["edge_index"]), mask=torch.tensor(s["mask"]))
              for s in gio.subgraphs(g, f"g{i}_sub")] for i in range(n_g)]
    fc.save_reference_cache(str(tmp_path), "zinc", "variation_neighborhoods", 0.3, "graph_reg", glist, Gc_list=[None] * n_g,
                            saved_graph_list=list(range(n_g)), extra_node=True)
    gc = fc.load_reference_cache(str(tmp_path), "zinc", "variation_neighborhoods", 0.3, extra_node=True)
    gpack, gX, graph_of_sub = fc.pack_from_reference_cache(gc, dev())
    pred = fg.infer.graph_level_Gs(gio.state_dict(g), gpack, gX, graph_of_sub, "graph_reg")
    assert_close(pred.cpu().numpy(), g["pred_gs"])


# ------------------------------------------------------------------------------------------ fp16 hidden state (opt-in)
@pytest.mark.parametrize("M,K,N", [(1000, 512, 512), (5000, 512, 47), (4224, 192, 384), (100003, 512, 512)])
def test_gemm_fp16x2_matches_fp64_of_the_rounded_operand(fg, M, K, N):
    """FITGNN_GEMM_FP16X2: A = ONE fp16 plane, W = fp16 hi/lo; the product of the ROUNDED A with the exact W must come out to
    fp32 accuracy (the only precision given up is the rounding of A to 11 bits, which the caller chose)."""
    g = torch.Generator().manual_seed(M + N)
    A = torch.randn(M, K, generator=g)
    W = torch.randn(N, K, generator=g) / K ** 0.5
    b = torch.randn(N, generator=g) * 0.1
    rs = torch.rand(M, generator=g) + 0.5
    a16, _ = fg.ops.split_f16(A.to(dev()), lo=False)
    Wp = fg.ops.split_f16(W.to(dev()))
    want = torch.nn.functional.elu(rs.double()[:1500, None] * (a16[:1500].cpu().double() @ W.double().T) + b.double())
    got = fg.ops.gemm_f16(a16, Wp, b.to(dev()), fg.ops.ACT_ELU, row_scale=rs.to(dev()))
    assert got.dtype == torch.float32
    assert float((got[:1500, :N].cpu().double() - want).abs().max()) <= 2e-5 * float(want.abs().max())
    if N % 8 == 0:  # fp16-plane output = the fp32 result rounded once
        got16 = fg.ops.gemm_f16(a16, Wp, b.to(dev()), fg.ops.ACT_ELU, row_scale=rs.to(dev()), out_f16=True)
        assert got16.dtype == torch.float16 and torch.equal(got16, got[:, :N].half())
    # the rounding itself: 2^-11 relative per element of A
    assert float((a16.float().cpu() - A).abs().max()) <= 2.0 ** -11 * float(A.abs().max())


def test_fp16_hidden_state_schedule_matches_oracle(fg):
    """PackedForward(precision='fp16x2') on the headline schedule: logits against the oracle inside the 1e-3 bound with a wide
    margin (measured ~2e-5 of max |log-prob|), and close to the bf16x3 result; ineligible packs fall back to bf16x3."""
    n, F, C = 30000, 100, 47
    ei, part, cw, k = planted(fg, n, 750000, seed=3)
    X = fg.synth.features(n, F, seed=3, device=dev())
    sd = fg.synth.init_state_dict(F, 512, C, seed=3)
    pack = fg.build_pack(ei, part, k, "none")
    f16 = fg.PackedForward(pack, sd, precision="fp16x2")
    assert f16.f16_hidden and f16.apack is not None
    want = oracle_none(ei, X, part, k, sd)
    got = f16(X).cpu().numpy()
    assert_close(got, want)
    assert np.abs(got - want).max() <= 2e-4 * max(1.0, np.abs(want).max())
    ref = fg.PackedForward(pack, sd, precision="bf16x3")(X).cpu().numpy()
    assert np.abs(got - ref).max() <= 2e-4 * np.abs(ref).max()
    # three layers: the middle fused transform takes the fp16 plane as input
    sd3 = fg.synth.init_state_dict(F, 256, C, num_layers=3, seed=4)
    g3 = fg.PackedForward(pack, sd3, precision="fp16x2")(X).cpu().numpy()
    r3 = fg.PackedForward(pack, sd3, precision="bf16x3")(X).cpu().numpy()
    assert np.abs(g3 - r3).max() <= 3e-4 * np.abs(r3).max()
    # a pack the fused schedule cannot take (cluster mode) runs the classic schedule on the fp16 plane
    cl = fg.PackedForward(fg.build_pack(ei[:, :100000], part, k, "cluster"), sd, precision="fp16x2")
    assert cl.f16_hidden and cl.f16_classic and cl.apack is None
    # a model the fp16 plane does not cover (one layer) silently runs bf16x3
    sd1 = fg.synth.init_state_dict(F, 512, C, num_layers=1, seed=4)
    assert not fg.PackedForward(pack, sd1, precision="fp16x2", fuse_aggregate=False).f16_hidden


@pytest.mark.parametrize("M,K,N", [(1000, 512, 512), (5000, 512, 512), (4224, 192, 384), (100003, 512, 512), (70001, 320, 512),
                                   (50000, 512, 47)])
def test_gemm_fp16_single_weight_plane_matches_fp64_of_the_rounded_operands(fg, M, K, N):
    """W_lo = NULL (precision='fp16'): A and W are ONE fp16 plane each, one MMA per k-step.  The product of the ROUNDED operands
    must come out to fp32 accuracy on every plan the shapes select: single-CTA streaming (M < 4096), streaming CTA pairs (too
    few row blocks for resident weights), W-stationary CTA pairs (M >= 148 * 256; K = 512 and K = 320), the narrow head; and the
    W-stationary pair plan must be bit-identical to the streaming one."""
    g = torch.Generator().manual_seed(M + N + 1)
    A = torch.randn(M, K, generator=g)
    W = torch.randn(N, K, generator=g) / K ** 0.5
    b = torch.randn(N, generator=g) * 0.1
    rs = torch.rand(M, generator=g) + 0.5
    a16, _ = fg.ops.split_f16(A.to(dev()), lo=False)
    w16, none = fg.ops.split_f16(W.to(dev()), lo=False)
    assert none is None
    rows = torch.cat([torch.arange(700), torch.arange(M - 700, M)]) if M > 1400 else torch.arange(M)
    want = torch.nn.functional.elu(rs.double()[rows, None] * (a16[rows].cpu().double() @ w16.cpu().double().T) + b.double())
    got = fg.ops.gemm_f16(a16, (w16, None), b.to(dev()), fg.ops.ACT_ELU, row_scale=rs.to(dev()))
    assert got.dtype == torch.float32
    assert float((got[rows, :N].cpu().double() - want).abs().max()) <= 2e-5 * float(want.abs().max())
    # against the two-plane product: the only difference is W's own rounding (2^-11 per element, averaging over K)
    two = fg.ops.gemm_f16(a16, fg.ops.split_f16(W.to(dev())), b.to(dev()), fg.ops.ACT_ELU, row_scale=rs.to(dev()))
    assert float((got - two).abs().max()) <= 2e-3 * float(two.abs().max())
    if N % 8 == 0:
        got16 = fg.ops.gemm_f16(a16, (w16, None), b.to(dev()), fg.ops.ACT_ELU, row_scale=rs.to(dev()), out_f16=True)
        assert got16.dtype == torch.float16 and torch.equal(got16, got[:, :N].half())
        old = fg._lib.set_tuning("gemm_pair_ws", 0)
        try:
            stream16 = fg.ops.gemm_f16(a16, (w16, None), b.to(dev()), fg.ops.ACT_ELU, row_scale=rs.to(dev()), out_f16=True)
            stream32 = fg.ops.gemm_f16(a16, (w16, None), b.to(dev()), fg.ops.ACT_ELU, row_scale=rs.to(dev()))
        finally:
            fg._lib.set_tuning("gemm_pair_ws", old)
        assert torch.equal(stream16, got16) and torch.equal(stream32, got)


def test_fp16_single_weight_plane_schedule_matches_oracle(fg):
    """PackedForward(precision='fp16') on the headline schedule at a size where the layer-2 transform runs as W-stationary CTA
    pairs (>= 148 * 256 rows): logits against the oracle inside the 1e-3 bound (measured ~2e-5 of max |log-prob|)."""
    n, F, C = 80000, 100, 47
    ei, part, cw, k = planted(fg, n, 2000000, seed=5)
    X = fg.synth.features(n, F, seed=5, device=dev())
    sd = fg.synth.init_state_dict(F, 512, C, seed=5)
    pack = fg.build_pack(ei, part, k, "none")
    f16 = fg.PackedForward(pack, sd, precision="fp16")
    assert f16.f16_hidden and f16.w_single and f16.apack is not None and f16.W[1][1] is None
    assert f16.f16_layer0 and f16.W0_f16[1] is None  # the first transform runs on fp16 planes too
    want = oracle_none(ei, X, part, k, sd, sub_ids=np.arange(min(k, 6000)))
    got = f16(X).cpu().numpy()
    assert_close(got[: want.shape[0]], want)
    assert np.abs(got[: want.shape[0]] - want).max() <= 2e-4 * max(1.0, np.abs(want).max())
    ref = fg.PackedForward(pack, sd, precision="fp16x2")(X).cpu().numpy()
    assert np.abs(got - ref).max() <= 2e-4 * np.abs(ref).max()
    # the two forms of the fused schedule: the next layer's aggregation in the previous transform's epilogue (default) / every
    # later layer as ONE conv kernel, aggregation on the raw accumulators (fitgnn_gcn_conv_aligned_f16, opt-in)
    assert not f16.conv_fused
    for prec in ("fp16", "fp16x2"):
        a = fg.PackedForward(pack, sd, precision=prec, conv_fused=True)
        b = fg.PackedForward(pack, sd, precision=prec, conv_fused=False)
        assert a.conv_fused and not b.conv_fused
        ga, gb = a(X).cpu().numpy(), b(X).cpu().numpy()
        assert_close(ga[: want.shape[0]], want)
        assert np.abs(ga - gb).max() <= 2e-4 * np.abs(gb).max()
    # layer 1 on bf16 hi/lo planes instead (what fp16x2 does): the difference is the rounding of the aggregated features
    import os
    os.environ["FITGNN_F16_LAYER0"] = "0"
    try:
        b0 = fg.PackedForward(pack, sd, precision="fp16")
    finally:
        del os.environ["FITGNN_F16_LAYER0"]
    assert not b0.f16_layer0
    assert np.abs(b0(X).cpu().numpy() - got).max() <= 2e-4 * np.abs(got).max()
    # classic schedule and three layers (the middle fused transform takes a single weight plane too)
    cl = fg.PackedForward(pack, sd, precision="fp16", fuse_aggregate=False)
    assert cl.f16_classic
    assert np.abs(cl(X).cpu().numpy() - ref).max() <= 2e-4 * np.abs(ref).max()
    sd3 = fg.synth.init_state_dict(F, 256, C, num_layers=3, seed=6)
    g3 = fg.PackedForward(pack, sd3, precision="fp16")(X).cpu().numpy()
    r3 = fg.PackedForward(pack, sd3, precision="bf16x3")(X).cpu().numpy()
    assert np.abs(g3 - r3).max() <= 3e-4 * np.abs(r3).max()


@pytest.mark.parametrize("width", [64, 128, 256, 512, 1024])
def test_spmm_f16_matches_the_fp32_spmm_of_the_same_plane(fg, width):
    """fitgnn_spmm_symnorm_f16 (fp16 plane in, fp32 sums, fp16 plane out) against fitgnn_spmm_symnorm on the same values in
    fp32: equal up to the final fp16 rounding (2^-11 relative); all rows / a row selection, bias + ELU, hub rows."""
    n = 20000
    ei, part, cw, k = planted(fg, n, 200000, seed=7, sizes="powerlaw")
    pack = fg.build_pack(ei, part, k, "none")
    g = torch.Generator(device="cuda").manual_seed(width)
    Xh = torch.randn(pack.n_rows, width, generator=g, device=dev()).half()
    bias = torch.randn(width, generator=g, device=dev())
    rows = torch.randperm(pack.n_rows, generator=g, device=dev())[: pack.n_rows // 3].sort().values.to(torch.int32)
    for out_rows in (None, rows):
        for b, act in ((None, fg.ops.ACT_NONE), (bias, fg.ops.ACT_ELU)):
            want = fg.ops.spmm_symnorm(pack.rowptr, pack.col, pack.dinv, Xh.float(), width, None, b, act, out_rows)
            got = fg.ops.spmm_symnorm_f16(pack.rowptr, pack.col, pack.dinv, Xh, width, None, b, act, out_rows)
            assert got.dtype == torch.float16 and got.shape == want.shape
            assert bool(((got.float() - want).abs() <= 4.9e-4 * want.abs() + 1e-7).all())
    # a star with 3000 leaves: the hub kernel (CTA-split row) on the fp16 plane
    m = 3001
    src = torch.cat([torch.zeros(m - 1, dtype=torch.long), torch.arange(1, m)])
    dst = torch.cat([torch.arange(1, m), torch.zeros(m - 1, dtype=torch.long)])
    rowptr, col, dinv = fg.ops.csr_from_coo(torch.stack([src, dst]).to(dev()), m)
    hubs = fg.ops.find_hubs(rowptr, None, m)
    assert hubs[1] == 1
    Xs = torch.randn(m, width, generator=g, device=dev()).half()
    want = fg.ops.spmm_symnorm(rowptr, col, dinv, Xs.float(), hubs=hubs)
    got = fg.ops.spmm_symnorm_f16(rowptr, col, dinv, Xs, hubs=hubs)
    assert bool(((got.float() - want).abs() <= 4.9e-4 * want.abs() + 1e-6).all())


@pytest.mark.parametrize("mode", ["none", "extra", "cluster"])
def test_classic_schedule_on_the_fp16_plane_matches_bf16x3_and_the_oracle(fg, mode):
    """precision='fp16x2' without the fused schedule (fuse_aggregate=False, or a pack that is not eligible): SpMM + transform
    per layer with the hidden state as ONE fp16 plane; power-law graph, so cluster mode has hub rows; row-mapped head."""
    n, F, H, C = 5000, 100, 512, 47
    ei = fg.synth.powerlaw_graph(n, 20000, seed=11)
    partition, comps, C_list = fg.synth.neighborhood_partition(ei, n, 0.3, seed=11)
    eid = torch.tensor(ei, device=dev())
    X = fg.synth.features(n, F, seed=11, device=dev())
    pack = fg.build_pack(eid, torch.tensor(partition.part), partition.k, mode)
    Xg = torch.cat([X, fg.coarsen.project(eid, X, partition)["Xc"]], 0) if mode == "cluster" else X
    for layers in (2, 3):
        sd = fg.synth.init_state_dict(F, H, C, num_layers=layers, seed=layers)
        f = fg.PackedForward(pack, sd, precision="fp16x2", fuse_aggregate=False)
        assert f.f16_classic
        got = f(Xg)
        ref = fg.PackedForward(pack, sd, precision="bf16x3", fuse_aggregate=False)(Xg)
        assert_close(got.cpu().numpy(), ref.cpu().numpy())
        assert float((got - ref).abs().max()) <= 3e-4 * float(ref.abs().max())
        if mode == "none" and layers == 2:
            ids = np.arange(min(partition.k, 300))
            subs = fo.subgraphs_from_partition(ei, X.cpu().numpy(), partition.part, ids)
            sel = [np.ones(s_["x"].shape[0], dtype=bool) for s_ in subs]
            want = fo.node_infer_batched({k_: v.cpu() for k_, v in sd.items()}, subs, sel, "node_cls", 128).numpy()
            assert_close(got[: want.shape[0]].cpu().numpy(), want)
            # the same rows through a row map into a caller's buffer
            perm = torch.randperm(f.n_out, device=dev()).to(torch.int32)
            out = torch.zeros(f.n_out, fg.ops.pad4(C), device=dev())
            fm = fg.PackedForward(pack, sd, precision="fp16x2", fuse_aggregate=False, out_map=perm)
            fm(Xg, out=out)
            assert torch.equal(out[perm.long(), :C], got[:, :C])
