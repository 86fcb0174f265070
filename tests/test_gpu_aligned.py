"""GPU tests of the group-aligned pack layout and the aggregation fused into the tensor-core transform's epilogue
(fitgnn_pack_align_*, fitgnn_gcn_transform_aggregate, fitgnn_gemm_head_rows).  Integer structure bit-exact against
the oracle's restatement of the layout; fp32 results within 1e-3 relative (RTOL) of the fp64 oracle."""
import numpy as np
import pytest
import torch

from oracle import fitgnn_oracle as fo

pytestmark = pytest.mark.gpu
RTOL = 1e-3


def dev():
    return torch.device("cuda:0")


@pytest.fixture(scope="module")
def fg():
    import fitgnn_b200
    return fitgnn_b200


def small_subgraph_graph(n, k, seed, extra_intra=1.0, inter=3.0, max_size=None):
    """Random graph whose partition has k clusters of small random sizes; intra-cluster tree + extras (with duplicate
    edges), plus inter-cluster edges that the 'none' pack drops."""
    rng = np.random.default_rng(seed)
    part = np.sort(rng.integers(0, k, n))
    if max_size:
        # cap cluster sizes by re-assigning the overflow to fresh clusters
        out, nxt, cnt = part.copy(), k, {}
        for i, p in enumerate(part):
            c = cnt.get(p, 0)
            if c >= max_size:
                out[i] = nxt + (c - max_size) // max_size + 1000 * p
            cnt[p] = c + 1
        part = out
    _, part = np.unique(part, return_inverse=True)
    k = int(part.max()) + 1
    perm = rng.permutation(n)
    start = np.searchsorted(part, np.arange(k))
    size = np.bincount(part, minlength=k)
    pos = np.arange(n) - start[part]
    child = np.nonzero(pos > 0)[0]
    parent = start[part[child]] + (rng.random(child.size) * pos[child]).astype(np.int64)
    s, d = [child], [parent]
    ne = int(extra_intra * child.size)
    if ne:
        c2 = child[rng.integers(0, child.size, ne)]
        p2 = start[part[c2]] + (rng.random(ne) * size[part[c2]]).astype(np.int64)
        ok = p2 != c2
        s.append(c2[ok]); d.append(p2[ok])
    ni = int(inter * n)
    u, v = rng.integers(0, n, ni), rng.integers(0, n, ni)
    ok = part[u] != part[v]
    s.append(u[ok]); d.append(v[ok])
    s, d = np.concatenate(s), np.concatenate(d)
    a, b = perm[s], perm[d]
    ei = np.stack([np.concatenate([a, b]), np.concatenate([b, a])]).astype(np.int64)
    part_by_node = np.empty(n, dtype=np.int64)
    part_by_node[perm] = part
    # renumber clusters by smallest member (reference convention)
    first = np.full(k, n)
    np.minimum.at(first, part_by_node, np.arange(n))
    new_id = np.empty(k, dtype=np.int64)
    new_id[np.argsort(first)] = np.arange(k)
    return ei, new_id[part_by_node].astype(np.int32), k


def pack_arrays(p):
    return {k: getattr(p, k).cpu().numpy() for k in ("rowptr", "col", "dinv", "gid", "sub_ptr", "core_rows", "is_core", "mask")}


@pytest.mark.parametrize("policy", ["order", "degree"])
@pytest.mark.parametrize("n,k,seed,max_size", [(40, 20, 0, None), (5000, 2300, 1, None), (3000, 400, 2, 13), (64, 2, 3, 32),
                                               (33, 33, 4, None)])
def test_aligned_pack_matches_oracle_layout(fg, n, k, seed, max_size, policy):
    ei, part, k = small_subgraph_graph(n, k, seed, max_size=max_size)
    pack = fg.build_pack(torch.tensor(ei, device=dev()), torch.tensor(part), k, "none")
    want = fo.aligned_pack(pack_arrays(pack), 32, policy)
    ap = pack.aligned(32, policy)
    if want is None:
        assert ap is None
        return
    assert ap is not None and ap.n_rows == want["n_rows"]
    for name in ("rowptr", "col", "gid", "sub_ptr", "core_rows", "is_core", "mask", "orig_row", "new_of_old"):
        assert np.array_equal(getattr(ap, name).cpu().numpy().astype(np.int64), want[name].astype(np.int64)), name
    assert np.array_equal(ap.dinv.cpu().numpy(), want["dinv"])  # bit-exact copy
    assert np.array_equal(ap.agg_desc.cpu().numpy().view(np.uint64), want["agg_desc"])
    assert ap.agg_ok == want["agg_ok"]
    # the property the fused kernel relies on: every entry of a row lies in the row's own group of 32
    rp, col = want["rowptr"], want["col"]
    rows = np.repeat(np.arange(ap.n_rows), np.diff(rp))
    assert np.array_equal(rows // 32, col // 32)
    # every source row appears exactly once; the aligned CSR is the source CSR relabelled
    orig = want["orig_row"]
    assert np.array_equal(np.sort(orig[orig >= 0]), np.arange(pack.n_rows))
    if policy == "degree" and want["n_rows"] > 64:
        n_order = fo.aligned_layout(pack_arrays(pack)["sub_ptr"], 32, "order")[1]
        assert want["n_rows"] <= n_order  # gap filling never pads more than the in-order placement


def test_unalignable_pack_returns_none(fg):
    ei, part, k = small_subgraph_graph(400, 4, 5)  # ~100 rows per subgraph
    pack = fg.build_pack(torch.tensor(ei, device=dev()), torch.tensor(part), k, "none")
    assert pack.aligned(32) is None
    sd = fo.init_state_dict(16, 256, 5, seed=0)
    f = fg.PackedForward(pack, sd, precision="bf16x3")  # auto: falls back to SpMM + GEMM
    assert f.apack is None
    with pytest.raises(ValueError):
        fg.PackedForward(pack, sd, precision="bf16x3", fuse_aggregate=True)


@pytest.mark.parametrize("n,k,K,N,seed", [(300, 130, 104, 512, 0), (6000, 2700, 64, 256, 1), (9000, 4000, 512, 512, 2),
                                          (1000, 90, 104, 512, 3)])
def test_transform_aggregate_matches_fp64(fg, n, k, K, N, seed):
    ei, part, k = small_subgraph_graph(n, k, seed, max_size=30)
    pack = fg.build_pack(torch.tensor(ei, device=dev()), torch.tensor(part), k, "none")
    ap = pack.aligned(32)
    assert ap is not None
    g = torch.Generator().manual_seed(seed)
    M = ap.n_rows
    A = torch.randn(M, K, generator=g)
    W = torch.randn(N, K, generator=g) / K ** 0.5
    b = torch.randn(N, generator=g) * 0.1
    A_pl = fg.ops.split_bf16(A.to(dev()))
    W_pl = fg.ops.split_bf16(W.to(dev()))
    if not ap.agg_ok:
        pytest.skip("a row has more than 12 neighbours")
    for split in (True, False):
        out = fg.ops.gcn_transform_aggregate(A_pl, W_pl, b.to(dev()), fg.ops.ACT_ELU, ap.agg_desc, ap.dinv, split_out=split)
        got = (out[0].float() + out[1].float()) if split else out
        h = torch.nn.functional.elu(A.double() @ W.double().T + b.double()).numpy()
        a = fo.aligned_pack(pack_arrays(pack), 32, "degree")
        want = fo.aggregate_dense(a["rowptr"], a["col"], a["dinv"], h)
        got = got.cpu().numpy()
        real = a["orig_row"] >= 0
        scale = np.abs(want).max()
        err = np.abs(got[real] - want[real]).max() / scale
        assert err < 1e-4, err
        assert (got[~real] == 0).all()  # padding rows are written as zeros


@pytest.mark.parametrize("N,head", [(47, "log_softmax"), (7, "softmax"), (1, "identity"), (48, "identity")])
def test_head_rows_drops_padding(fg, N, head):
    g = torch.Generator().manual_seed(N)
    M, K = 1000, 512
    A = torch.randn(M, K, generator=g)
    W = torch.randn(N, K, generator=g) / K ** 0.5
    b = torch.randn(N, generator=g)
    keep = torch.rand(M, generator=g) > 0.1
    row_map = torch.full((M,), -1, dtype=torch.int32)
    n_keep = int(keep.sum())
    row_map[keep] = torch.randperm(n_keep, generator=g).to(torch.int32)
    ld = (N + 3) // 4 * 4
    out = torch.full((n_keep, ld), 7.0, device=dev())
    hd = {"identity": 0, "log_softmax": 1, "softmax": 2}[head]
    fg.ops.gemm_head_rows(fg.ops.split_bf16(A.to(dev())), fg.ops.split_bf16(W.to(dev())), b.to(dev()), fg.ops.ACT_NONE, hd,
                          row_map.to(dev()), out, K=K, N=N)
    z = A.double() @ W.double().T + b.double()
    if head == "log_softmax":
        z = torch.log_softmax(z, 1)
    elif head == "softmax":
        z = torch.softmax(z, 1)
    want = torch.empty(n_keep, N, dtype=torch.float64)
    want[row_map[keep].long()] = z[keep]
    got = out.cpu()[:, :N].double()
    assert (got - want).abs().max() / max(1.0, want.abs().max()) < 1e-4
    pad = out.cpu()[:, N:]
    assert ((pad == 7.0) | (pad == 0.0)).all()  # pitch padding: untouched or zero-filled, never garbage


@pytest.mark.parametrize("n,k,F,H,C,layers", [(3000, 1300, 100, 512, 47, 2), (2000, 900, 30, 256, 5, 3), (500, 400, 100, 512, 7, 2)])
def test_forward_fused_aggregation_matches_classic_and_oracle(fg, n, k, F, H, C, layers):
    ei, part, k = small_subgraph_graph(n, k, 11, max_size=12)
    pack = fg.build_pack(torch.tensor(ei, device=dev()), torch.tensor(part), k, "none")
    X = fg.synth.features(n, F, seed=1).to(dev())
    sd = fo.init_state_dict(F, H, C, num_layers=layers, seed=2)
    fused = fg.PackedForward(pack, sd, precision="bf16x3", fuse_aggregate=True)
    classic = fg.PackedForward(pack, sd, precision="bf16x3", fuse_aggregate=False)
    assert fused.apack is not None and classic.apack is None
    a, b = fused(X), classic(X)
    assert a.shape == b.shape == (pack.n_rows, C)
    assert (a - b).abs().max().item() < 1e-4 * max(1.0, b.abs().max().item())
    # oracle: per-subgraph fp32 torch-CPU restatement of network.py:29-35
    subs = fo.subgraphs_from_partition(ei, X.cpu().numpy(), part, np.arange(k))
    want = fo.node_infer_batched(sd, subs, [np.ones(s["x"].shape[0], dtype=bool) for s in subs], "node_cls", 128).numpy()
    got = a.cpu().numpy()
    assert np.abs(got - want).max() / max(1.0, np.abs(want).max()) < RTOL
    # writing into a caller-provided (padded-pitch) buffer
    buf = torch.zeros(pack.n_rows, (C + 3) // 4 * 4, device=dev())
    fused(X, out=buf)
    assert torch.equal(buf[:, :C], a)


@pytest.mark.parametrize("precision,fuse", [("bf16x3", True), ("bf16x3", False), ("fp32", False)])
@pytest.mark.parametrize("mode,F", [("none", 100), ("none", 30), ("extra", 50)])
def test_pack_ordered_features_equal_table_features(fg, mode, F, precision, fuse):
    """PackedForward.pack_features: one row per pack row (the reference's collated batch.x, utils.py:248 + run.py:336)
    must give bit-identical logits to the node-ordered table gathered through gid."""
    if fuse and mode != "none":
        pytest.skip("the fused aggregation needs every row to be an output row")
    n, k = 2500, 1100
    ei, part, k = small_subgraph_graph(n, k, 21, max_size=12)
    pack = fg.build_pack(torch.tensor(ei, device=dev()), torch.tensor(part), k, mode)
    X = fg.synth.features(n, F, seed=3).to(dev())
    sd = fo.init_state_dict(F, 256, 6, num_layers=2, seed=4)
    fwd = fg.PackedForward(pack, sd, precision=precision, fuse_aggregate=fuse)
    assert (fwd.apack is not None) == fuse
    Xp = fwd.pack_features(X)
    rows = (fwd.apack if fuse else pack).n_rows
    assert Xp.shape[0] == rows and Xp.shape[1] % 4 == 0 and Xp.is_contiguous()
    a = fwd(Xp, packed=True).clone()
    b = fwd(X)
    assert torch.equal(a, b)
    with pytest.raises(AssertionError):
        fwd(X[:-1], packed=False)


@pytest.mark.parametrize("split", [True, False])
@pytest.mark.parametrize("packed", [True, False])
@pytest.mark.parametrize("n,k,seed,max_size,width", [(40, 20, 0, None, 4), (5000, 2300, 1, None, 100), (3000, 400, 2, 13, 128),
                                                     (64, 2, 3, 32, 36), (33, 33, 4, None, 100), (20000, 9000, 5, 12, 100)])
def test_grouped_spmm_is_bit_identical_to_generic_spmm(fg, n, k, seed, max_size, width, packed, split):
    """fitgnn_spmm_symnorm_grouped (one warp per 32-row group, source rows staged in shared memory) against
    fitgnn_spmm_symnorm on the same group-aligned pack: same accumulation order -> identical bits; both feature layouts
    (pack-ordered rows / node table through gid), fp32 and bf16 hi/lo outputs, padding rows, a ragged last group."""
    ei, part, k = small_subgraph_graph(n, k, seed, max_size=max_size)
    pack = fg.build_pack(torch.tensor(ei, device=dev()), torch.tensor(part), k, "none")
    ap = pack.aligned(32, "degree")
    assert ap is not None
    g = torch.Generator(device="cuda").manual_seed(seed)
    X = torch.randn(n, width, generator=g, device=dev())
    src = None if packed else ap.gid
    Xin = X[ap.gid.long()].contiguous() if packed else X
    want = fg.ops.spmm_symnorm(ap.rowptr, ap.col, ap.dinv, Xin, width, src, split=split)
    got = fg.ops.spmm_symnorm_grouped(ap.rowptr, ap.col, ap.dinv, Xin, width, src, split=split)
    if split:
        assert torch.equal(got[0], want[0]) and torch.equal(got[1], want[1])
    else:
        assert torch.equal(got, want)
        # and against the fp64 dense restatement of D^-1/2 (A+I) D^-1/2 on the aligned CSR
        rp, col, dinv = ap.rowptr.cpu().numpy(), ap.col.cpu().numpy(), ap.dinv.cpu().numpy().astype(np.float64)
        Xn = Xin.cpu().numpy().astype(np.float64)
        if not packed:
            Xn = Xn[ap.gid.cpu().numpy()]
        ref = np.zeros((ap.n_rows, width))
        rows = np.repeat(np.arange(ap.n_rows), np.diff(rp))
        np.add.at(ref, rows, (dinv[rows] * dinv[col])[:, None] * Xn[col])
        assert np.abs(got.cpu().numpy() - ref).max() <= 1e-5 * max(1.0, np.abs(ref).max())


def test_grouped_spmm_rejects_wide_rows(fg):
    ei, part, k = small_subgraph_graph(200, 90, 0, max_size=10)
    ap = fg.build_pack(torch.tensor(ei, device=dev()), torch.tensor(part), k, "none").aligned(32, "order")
    X = torch.zeros(ap.n_rows, 132, device=dev())
    with pytest.raises(Exception, match="EUNSUP|width"):
        fg.ops.spmm_symnorm_grouped(ap.rowptr, ap.col, ap.dinv, X, 132, None)


def test_grouped_spmm_fills_the_pad_columns(fg):
    """pad_value: planes with pitch width + 4 get their pad columns written too (hi[:, width] = pad_value, rest 0), the
    data columns stay bit-identical; planes with another pitch are left alone."""
    ei, part, k = small_subgraph_graph(4000, 1800, 7, max_size=12)
    ap = fg.build_pack(torch.tensor(ei, device=dev()), torch.tensor(part), k, "none").aligned(32, "degree")
    width = 100
    X = torch.randn(ap.n_rows, width, device=dev())
    want = fg.ops.spmm_symnorm(ap.rowptr, ap.col, ap.dinv, X, width, None, split=True)
    for pitch in (width + 4, width + 12):
        hi = torch.full((ap.n_rows, pitch), 7.0, dtype=torch.bfloat16, device=dev())
        lo = torch.full((ap.n_rows, pitch), 7.0, dtype=torch.bfloat16, device=dev())
        fg.ops.spmm_symnorm_grouped(ap.rowptr, ap.col, ap.dinv, X, width, None, split=True, out=(hi, lo), pad_value=1.0)
        assert torch.equal(hi[:, :width], want[0]) and torch.equal(lo[:, :width], want[1])
        if pitch == width + 4:
            assert bool((hi[:, width] == 1.0).all()) and bool((hi[:, width + 1:] == 0).all()) and bool((lo[:, width:] == 0).all())
        else:
            assert bool((hi[:, width:] == 7.0).all()) and bool((lo[:, width:] == 7.0).all())


@pytest.mark.parametrize("packed", [True, False])
@pytest.mark.parametrize("n,k,seed,max_size,width", [(40, 20, 0, None, 4), (5000, 2300, 1, None, 100), (3000, 400, 2, 13, 128),
                                                     (33, 33, 4, None, 100), (20000, 9000, 5, 12, 100)])
def test_grouped_spmm_fp16_plane_is_the_rounded_fp32_result(fg, n, k, seed, max_size, width, packed):
    """fitgnn_spmm_symnorm_grouped_f16: the same fp32 sums as the fp32 kernel, rounded ONCE to fp16 (the layer-1 operand of
    precision='fp16'); pad columns written when the pitch is width + 4, left alone otherwise."""
    ei, part, k = small_subgraph_graph(n, k, seed, max_size=max_size)
    pack = fg.build_pack(torch.tensor(ei, device=dev()), torch.tensor(part), k, "none")
    ap = pack.aligned(32, "degree")
    g = torch.Generator(device="cuda").manual_seed(seed)
    X = torch.randn(n, width, generator=g, device=dev())
    src = None if packed else ap.gid
    Xin = X[ap.gid.long()].contiguous() if packed else X
    want = fg.ops.spmm_symnorm_grouped(ap.rowptr, ap.col, ap.dinv, Xin, width, src).half()
    got = fg.ops.spmm_symnorm_grouped_f16(ap.rowptr, ap.col, ap.dinv, Xin, width, src)
    assert got.dtype == torch.float16 and torch.equal(got, want)
    for pitch in (width + 4, width + 12):
        if width >= 128:
            continue
        out = torch.full((ap.n_rows, pitch), 7.0, dtype=torch.float16, device=dev())
        fg.ops.spmm_symnorm_grouped_f16(ap.rowptr, ap.col, ap.dinv, Xin, width, src, out=out, pad_value=1.0)
        assert torch.equal(out[:, :width], want)
        if pitch == width + 4:
            assert bool((out[:, width] == 1.0).all()) and bool((out[:, width + 1:] == 0).all())
        else:
            assert bool((out[:, width:] == 7.0).all())
