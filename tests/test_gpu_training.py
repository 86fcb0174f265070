"""GPU: backward of the GCN operator (the training path) against torch-CPU autograd on the oracle's
restatement of PyG GCNConv; and a few optimiser steps of node_train_Gc / node_train_Gs_GD semantics."""
import argparse

import numpy as np
import pytest
import torch

from oracle import fitgnn_oracle as fo
from tests import golden_io as gio

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def fg():
    import fitgnn_b200
    return fitgnn_b200


def close(a, b, rtol=1e-3):
    a, b = a.detach().cpu().double().numpy(), b.detach().cpu().double().numpy()
    scale = max(1e-12, np.abs(b).max())
    assert np.abs(a - b).max() <= rtol * scale, (np.abs(a - b).max(), scale)


@pytest.mark.parametrize("fin,fout", [(12, 32), (100, 64), (300, 64)])  # aggregate-first and transform-first
def test_gcnconv_gradients_match_oracle(fg, fin, fout):
    d = gio.load("node_small")
    ref = gio.subgraphs(d, "extra_sub")
    x, ei = fo.collate(ref[:64])
    g = torch.Generator().manual_seed(fin)
    x = torch.rand(x.shape[0], fin, generator=g)
    conv = fg.GCNConv(fin, fout)
    with torch.no_grad():
        conv.bias.uniform_(-0.1, 0.1)
    w0, b0 = conv.lin.weight.detach().clone(), conv.bias.detach().clone()
    tgt = torch.rand(x.shape[0], fout, generator=g)
    # oracle: torch CPU autograd through gcn_conv_torch + ELU
    xo = x.clone().requires_grad_(True); wo = w0.clone().requires_grad_(True); bo = b0.clone().requires_grad_(True)
    lo = ((torch.nn.functional.elu(fo.gcn_conv_torch(xo, ei, wo, bo)) - tgt) ** 2).sum()
    lo.backward()
    conv = conv.to(DEV)
    xg = x.to(DEV).requires_grad_(True)
    out = conv(xg, ei.to(DEV), act=fg.ops.ACT_ELU)
    assert out.requires_grad
    loss = ((out - tgt.to(DEV)) ** 2).sum()
    loss.backward()
    close(loss, lo)
    close(conv.lin.weight.grad, wo.grad)
    close(conv.bias.grad, bo.grad)
    close(xg.grad, xo.grad)


def test_training_steps_follow_the_oracle(fg):
    """Three Adam steps of the reference's GD training semantics (one loss over all train nodes, run.py:199-204) with
    dropout disabled: parameters track a torch-CPU run of the oracle model."""
    d = gio.load("node_small")
    ref = gio.subgraphs(d, "extra_sub")
    x, ei = fo.collate(ref)
    y = torch.tensor(np.concatenate([r["y"] for r in ref])).long()
    tm = torch.tensor(np.concatenate([r["train_mask"] for r in ref]))
    args = argparse.Namespace(num_layers1=2, num_features=x.shape[1], hidden=32, num_classes=int(d["n_classes"]),
                              layer_name="GCNConv")
    sd = gio.state_dict(d)
    model = fg.Classify_node(args); model.load_state_dict(sd); model = model.to(DEV)
    opt = torch.optim.Adam(model.parameters(), lr=0.01, weight_decay=0.0005)  # main.py:193-194 defaults
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    opt_o = torch.optim.Adam(params.values(), lr=0.01, weight_decay=0.0005)
    model.eval()  # dropout off so the two runs are comparable; gradients still flow (as in the reference's eval loops)
    for _ in range(3):
        opt.zero_grad()
        out = model(x.to(DEV), ei.to(DEV))
        loss = torch.nn.functional.nll_loss(out[tm.to(DEV)], y.to(DEV)[tm.to(DEV)])
        loss.backward(); opt.step()
        opt_o.zero_grad()
        lo = torch.nn.functional.nll_loss(fo.classify_node(params, x, ei)[tm], y[tm])
        lo.backward(); opt_o.step()
        close(loss, lo)
    for k, v in model.state_dict().items():
        close(v, params[k], rtol=2e-3)
    # train mode: dropout is active, output differs from eval and is stochastic
    model.train()
    a = model(x.to(DEV), ei.to(DEV)); b = model(x.to(DEV), ei.to(DEV))
    assert not torch.equal(a, b)


def test_packed_train_step_and_pack_roundtrip(fg, tmp_path):
    """train_step_Gs on a pack (all subgraphs in one forward/backward) lowers the loss and equals the collated-batch loss;
    Pack.save / Pack.load round-trips bit-exactly."""
    d = gio.load("node_small")
    comps = gio.components(d, "extra")
    cos = gio.coarsenings_for_oracle(d, "extra", comps)
    partition = fg.coarsen.partition_from_components(comps, [c["C"] if c else None for c in cos], int(d["n"]))
    pack = fg.build_pack(torch.tensor(d["edge_index"], device=DEV), torch.tensor(partition.part), partition.k, "extra")
    path = tmp_path / "pack.pt"
    pack.save(path)
    again = fg.Pack.load(path, DEV)
    for name in fg.Pack._ARRAYS:
        assert torch.equal(getattr(pack, name), getattr(again, name)), name
    assert (again.n_rows, again.nnz, again.mode) == (pack.n_rows, pack.nnz, pack.mode)
    args = argparse.Namespace(num_layers1=2, num_features=d["x"].shape[1], hidden=32, num_classes=int(d["n_classes"]),
                              layer_name="GCNConv")
    model = fg.Classify_node(args); model.load_state_dict(gio.state_dict(d)); model = model.to(DEV)
    X, y, tm = torch.tensor(d["x"], device=DEV), torch.tensor(d["y"]), torch.tensor(d["train_mask"])
    # loss of the packed forward == loss of the reference-style collated forward (eval mode: no dropout)
    model.eval()
    ref = gio.subgraphs(d, "extra_sub")
    xc, eic = fo.collate(ref)
    yc = torch.tensor(np.concatenate([r["y"] for r in ref])).long().to(DEV)
    tmc = torch.tensor(np.concatenate([r["train_mask"] for r in ref])).to(DEV)
    l_ref = torch.nn.functional.nll_loss(model(xc.to(DEV), eic.to(DEV))[tmc], yc[tmc])
    out = fg.train.forward_on_pack(model, pack, X)
    rows = pack.split_masks(tm)
    l_pack = torch.nn.functional.nll_loss(out[rows], y.to(DEV)[pack.gid.long()][rows])
    close(l_pack, l_ref, rtol=1e-5)
    opt = torch.optim.Adam(model.parameters(), lr=0.01, weight_decay=0.0005)
    torch.manual_seed(0)
    losses = [fg.train.train_step_Gs(model, again, X, y, tm, opt) for _ in range(15)]
    assert losses[-1] < losses[0]


@pytest.mark.parametrize("task", ["graph_reg", "graph_cls"])
def test_graph_level_models_train_every_parameter(fg, task):
    """Regress/Classify_graph_gs and _gc under grad: the pooled tensor keeps its grad_fn (segment pooling is an autograd
    Function), so loss.backward() reaches the conv weights — gradients against torch-CPU autograd through the oracle's
    graph_gs_forward / graph_gc_forward (network.py:118-135, :189-204, :87-95, :158-166)."""
    from oracle.ref_shims import Data
    d = gio.load("graph_small")
    n_g = int(d["n_kept"])
    C = 1 if task == "graph_reg" else 3
    sd = fo.init_state_dict(1, 16, C, seed=9)
    args = argparse.Namespace(num_layers1=2, num_features=1, hidden=16, num_classes=C, layer_name="GCNConv")
    set_gs = [[Data(x=torch.tensor(s["x"]), edge_index=torch.tensor(s["edge_index"]), mask=torch.tensor(s["mask"]))
               for s in gio.subgraphs(d, f"g{g}_sub")] for g in range(n_g)]
    bt = torch.tensor(d["batch_tensor"])
    g_ = torch.Generator().manual_seed(0)
    tgt = torch.rand(n_g, C, generator=g_)
    # --- *_gs
    Model = fg.Regress_graph_gs if task == "graph_reg" else fg.Classify_graph_gs
    model = Model(args); model.load_state_dict(sd); model = model.to(DEV).eval()  # eval: dropout off, gradients on
    pred = model(set_gs, bt)
    assert pred.requires_grad
    loss = ((pred - tgt.to(DEV)) ** 2).sum()
    loss.backward()
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    want = fo.graph_gs_forward(params, [[dict(x=g.x, edge_index=g.edge_index, mask=g.mask) for g in gs] for gs in set_gs], bt, task)
    lo = ((want - tgt) ** 2).sum()
    lo.backward()
    close(loss, lo)
    for k, v in model.named_parameters():
        assert v.grad is not None and float(v.grad.abs().max()) > 0, k
        close(v.grad, params[k].grad, rtol=2e-3)
    # --- *_gc
    gx = torch.tensor(np.concatenate([d[f"g{g}_gc_x"] for g in range(n_g)])).float()
    off, eis = 0, []
    for g in range(n_g):
        eis.append(d[f"g{g}_gc_edge"] + off); off += d[f"g{g}_gc_x"].shape[0]
    ei = torch.tensor(np.concatenate(eis, 1))
    batch = torch.tensor(d["gc_batch"])
    ModelC = fg.Regress_graph_gc if task == "graph_reg" else fg.Classify_graph_gc
    mc = ModelC(args); mc.load_state_dict(sd); mc = mc.to(DEV).eval()
    pred = mc(Data(x=gx.to(DEV), edge_index=ei.to(DEV), batch=batch.to(DEV)))
    loss = ((pred - tgt.to(DEV)) ** 2).sum()
    loss.backward()
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    lo = ((fo.graph_gc_forward(params, gx, ei, batch, task) - tgt) ** 2).sum()
    lo.backward()
    close(loss, lo)
    for k, v in mc.named_parameters():
        assert v.grad is not None, k
        close(v.grad, params[k].grad, rtol=2e-3)


# ------------------------------------------------------------------------------------------ training kernels (tensor cores)
@pytest.mark.parametrize("R,out,inn", [(5000, 512, 104), (100003, 512, 512), (300, 47, 512), (4096, 32, 7), (70001, 512, 100)])
def test_gemm_tn_matches_fp64(fg, R, out, inn):
    """dW = Gᵀ·A over the rows (fitgnn_gemm_tn: transposing bf16 hi/lo split + batched split-K tcgen05 GEMM + ordered
    reduction) against an fp64 product of the fp32 operands; deterministic (two runs bit-identical)."""
    g = torch.Generator().manual_seed(R)
    G = torch.randn(R, out, generator=g)
    A = torch.randn(R, inn, generator=g)
    want = G.double().t() @ A.double()
    got = fg.ops.gemm_tn(G.to(DEV), A.to(DEV))
    assert got.shape == (out, inn)
    assert float((got.cpu().double() - want).abs().max()) <= 1e-4 * float(want.abs().max())
    assert torch.equal(got, fg.ops.gemm_tn(G.to(DEV), A.to(DEV)))
    # strided operands (a column slice of a wider matrix)
    Gw = torch.randn(R, out + 8, generator=g).to(DEV)
    got2 = fg.ops.gemm_tn(Gw[:, :out], A.to(DEV))
    want2 = Gw[:, :out].cpu().double().t() @ A.double()
    assert float((got2.cpu().double() - want2).abs().max()) <= 1e-4 * float(want2.abs().max())


def test_philox_dropout_and_its_backward(fg):
    rows, cols, p = 3001, 130, 0.5
    x = torch.rand(rows, cols, device=DEV) + 0.5
    y1 = fg.ops.dropout(x, p, seed=1234)
    y2 = fg.ops.dropout(x, p, seed=1234)
    y3 = fg.ops.dropout(x, p, seed=1235)
    assert torch.equal(y1, y2) and not torch.equal(y1, y3)
    keep = y1 != 0
    assert abs(float(keep.float().mean()) - (1 - p)) < 0.01
    assert torch.allclose(y1[keep], x[keep] / (1 - p))
    # every column / row sees both outcomes (no stripes)
    assert float(keep.float().mean(0).min()) > 0.4 and float(keep.float().mean(1).max()) < 0.7
    # backward: the same (seed) regenerates the mask; ELU' from the activation output
    h = torch.randn(rows, cols, device=DEV)
    h = torch.where(h > 0, h, torch.expm1(h))  # an ELU output
    g = torch.randn(rows, cols, device=DEV)
    gz = fg.ops.elu_dropout_backward(g, h, fg.ops.ACT_ELU, p, seed=1234)
    want = g * keep / (1 - p) * torch.where(h > 0, torch.ones_like(h), h + 1)
    assert torch.allclose(gz, want, rtol=1e-6, atol=1e-7)
    assert torch.allclose(fg.ops.elu_dropout_backward(g, h, fg.ops.ACT_NONE, 0.0), g)
    # p = 0.3: keep rate
    assert abs(float((fg.ops.dropout(x, 0.3, seed=7) != 0).float().mean()) - 0.7) < 0.01


def test_fused_adam_matches_torch_adam(fg):
    torch.manual_seed(0)
    m1 = torch.nn.Sequential(torch.nn.Linear(40, 64), torch.nn.ELU(), torch.nn.Linear(64, 5)).to(DEV)
    m2 = torch.nn.Sequential(torch.nn.Linear(40, 64), torch.nn.ELU(), torch.nn.Linear(64, 5)).to(DEV)
    m2.load_state_dict(m1.state_dict())
    o1 = torch.optim.Adam(m1.parameters(), lr=0.01, weight_decay=0.0005)
    o2 = fg.train.FusedAdam(m2.parameters(), lr=0.01, weight_decay=0.0005)
    x = torch.randn(200, 40, device=DEV)
    y = torch.randint(0, 5, (200,), device=DEV)
    for _ in range(6):
        for m, o in ((m1, o1), (m2, o2)):
            o.zero_grad()
            torch.nn.functional.cross_entropy(m(x), y).backward()
            o.step()
    for a, b in zip(m1.parameters(), m2.parameters()):
        assert torch.allclose(a, b, rtol=1e-5, atol=1e-6)
    assert all(p.data_ptr() >= o2.flat.data_ptr() for p in m2.parameters())  # parameters live in the flat buffer


def test_train_mode_dropout_gradients_match_oracle(fg):
    """conv -> ELU -> dropout(0.5) as one operator in train mode (network.py:31-33): forward and gradients against torch-CPU
    autograd on the oracle with the SAME mask (regenerated from the seed the operator drew from torch's generator)."""
    d = gio.load("node_small")
    ref = gio.subgraphs(d, "extra_sub")
    x, ei = fo.collate(ref[:64])
    fin, fout = 24, 64
    g = torch.Generator().manual_seed(3)
    x = torch.rand(x.shape[0], fin, generator=g)
    conv = fg.GCNConv(fin, fout)
    with torch.no_grad():
        conv.bias.uniform_(-0.1, 0.1)
    w0, b0 = conv.lin.weight.detach().clone(), conv.bias.detach().clone()
    tgt = torch.rand(x.shape[0], fout, generator=g)
    torch.manual_seed(11)
    seed = int(torch.randint(0, 2 ** 62, (1,)).item())
    mask = (fg.ops.dropout(torch.ones(x.shape[0], fout, device=DEV), 0.5, seed) != 0).cpu()
    xo = x.clone().requires_grad_(True); wo = w0.clone().requires_grad_(True); bo = b0.clone().requires_grad_(True)
    yo = torch.nn.functional.elu(fo.gcn_conv_torch(xo, ei, wo, bo)) * mask * 2.0
    lo = ((yo - tgt) ** 2).sum()
    lo.backward()
    conv = conv.to(DEV)
    xg = x.to(DEV).requires_grad_(True)
    torch.manual_seed(11)
    out = conv(xg, ei.to(DEV), act=fg.ops.ACT_ELU, dropout_p=0.5)
    close(out, yo)
    loss = ((out - tgt.to(DEV)) ** 2).sum()
    loss.backward()
    close(loss, lo)
    close(conv.lin.weight.grad, wo.grad)
    close(conv.bias.grad, bo.grad)
    close(xg.grad, xo.grad)


def test_packed_train_step_with_fused_adam(fg):
    """train_step_Gs with FusedAdam (flat parameter / gradient buffers, one optimiser kernel): same losses as torch.optim.Adam
    on the same model in eval-mode-equivalent conditions (dropout off), and the loss goes down in train mode."""
    d = gio.load("node_small")
    comps = gio.components(d, "extra")
    cos = gio.coarsenings_for_oracle(d, "extra", comps)
    partition = fg.coarsen.partition_from_components(comps, [c["C"] if c else None for c in cos], int(d["n"]))
    pack = fg.build_pack(torch.tensor(d["edge_index"], device=DEV), torch.tensor(partition.part), partition.k, "extra")
    args = argparse.Namespace(num_layers1=2, num_features=d["x"].shape[1], hidden=32, num_classes=int(d["n_classes"]),
                              layer_name="GCNConv")
    X, y, tm = torch.tensor(d["x"], device=DEV), torch.tensor(d["y"]), torch.tensor(d["train_mask"])
    model = fg.Classify_node(args); model.load_state_dict(gio.state_dict(d)); model = model.to(DEV)
    opt = fg.train.FusedAdam(model.parameters(), lr=0.01, weight_decay=0.0005)
    torch.manual_seed(0)
    losses = [fg.train.train_step_Gs(model, pack, X, y, tm, opt) for _ in range(15)]
    assert all(np.isfinite(losses)) and losses[-1] < losses[0]
