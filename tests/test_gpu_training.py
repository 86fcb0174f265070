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
