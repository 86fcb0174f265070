"""Training steps on the pack: the reference's GD semantics — one forward over ALL subgraphs, one loss
over all selected nodes, one backward, one optimiser step (node_train_Gs_GD run.py:177-215; node_train_Gc run.py:26-37
is the same call on the coarsened graph) — with the conv forward/backward on libfitgnn_b200 (autograd.GCNConvFn), and
the gradient all-reduce the multi-GPU path needs (one flat buffer per step)."""
from __future__ import annotations

import torch

from .autograd import CsrPair
from .pack import Pack


def pack_csr(pack: Pack) -> CsrPair:
    return CsrPair(pack.rowptr, pack.col, pack.dinv)


def forward_on_pack(model, pack: Pack, X, csr: CsrPair | None = None):
    """Differentiable forward of a node model (fitgnn_b200.nn.Classify_node / Regress_node) over every row of the pack;
    X is the de-duplicated feature table ([N (+k), F]).  Returns [n_rows, C] with grad_fn."""
    from . import ops
    from .autograd import gcn_conv
    csr = csr or pack_csr(pack)
    x = X[pack.gid.long()]
    for i in range(model.num_layers):
        x = gcn_conv(x, model.conv[i].lin.weight, model.conv[i].bias, csr, ops.ACT_ELU)
        if model.training:
            x = torch.nn.functional.dropout(x, training=True)  # network.py:33 (p = 0.5)
    y = model.lt1(x)
    return torch.nn.functional.log_softmax(y, dim=1) if model._head == "log_softmax" else y


def allreduce_gradients(model, world_size: int, group=None):
    """One all-reduce (average) of the flat gradient buffer per optimiser step."""
    if world_size <= 1:
        return
    import torch.distributed as dist
    grads = [p.grad for p in model.parameters() if p.grad is not None]
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, group=group)
    flat /= world_size
    off = 0
    for g in grads:
        g.copy_(flat[off: off + g.numel()].view_as(g))
        off += g.numel()


def train_step_Gs(model, pack: Pack, X, y, train_mask, optimizer, loss_fn=None, world_size=1, csr=None):
    """node_train_Gs_GD (run.py:177-215): loss over every train node of the (rank-local) pack, backward, step.
    y / train_mask are global ([N]); extra / cluster rows never contribute (Pack.split_masks)."""
    model.train()
    optimizer.zero_grad()
    out = forward_on_pack(model, pack, X, csr)
    rows = pack.split_masks(train_mask)
    tgt = y.to(out.device)[pack.gid.long().clamp(max=y.numel() - 1)]
    loss_fn = loss_fn or torch.nn.functional.nll_loss
    loss = loss_fn(out[rows], tgt[rows])
    loss.backward()
    allreduce_gradients(model, world_size)
    optimizer.step()
    return float(loss.detach())
