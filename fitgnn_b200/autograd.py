"""Backward of the GCN operator on the same kernels (the first "next" row of the scope table).

The reference trains through PyG autograd (node_train_Gc run.py:26-37, node_train_Gs_GD run.py:177-215).  Here the
conv (+ fused ELU) is a torch.autograd.Function whose forward AND backward run on libfitgnn_b200:

    out = act(Â · (x Wᵀ) + b)
    gz  = g ⊙ act'(out)                      ELU'(z) = 1 (z > 0) else out + 1
    db  = Σ_rows gz
    aggregate-first (in <= out):  A = Â x (saved);  dW = gzᵀ A;  dx = Âᵀ (gz W)
    transform-first (in >  out):  dZ = Âᵀ gz;       dW = dZᵀ x;  dx = dZ W

Âᵀ is applied with the CSR of the reversed edges and the ORIGINAL deg^-1/2 vector (for the undirected graphs of every
reference dataset Âᵀ = Â).  GEMMs use the exact-fp32 kernel; transposed operands are materialised with torch (index
plumbing) — the tensor-core TN kernel is future work, training throughput is not the round-1 target.
"""
from __future__ import annotations

import torch

from . import ops


class CsrPair:
    """Forward CSR (rows = targets) + CSR of the reversed graph (rows = sources), both with the forward dinv."""

    def __init__(self, rowptr, col, dinv, rowptr_t=None, col_t=None):
        self.rowptr, self.col, self.dinv = rowptr, col, dinv
        self.rowptr_t, self.col_t = rowptr_t, col_t

    @staticmethod
    def from_edge_index(edge_index, n):
        rowptr, col, dinv = ops.csr_from_coo(edge_index, n)
        return CsrPair(rowptr, col, dinv)

    def transposed(self):
        if self.rowptr_t is None:
            n = self.rowptr.numel() - 1
            deg = (self.rowptr[1:] - self.rowptr[:-1]).long()
            tgt = torch.repeat_interleave(torch.arange(n, device=self.col.device), deg)
            rev = torch.stack([tgt, self.col.long()])  # reversed edges: source := old target
            self.rowptr_t, self.col_t, _ = ops.csr_from_coo(rev.contiguous(), n)
        return self.rowptr_t, self.col_t


def _pad4(t):
    f = t.shape[1]
    if f % 4 == 0:
        return t.contiguous()
    out = torch.zeros(t.shape[0], ops.pad4(f), dtype=t.dtype, device=t.device)
    out[:, :f] = t
    return out


class GCNConvFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, csr: CsrPair, act: int):
        xin = _pad4(x.detach().float())
        w = _pad4(weight.detach())
        b = bias.detach().contiguous() if bias is not None else None
        fin, fout = weight.shape[1], weight.shape[0]
        ctx.csr, ctx.act, ctx.fin = csr, act, fin
        ctx.transform_first = fin > fout
        if ctx.transform_first:
            z = ops.gemm_bias_act(xin, w, None, ops.ACT_NONE, K=xin.shape[1])
            out = ops.spmm_symnorm(csr.rowptr, csr.col, csr.dinv, z, bias=b, act=act)
            saved_in = xin
        else:
            a = ops.spmm_symnorm(csr.rowptr, csr.col, csr.dinv, xin)
            out = ops.gemm_bias_act(a, w, b, act, K=xin.shape[1])
            saved_in = a
        ctx.save_for_backward(saved_in, w, out)
        ctx.has_bias = bias is not None
        return out

    @staticmethod
    def backward(ctx, g):
        saved_in, w, out = ctx.saved_tensors
        csr = ctx.csr
        g = g.contiguous().float()
        gz = torch.where(out > 0, g, g * (out + 1.0)) if ctx.act == ops.ACT_ELU else g
        db = gz.sum(0) if ctx.has_bias else None
        rp_t, col_t = csr.transposed()
        fin = ctx.fin
        wt = w.t().contiguous()  # [in_p, out]
        if ctx.transform_first:
            dz = ops.spmm_symnorm(rp_t, col_t, csr.dinv, gz)                 # Âᵀ gz          [n, out]
            dw = ops.gemm_bias_act(dz.t().contiguous(), saved_in.t().contiguous())   # dZᵀ x  [out, in_p]
            dx = ops.gemm_bias_act(dz, wt) if ctx.needs_input_grad[0] else None      # dZ W   [n, in_p]
        else:
            dw = ops.gemm_bias_act(gz.t().contiguous(), saved_in.t().contiguous())   # gzᵀ A  [out, in_p]
            dx = None
            if ctx.needs_input_grad[0]:
                da = ops.gemm_bias_act(gz, wt)                                       # gz W   [n, in_p]
                dx = ops.spmm_symnorm(rp_t, col_t, csr.dinv, da)                     # Âᵀ dA
        if dx is not None:
            dx = dx[:, :fin]
        return dx, dw[:, :fin], db, None, None


def gcn_conv(x, weight, bias, csr: CsrPair, act: int = ops.ACT_NONE):
    return GCNConvFn.apply(x, weight, bias, csr, act)
