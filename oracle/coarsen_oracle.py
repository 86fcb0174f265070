"""CPU restatement of the reference's multilevel local-variation coarsening — TEST INFRASTRUCTURE, not product code (only
tests/ may import it).  Follows /root/reference/graph_coarsening/coarsening_utils.py:

    coarsen                      :18-182   (method 'variation_neighborhoods', the reference's default, utils.py:159)
    contract_variation_linear    :530-650  (candidate family = closed neighbourhoods :583-588, cost :554-560, the sequential
                                            contraction over a SortedList :606-648)
    get_coarsening_matrix        :212-254, coarsen_matrix :201-205, graph_utils.zero_diag :79-87

plain numpy / scipy, same operations in the same order, so that the costs — and with them the contraction order — come out
bit-identical.  Pinned by tests/golden/coarsen_algo.npz (tests/golden/make_golden_coarsen.py runs the unmodified reference
with (Uk, lk) passed through its own arguments; its internal eigsh starts from a random vector and is not reproducible).
Graphs are scipy CSR weight matrices (symmetric, zero diagonal) — what pygsp's Graph holds as W."""
import heapq

import numpy as np
import scipy.sparse as sp


def spectral_matrix(Uk, lk, K):
    """B = Uk[:, :K] diag(lk^-1/2) with (near-)zero eigenvalues dropped, coarsening_utils.py:78-83 / :90-95."""
    lk = np.array(lk, dtype=np.float64)
    mask = lk < 1e-10
    lk[mask] = 1
    lsinv = lk ** (-0.5)
    lsinv[mask] = 0
    return Uk[:, :K] @ np.diag(lsinv[:K])


def subgraph_cost(W, deg, A, nodes):
    """:554-560 — ||B^T L B||_F / (nc - 1), L = diag(2 deg - W_S 1) - W_S on the induced subgraph, B = centred rows of A."""
    nc = len(nodes)
    ones = np.ones(nc)
    Ws = W[nodes, :][:, nodes]
    L = np.diag(2 * deg[nodes] - Ws.dot(ones)) - Ws  # np.matrix, as in the reference (dense minus sparse)
    B = (np.eye(nc) - np.outer(ones, ones) / nc) @ A[nodes, :]
    return np.linalg.norm(B.T @ L @ B) / (nc - 1)


def contract_variation_neighborhoods(W, A, r):
    """:530-650 with mode 'neighborhood'.  W: CSR weights of the current level; returns the list of contraction sets.
    The reference keeps its candidates in a SortedList ordered by cost only: equal costs keep insertion order and pop(0) takes
    the oldest — a heap keyed (cost, insertion number) pops in exactly that order."""
    N = W.shape[0]
    deg = np.ravel(W.sum(axis=0))
    Wb = ((W > 0) + sp.eye(N, dtype=bool, format="csr")).tocsr()
    Wb.sort_indices()
    Wl = W.tolil()  # the reference slices a lil matrix (:539); values and summation order equal the CSR's
    heap, seq = [], 0
    for i in range(N):
        s = Wb.indices[Wb.indptr[i]: Wb.indptr[i + 1]].copy()
        heap.append((subgraph_cost(Wl, deg, A, s), seq, s))
        seq += 1
    heapq.heapify(heap)
    marked = np.zeros(N, dtype=bool)
    out = []
    n_reduce = np.floor(r * N)
    while heap:
        cost, _, s = heapq.heappop(heap)
        m = marked[s]
        if not m.any():
            gain = len(s) - 1
            if gain > n_reduce:
                continue
            marked[s] = True
            out.append(s)
            n_reduce -= gain
            if n_reduce <= 0:
                break
        else:
            s = s[~m]
            if len(s) > 1:
                heapq.heappush(heap, (subgraph_cost(Wl, deg, A, s), seq, s))
                seq += 1
    return out


def coarsening_matrix(N, sets):
    """:212-254 — row of a contracted set = its first (smallest) member, entries 1/sqrt(|set|); other rows identity."""
    C = sp.eye(N, format="lil")
    drop = []
    for s in sets:
        C[s[0], s] = 1 / np.sqrt(len(s))
        drop.extend(s[1:])
    keep = np.setdiff1d(np.arange(N), np.array(drop, dtype=np.int64))
    return sp.csc_matrix(C.tocsr()[keep, :])


def coarsen_weights(W, iC):
    """:201-205 + zero_diag + the symmetrisation of :139."""
    D = sp.diags(np.array(1 / np.sum(iC, 0))[0])
    Pinv = (iC.dot(D)).T
    Wc = sp.lil_matrix((Pinv.T).dot(W.dot(Pinv)))
    Wc.setdiag(0)
    Wc = sp.csr_matrix(Wc)
    Wc.eliminate_zeros()
    return sp.csr_matrix((Wc + Wc.T) / 2)


def coarsen(W, Uk, lk, K=10, r=0.5, max_levels=10, max_level_r=0.99):
    """:18-182.  Returns (C csc [n_c, N], Wc csr, number of levels)."""
    r = np.clip(r, 0, 0.999)
    N = W.shape[0]
    n, n_target = N, np.ceil((1 - r) * N)
    C = sp.eye(N, format="csc")
    Wc = sp.csr_matrix(W)
    levels = 0
    B = iC = None
    for level in range(1, max_levels + 1):
        Wl = Wc
        r_cur = np.clip(1 - n_target / n, 0.0, max_level_r)
        if level == 1:
            B = spectral_matrix(Uk, lk, K)
            A = B
        else:
            B = iC.dot(B)
            L = (sp.diags(np.ravel(Wl.sum(axis=0)), 0) - Wl).tocsc()
            d, V = np.linalg.eig(B.T @ L.dot(B))
            mask = d == 0
            d[mask] = 1
            dinvsqrt = d ** (-1 / 2)
            dinvsqrt[mask] = 0
            A = B @ np.diag(dinvsqrt) @ V
        sets = contract_variation_neighborhoods(Wl, A, r_cur)
        iC = coarsening_matrix(Wl.shape[0], sets)
        levels += 1
        if iC.shape[1] - iC.shape[0] <= 2:
            break
        C = iC.dot(C)
        Wc = coarsen_weights(Wl, iC)
        n = Wc.shape[0]
        if n <= n_target:
            break
    return sp.csc_matrix(C), Wc, levels
