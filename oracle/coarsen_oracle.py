"""CPU restatement of the reference's multilevel local-variation coarsening — TEST INFRASTRUCTURE, not product code (only
tests/ may import it).  Follows /root/reference/graph_coarsening/coarsening_utils.py:

    coarsen                      :18-182   (methods 'variation_neighborhoods' — the reference's default, utils.py:159 —
                                            'variation_cliques' and 'variation_edges')
    contract_variation_linear    :530-650  (candidate family = closed neighbourhoods :583-588, cost :554-560, the sequential
                                            contraction over a SortedList :606-648)
    contract_variation_edges     :483-527  (edge costs)
    matching_greedy              :931-989  (one pass over the edges in np.argsort(-weights) order, numpy's default sort)
('heavy_edge', :689-696, is NOT restated: with the scipy / numpy of this image `np.max(G.W, 0)` on pygsp's lil matrix returns the
matrix itself, so `wmax` becomes row 0 of W + 1e-5 instead of the column maxima — an accident of the installed versions, nothing
to pin parity on.)
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


def contract_variation_neighborhoods(W, A, r, mode="neighborhood"):
    """:530-650 with mode 'neighborhood' (family = closed neighbourhoods, :583-588) or 'cliques' (family = the maximal cliques
    networkx enumerates, node order inside a clique as it yields them, :590-595).  W: CSR weights of the current level;
    returns the list of contraction sets.
    The reference keeps its candidates in a SortedList ordered by cost only: equal costs keep insertion order and pop(0) takes
    the oldest — a heap keyed (cost, insertion number) pops in exactly that order."""
    N = W.shape[0]
    deg = np.ravel(W.sum(axis=0))
    Wb = ((W > 0) + sp.eye(N, dtype=bool, format="csr")).tocsr()
    Wb.sort_indices()
    Wl = W.tolil()  # the reference slices a lil matrix (:539); values and summation order equal the CSR's
    heap, seq = [], 0
    if mode == "cliques":
        import networkx as nx
        family = [np.array(c) for c in nx.find_cliques(nx.from_scipy_sparse_array(Wl))]
    else:
        family = [Wb.indices[Wb.indptr[i]: Wb.indptr[i + 1]].copy() for i in range(N)]
    for s in family:
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


def edge_list(W):
    """pygsp Graph.get_edge_list: the lower triangle's entries in row-major order (v_in > v_out) and their weights."""
    T = sp.tril(sp.csr_matrix(W)).tocoo()
    return T.row, T.col, T.data


def variation_edge_costs(W, A):
    """:483-513 — cost of contracting one edge: ||B^T L B||_F with the 2 x 2 Laplacian of the pair and B = centred rows of A."""
    deg = np.ravel(W.sum(axis=0))
    vi, vo, ww = edge_list(W)
    ones = np.ones(2)
    Pibot = np.eye(2) - np.outer(ones, ones) / 2
    out = np.empty(len(vi))
    for e in range(len(vi)):
        edge, w = np.array([vi[e], vo[e]]).astype(np.int32), ww[e]
        deg_new = 2 * deg[edge] - w
        L = np.array([[deg_new[0], -w], [-w, deg_new[1]]])
        B = Pibot @ A[edge, :]
        out[e] = np.linalg.norm(B.T @ L @ B)
    return out


def matching_greedy(W, weights, r):
    """:931-989 — heaviest edge first (np.argsort(-weights), default sort), skip edges with a matched endpoint, stop once
    n <= (1 - r) N."""
    N = W.shape[0]
    vi, vo, _ = edge_list(W)
    idx = np.argsort(-weights)
    marked = np.zeros(N, dtype=bool)
    n, n_target = N, (1 - r) * N
    out = []
    for e in idx:
        i, j = vi[e], vo[e]
        if marked[i] or marked[j]:
            continue
        marked[[i, j]] = True
        n -= 1
        out.append(np.array([i, j]))
        if n <= n_target:
            break
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


def coarsen(W, Uk, lk, K=10, r=0.5, max_levels=10, max_level_r=0.99, method="variation_neighborhoods"):
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
        assert method in ("variation_neighborhoods", "variation_cliques", "variation_edges")
        if True:
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
            if method == "variation_edges":  # :105-108, :523-525 (algorithm 'greedy': matching_greedy on -(-cost))
                sets = matching_greedy(Wl, -variation_edge_costs(Wl, A), r_cur)
            else:
                sets = contract_variation_neighborhoods(Wl, A, r_cur, "cliques" if method == "variation_cliques" else "neighborhood")
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
