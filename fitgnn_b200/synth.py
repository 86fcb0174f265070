"""Seeded synthetic inputs of the shapes BASELINE.json names (there is no network for the real datasets).

The coarsening algorithm itself is out of scope (it produces the partition once, on the CPU); on the GPU box
/root/reference does not exist, so the partitions used by tests / bench come from
  * `neighborhood_partition` — a seeded greedy contraction of closed neighbourhoods, the same family of
    contraction sets as the reference's `variation_neighborhoods` (coarsening_utils.py:572-576), single level,
    C weights 1/sqrt(cluster size) (coarsening_utils.py:239), components handled like utils.py:144-166;
  * `planted_partition` — clusters known by construction, for the ogbn-products-shaped config 5 (the
    "community-detection proxy"), generated directly on the device.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp
import torch

from .coarsen import Partition, partition_from_components

# (nodes, undirected edges, features, classes, coarsening_ratio) — dataset_info.csv:5,7,4 and SURVEY §8d
SHAPES = {
    "cora": (2708, 5278, 1433, 7, 0.3),
    "pubmed": (19717, 44324, 500, 3, 0.5),
    "physics": (34493, 247962, 8415, 5, 0.1),
    "products": (2449029, 61859140, 100, 47, 0.5),
}


def powerlaw_graph(n, n_undirected, seed=0):
    """Connected-ish preferential-attachment graph with ~n_undirected edges; returns int64 [2, 2E] COO with both
    directions, unsorted."""
    rng = np.random.default_rng(seed)
    m = max(1, int(round(n_undirected / n)))
    src = np.repeat(np.arange(1, n), m)
    # attach to an earlier node, biased to low ids (heavy-tailed degrees)
    dst = (rng.random(src.size) ** 2 * src).astype(np.int64)
    extra = n_undirected - src.size
    if extra > 0:
        a = rng.integers(0, n, extra)
        b = (rng.random(extra) ** 2 * n).astype(np.int64)
        src, dst = np.concatenate([src, a]), np.concatenate([dst, b])
    keep = src != dst
    und = np.unique(np.stack([np.minimum(src[keep], dst[keep]), np.maximum(src[keep], dst[keep])], 1), axis=0)
    perm = rng.permutation(n)
    und = perm[und]
    ei = np.concatenate([und, und[:, ::-1]], 0)
    return np.ascontiguousarray(ei[rng.permutation(len(ei))].T)


def components_sorted(edge_index, n):
    """pygsp extract_components order (by smallest member) then stable size-descending sort (utils.py:144-146)."""
    A = sp.coo_matrix((np.ones(edge_index.shape[1], dtype=np.int8), (edge_index[0], edge_index[1])), shape=(n, n)).tocsr()
    ncomp, labels = sp.csgraph.connected_components(A, directed=False)
    order = np.argsort(labels, kind="stable")
    bounds = np.searchsorted(labels[order], np.arange(ncomp + 1))
    comps = [order[bounds[c]:bounds[c + 1]] for c in range(ncomp)]
    comps.sort(key=lambda c: c[0])
    return sorted(comps, key=len, reverse=True)


def neighborhood_partition(edge_index, n, ratio, seed=0, max_cluster=None):
    """Greedy independent set of closed neighbourhoods per component until ceil(ratio * n_comp) supernodes remain
    (coarsen() targets n_target = ceil((1 - r) N) with r = 1 - ratio, coarsening_utils.py:57-61, main.py:278).
    Returns (Partition, comps, C_list) with C_list[i] a scipy CSC [k_i, n_i] (None for single-node components)."""
    rng = np.random.default_rng(seed)
    A = sp.coo_matrix((np.ones(edge_index.shape[1], dtype=np.int8), (edge_index[0], edge_index[1])), shape=(n, n)).tocsr()
    A.setdiag(0)
    A.eliminate_zeros()
    indptr, indices = A.indptr, A.indices
    comps = components_sorted(edge_index, n)
    C_list = []
    for comp in comps:
        nc = len(comp)
        if nc == 1:
            C_list.append(None)
            continue
        local = {int(v): i for i, v in enumerate(comp)}
        n_reduce = nc - int(np.ceil(ratio * nc))
        marked = np.zeros(nc, dtype=bool)
        owner = np.arange(nc)
        deg = np.array([indptr[v + 1] - indptr[v] for v in comp])
        # low-degree centres first (small neighbourhoods cost least), ties broken by a seeded shuffle
        order = np.lexsort((rng.random(nc), deg))
        for i in order:
            if n_reduce <= 0:
                break
            if marked[i]:
                continue
            nb = [local[int(u)] for u in indices[indptr[comp[i]]:indptr[comp[i] + 1]]]
            members = [i] + [j for j in nb if not marked[j]]
            if max_cluster:
                members = members[:max_cluster]
            gain = len(members) - 1
            if gain == 0:
                continue
            if gain > n_reduce:
                members = members[: n_reduce + 1]
                gain = n_reduce
            marked[members] = True
            owner[members] = min(members)
            n_reduce -= gain
        reps, inv = np.unique(owner, return_inverse=True)  # supernodes ordered by smallest member
        sizes = np.bincount(inv)
        C = sp.csc_matrix((1.0 / np.sqrt(sizes[inv]), (inv, np.arange(nc))), shape=(len(reps), nc))
        C_list.append(C)
    return partition_from_components(comps, C_list, n), comps, C_list


def features(n, F, seed=0, kind="dense", device="cpu"):
    """'dense': U[0,1) rows, L1-normalised (--normalize_features, main.py:38); 'bow': ~18 ones per row, normalised."""
    g = torch.Generator(device=device).manual_seed(seed)
    if kind == "bow":
        x = torch.zeros(n, F, device=device)
        idx = torch.randint(0, F, (n, 18), generator=g, device=device)
        x.scatter_(1, idx, 1.0)
    else:
        x = torch.rand(n, F, generator=g, device=device)
    return x / x.sum(1, keepdim=True).clamp(min=1e-12)


def powerlaw_sizes(n, alpha, max_size, generator, device):
    """Cluster sizes drawn from a truncated power law P(s) ~ s^-alpha, s in [1, max_size], cut so that they sum to n.
    alpha = 1.8, max_size = 500 gives mean ~7, p99 ~190 — the heavy tail real coarsenings show (core size mean
    10 / p99 178 / max 490 on the Physics-shaped graph), unlike the near-Poisson sizes of a uniform assignment."""
    s = torch.arange(1, max_size + 1, device=device, dtype=torch.float64)
    cdf = torch.cumsum(s ** -alpha, 0)
    cdf = cdf / cdf[-1]
    mean = float((s * (s ** -alpha)).sum() / (s ** -alpha).sum())
    sizes = torch.zeros(0, dtype=torch.long, device=device)
    while int(sizes.sum()) < n:
        m = int(1.1 * (n - int(sizes.sum())) / mean) + 16
        u = torch.rand(m, generator=generator, device=device, dtype=torch.float64)
        sizes = torch.cat([sizes, torch.searchsorted(cdf, u).clamp(max=max_size - 1) + 1])
    csum = torch.cumsum(sizes, 0)
    k = int(torch.searchsorted(csum, torch.tensor([n], device=device)).item()) + 1
    sizes = sizes[:k].clone()
    sizes[-1] -= int(csum[k - 1]) - n  # the last cluster takes what is left
    return sizes


def planted_partition(n, n_undirected, ratio, seed=0, device="cuda", locality=256, intra_extra=0.5, sizes="uniform",
                      alpha=1.8, max_size=500):
    """ogbn-products-shaped graph with clusters known by construction, generated on `device`.

    sizes='uniform': ~ratio*n clusters with multinomial sizes (mean 1/ratio); sizes='powerlaw': heavy-tailed sizes
    (`powerlaw_sizes(n, alpha, max_size)`; `ratio` is ignored, k follows from the size law).  Every cluster is connected
    by a random tree plus `intra_extra`*(size-1) extra internal edges; the remaining undirected edges join a node to a
    node of a nearby cluster (|offset| ~ Laplace(locality) in cluster-contiguous order), which gives the coarsened graph
    a banded, community-like pattern.  Node ids are then shuffled so that nothing is pre-sorted.
    Returns (edge_index int64 [2, 2E] both directions, part int32 [n], cweight float64 [n], k)."""
    g = torch.Generator(device=device).manual_seed(seed)
    if sizes == "powerlaw":
        sizes = powerlaw_sizes(n, alpha, max_size, g, device)
        k = sizes.numel()
        part_sorted = torch.repeat_interleave(torch.arange(k, device=device), sizes)
    elif sizes == "uniform":
        k0 = max(1, int(round(ratio * n)))
        cid = torch.sort(torch.randint(0, k0, (n,), generator=g, device=device)).values
        uniq, part_sorted = torch.unique_consecutive(cid, return_inverse=True)
        k = uniq.numel()
        sizes = torch.bincount(part_sorted, minlength=k)
    else:
        raise ValueError(f"sizes={sizes!r}")
    start = torch.cumsum(sizes, 0) - sizes
    pos_in = torch.arange(n, device=device) - start[part_sorted]
    # random tree inside each cluster: node -> a random earlier node of the same cluster
    has_parent = pos_in > 0
    child = torch.nonzero(has_parent).view(-1)
    r = torch.rand(child.numel(), generator=g, device=device)
    parent = start[part_sorted[child]] + (r * pos_in[child]).long()
    src, dst = [child], [parent]
    n_extra = int(intra_extra * child.numel())
    if n_extra > 0:
        c2 = child[torch.randint(0, child.numel(), (n_extra,), generator=g, device=device)]
        r2 = torch.rand(n_extra, generator=g, device=device)
        p2 = start[part_sorted[c2]] + (r2 * sizes[part_sorted[c2]]).long()
        ok = p2 != c2
        src.append(c2[ok]); dst.append(p2[ok])
    n_inter = max(0, n_undirected - sum(s.numel() for s in src))
    if n_inter > 0:
        u = torch.randint(0, n, (n_inter,), generator=g, device=device)
        mag = -locality * torch.log(torch.rand(n_inter, generator=g, device=device).clamp(min=1e-12))
        sign = torch.randint(0, 2, (n_inter,), generator=g, device=device) * 2 - 1
        v = (u + sign * (mag.long() + 1)).clamp(0, n - 1)
        ok = part_sorted[u] != part_sorted[v]
        src.append(u[ok]); dst.append(v[ok])
    src, dst = torch.cat(src), torch.cat(dst)
    perm = torch.randperm(n, generator=g, device=device)  # sorted position -> node id
    part = torch.empty(n, dtype=torch.int32, device=device)
    part[perm] = part_sorted.to(torch.int32)
    cw = torch.empty(n, dtype=torch.float64, device=device)
    cw[perm] = 1.0 / torch.sqrt(sizes[part_sorted].double())
    a, b = perm[src], perm[dst]
    ei = torch.stack([torch.cat([a, b]), torch.cat([b, a])])
    shuffle = torch.randperm(ei.shape[1], generator=g, device=device)
    return ei[:, shuffle].contiguous(), part, cw, int(k)


def relabel_partition_reference_order(part, device=None):
    """Renumber clusters by their smallest member (the reference numbers supernodes that way,
    coarsening_utils.py:174-178) so a planted partition follows the same convention."""
    part = part.long()
    n = part.numel()
    k = int(part.max().item()) + 1
    first = torch.full((k,), n, dtype=torch.long, device=part.device)
    first.scatter_reduce_(0, part, torch.arange(n, device=part.device), reduce="amin")
    order = torch.argsort(first)
    new_id = torch.empty(k, dtype=torch.long, device=part.device)
    new_id[order] = torch.arange(k, device=part.device)
    return new_id[part].to(torch.int32)


def molecule_graphs(n_graphs, seed=0, mean_nodes=23):
    """ZINC-shaped small graphs: random tree + a few ring-closing edges, atom type in [0, 21).  Returns a list of
    (x int64 [n,1], edge_index int64 [2,2e], y float32 [1])."""
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(n_graphs):
        n = int(np.clip(rng.normal(mean_nodes, 4), 8, 38))
        e = [(int(rng.integers(max(0, i - 3), i)), i) for i in range(1, n)]
        for _ in range(int(rng.integers(1, 4))):
            a = int(rng.integers(0, n - 5))
            e.append((a, a + 5))
        und = np.unique(np.array(e, dtype=np.int64), axis=0)
        ei = np.concatenate([und, und[:, ::-1]], 0).T
        out.append((rng.integers(0, 21, (n, 1)).astype(np.int64), np.ascontiguousarray(ei),
                    rng.normal(size=(1,)).astype(np.float32)))
    return out


def init_state_dict(num_features, hidden, num_classes, num_layers=2, seed=0, bias_scale=0.1):
    """Random-init parameters under the reference's state_dict keys (network.py:11-22): lin.weight glorot-uniform
    like PyG; biases non-zero so the bias paths are exercised (there are no checkpoints to download)."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    dims = [num_features] + [hidden] * num_layers
    for i in range(num_layers):
        a = (6.0 / (dims[i] + dims[i + 1])) ** 0.5
        sd[f"conv.{i}.lin.weight"] = (torch.rand(dims[i + 1], dims[i], generator=g) * 2 - 1) * a
        sd[f"conv.{i}.bias"] = (torch.rand(dims[i + 1], generator=g) * 2 - 1) * bias_scale
    a = (1.0 / hidden) ** 0.5
    sd["lt1.weight"] = (torch.rand(num_classes, hidden, generator=g) * 2 - 1) * a
    sd["lt1.bias"] = (torch.rand(num_classes, generator=g) * 2 - 1) * a
    return sd
