"""TEST INFRASTRUCTURE ONLY — never imported by the product package.

Minimal stand-ins for the third-party modules the reference imports at module top
(`pygsp`, `matplotlib`, `torch_geometric`, `torch_scatter`, `torch_sparse`, `igraph`, `leidenalg`),
none of which is installed in this image.  With them registered in `sys.modules` the
UNMODIFIED reference files under /root/reference (`graph_coarsening/coarsening_utils.py`, `utils.py`,
`network.py`) import and run in this container, which is how `tests/golden/make_golden.py` produces
fixtures from the reference's own code.

What each stand-in restates (the behaviour of the library version the reference targets):
  * pygsp.graphs.Graph            — pygsp 0.5.1 `Graph` (W as LIL, A = W > 0, combinatorial L, dw, Ne,
                                    extract_components, subgraph, get_edge_list); the reference ships its
                                    own copy of extract_components at utils.py:73-104 with the same body.
  * torch_geometric.data.Data     — attribute bag with `.subgraph(subset)` = node-induced subgraph with
                                    relabelled, order-preserving edges (PyG >= 2.1 `Data.subgraph`).
  * torch_geometric.nn.GCNConv    — the oracle's restatement of PyG GCNConv (oracle.fitgnn_oracle.gcn_conv).
PyG itself is absent, so the GCNConv arithmetic stays a restatement: parity for that operator is
anchored on the known-answer vector of SURVEY.md §8c, not on PyG's own output ("parity unpinned").
"""
import sys
import types

import numpy as np
import scipy.sparse as sp
import torch

REFERENCE_ROOT = "/root/reference"


# ------------------------------------------------------------------------------------------ pygsp
class Graph:
    """pygsp 0.5.1 graphs.Graph, the subset graph_coarsening/ and utils.py touch."""

    def __init__(self, W, gtype="unknown", lap_type="combinatorial", coords=None, plotting=None):
        if sp.issparse(W):
            W = sp.lil_matrix(W)
        else:
            W = sp.lil_matrix(np.asarray(W))
        if W.shape[0] != W.shape[1]:
            raise ValueError("W has incorrect shape")
        self.N = W.shape[0]
        self.W = W
        self.gtype = gtype
        self.lap_type = lap_type
        self.info = {}
        if coords is not None:
            self.coords = coords
        self._A = None
        self._L = None
        # pygsp: Ne counts undirected edges once (upper triangle incl. diagonal)
        self.Ne = int(sp.triu(self.W).nnz)

    @property
    def A(self):
        if self._A is None:
            self._A = self.W > 0
        return self._A

    def is_directed(self):
        return (abs(self.W - self.W.T) > 1e-12).nnz != 0

    @property
    def dw(self):
        return np.ravel(self.W.sum(axis=0))

    @property
    def d(self):
        return np.ravel(self.A.sum(axis=1))

    @property
    def L(self):
        if self._L is None:
            D = sp.diags(self.dw, 0)
            self._L = (D - self.W).tocsc()
        return self._L

    def get_edge_list(self):
        v_in, v_out = sp.tril(self.W).nonzero()
        weights = self.W[v_in, v_out]
        weights = (weights.toarray() if sp.issparse(weights) else np.asarray(weights)).squeeze()  # pygsp: .toarray().squeeze()
        return v_in, v_out, weights

    def subgraph(self, ind):
        sub_W = self.W.tocsr()[ind, :].tocsc()[:, ind]
        return Graph(sub_W, gtype="sub-{}".format(self.gtype))

    def extract_components(self):
        graphs = []
        visited = np.zeros(self.A.shape[0], dtype=bool)
        A = self.A.tocsr()
        while not visited.all():
            stack = set([np.nonzero(~visited)[0][0]])
            comp = []
            while len(stack):
                v = stack.pop()
                if not visited[v]:
                    comp.append(v)
                    visited[v] = True
                    stack.update(set([idx for idx in A[v, :].nonzero()[1] if not visited[idx]]))
            comp = sorted(comp)
            G = self.subgraph(comp)
            G.info = {"orig_idx": comp}
            graphs.append(G)
        return graphs


# ------------------------------------------------------------------------------- torch_geometric
def pyg_subgraph(subset, edge_index, relabel_nodes=False, num_nodes=None):
    """torch_geometric.utils.subgraph for an index subset: keep edges with both ends in `subset`,
    in their original order; relabel to the position of the node in `subset`."""
    subset = torch.as_tensor(subset, dtype=torch.long)
    if num_nodes is None:
        num_nodes = int(max(int(edge_index.max()) + 1 if edge_index.numel() else 0,
                            int(subset.max()) + 1 if subset.numel() else 0))
    node_mask = torch.zeros(num_nodes, dtype=torch.bool)
    node_mask[subset] = True
    edge_mask = node_mask[edge_index[0]] & node_mask[edge_index[1]]
    ei = edge_index[:, edge_mask]
    if relabel_nodes:
        node_idx = torch.zeros(num_nodes, dtype=torch.long)
        node_idx[subset] = torch.arange(subset.numel())
        ei = node_idx[ei]
    return ei, None


class Data:
    """torch_geometric.data.Data: an attribute bag; node-level tensors are those whose first
    dimension equals num_nodes."""

    def __init__(self, x=None, edge_index=None, y=None, **kwargs):
        self.x, self.edge_index, self.y = x, edge_index, y
        for k, v in kwargs.items():
            setattr(self, k, v)

    @property
    def num_nodes(self):
        if self.x is not None:
            return self.x.shape[0]
        return int(self.edge_index.max()) + 1

    def keys(self):
        return [k for k in self.__dict__ if not k.startswith("_")]

    def to(self, device):
        for k in self.keys():
            v = getattr(self, k)
            if torch.is_tensor(v):
                setattr(self, k, v.to(device))
        return self

    def cpu(self):
        return self.to("cpu")

    def subgraph(self, subset):
        subset = torch.as_tensor(subset)
        if subset.dtype == torch.bool:
            subset = subset.nonzero().view(-1)
        subset = subset.long().cpu()
        n = self.num_nodes
        ei, _ = pyg_subgraph(subset, self.edge_index.cpu(), relabel_nodes=True, num_nodes=n)
        out = Data()
        for k in self.keys():
            v = getattr(self, k)
            if k == "edge_index":
                out.edge_index = ei
            elif torch.is_tensor(v) and v.dim() > 0 and v.shape[0] == n and k != "edge_index":
                setattr(out, k, v[subset])
            else:
                setattr(out, k, v)
        return out


class Batch(Data):
    @staticmethod
    def from_data_list(data_list):
        xs, eis, ys, batch, ptr = [], [], [], [], [0]
        extra = {}
        off = 0
        for i, d in enumerate(data_list):
            n = d.x.shape[0]
            xs.append(d.x)
            eis.append(d.edge_index + off)
            if d.y is not None:
                ys.append(d.y)
            batch.append(torch.full((n,), i, dtype=torch.long))
            for k in d.keys():
                v = getattr(d, k)
                if k in ("x", "edge_index", "y"):
                    continue
                if torch.is_tensor(v) and v.dim() > 0 and v.shape[0] == n:
                    extra.setdefault(k, []).append(v)
            off += n
            ptr.append(off)
        b = Batch(x=torch.cat(xs, 0), edge_index=torch.cat(eis, 1), y=torch.cat(ys, 0) if ys else None)
        b.batch = torch.cat(batch, 0)
        b.ptr = torch.tensor(ptr)
        for k, v in extra.items():
            if len(v) == len(data_list):
                setattr(b, k, torch.cat(v, 0))
        return b


class DataLoader:
    """torch_geometric.loader.DataLoader with shuffle=False: consecutive chunks -> Batch."""

    def __init__(self, dataset, batch_size=1, shuffle=False, **kw):
        assert not shuffle, "shim supports shuffle=False only"
        self.dataset, self.batch_size = list(dataset), batch_size

    def __len__(self):
        return (len(self.dataset) + self.batch_size - 1) // self.batch_size

    def __iter__(self):
        for i in range(0, len(self.dataset), self.batch_size):
            yield Batch.from_data_list(self.dataset[i:i + self.batch_size])


def to_scipy_sparse_matrix(edge_index, edge_attr=None, num_nodes=None):
    row, col = edge_index.cpu().numpy()
    if num_nodes is None:
        num_nodes = int(max(row.max(), col.max())) + 1
    data = np.ones(row.shape[0]) if edge_attr is None else edge_attr.cpu().numpy()
    return sp.coo_matrix((data, (row, col)), (num_nodes, num_nodes))


def degree(index, num_nodes=None, dtype=None):
    n = int(index.max()) + 1 if num_nodes is None else num_nodes
    out = torch.zeros(n, dtype=dtype or torch.float)
    return out.scatter_add_(0, index, torch.ones(index.numel(), dtype=out.dtype))


def to_dense_adj(edge_index, max_num_nodes=None):
    n = int(edge_index.max()) + 1 if max_num_nodes is None else max_num_nodes
    adj = torch.zeros(1, n, n)
    adj[0, edge_index[0], edge_index[1]] = 1
    return adj


def global_max_pool(x, batch, size=None):
    size = int(batch.max()) + 1 if size is None else size
    out = torch.full((size, x.shape[1]), float("-inf"), dtype=x.dtype)
    out = out.scatter_reduce(0, batch.view(-1, 1).expand_as(x), x, reduce="amax", include_self=True)
    return torch.where(torch.isinf(out), torch.zeros_like(out), out)


def global_mean_pool(x, batch, size=None):
    size = int(batch.max()) + 1 if size is None else size
    s = torch.zeros((size, x.shape[1]), dtype=x.dtype).index_add_(0, batch, x)
    cnt = torch.zeros(size, dtype=x.dtype).index_add_(0, batch, torch.ones(batch.numel(), dtype=x.dtype))
    return s / cnt.clamp(min=1).view(-1, 1)


def _make_gcnconv():
    from oracle import fitgnn_oracle as fo

    class GCNConv(torch.nn.Module):
        """PyG GCNConv surface (lin.weight [out,in] glorot, bias zeros) on the oracle arithmetic."""

        def __init__(self, in_channels, out_channels, **kw):
            super().__init__()
            self.in_channels, self.out_channels = in_channels, out_channels
            self.lin = torch.nn.Linear(in_channels, out_channels, bias=False)
            self.bias = torch.nn.Parameter(torch.zeros(out_channels))
            self.reset_parameters()

        def reset_parameters(self):
            torch.nn.init.xavier_uniform_(self.lin.weight)
            torch.nn.init.zeros_(self.bias)

        def forward(self, x, edge_index):
            return fo.gcn_conv_torch(x, edge_index, self.lin.weight, self.bias)

    return GCNConv


def install():
    """Register the stand-ins and put /root/reference on sys.path.  Idempotent."""
    if "pygsp" in sys.modules and getattr(sys.modules["pygsp"], "_fitgnn_shim", False):
        return
    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    graphs = mod("pygsp.graphs", Graph=Graph)
    pygsp = mod("pygsp", graphs=graphs, filters=mod("pygsp.filters"), reduction=mod("pygsp.reduction"))
    pygsp._fitgnn_shim = True

    pylab = mod("matplotlib.pylab")
    mod("matplotlib", pylab=pylab)
    mod("matplotlib.pyplot")
    m3 = mod("mpl_toolkits.mplot3d", Axes3D=object)
    mod("mpl_toolkits", mplot3d=m3)

    GCNConv = _make_gcnconv()
    tg_nn = mod("torch_geometric.nn", GCNConv=GCNConv, global_max_pool=global_max_pool,
                global_mean_pool=global_mean_pool)
    tg_data = mod("torch_geometric.data", Data=Data, Batch=Batch)
    tg_loader = mod("torch_geometric.loader", DataLoader=DataLoader)
    tg_utils = mod("torch_geometric.utils", subgraph=pyg_subgraph, to_scipy_sparse_matrix=to_scipy_sparse_matrix,
                   degree=degree, to_dense_adj=to_dense_adj)
    tg_profile = mod("torch_geometric.profile", get_data_size=lambda d: 0)
    mod("torch_geometric", nn=tg_nn, data=tg_data, loader=tg_loader, utils=tg_utils, profile=tg_profile)
    mod("torch_scatter")
    mod("torch_sparse")
    mod("igraph")
    mod("leidenalg")
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
