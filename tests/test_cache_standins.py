"""The reference's on-disk cache (main.py:131-172) read WITHOUT torch_geometric / pygsp installed: files are written with
classes that carry PyG's and pygsp's module paths and PyG 2.x's pickled state layout (Data.__dict__ = {'_store': GlobalStorage},
GlobalStorage state = {'_mapping': {...}, '_parent': the Data}; torch_geometric/data/data.py, storage.py), the fake packages
are removed again, and the reader's stand-ins must give back the attributes everything downstream reads."""
import sys
import types
import weakref

import numpy as np
import pytest
import scipy.sparse as sp
import torch


def _install_fakes():
    tgdd = types.ModuleType("torch_geometric.data.data")
    tgds = types.ModuleType("torch_geometric.data.storage")
    pggg = types.ModuleType("pygsp.graphs.graph")

    class GlobalStorage:
        def __init__(self, parent, **kw):
            self._mapping = dict(kw)
            self._parent = weakref.ref(parent)

        def __getstate__(self):  # BaseStorage.__getstate__: the weak reference is replaced by the object
            out = self.__dict__.copy()
            out["_parent"] = out["_parent"]() if out.get("_parent") is not None else None
            return out

        def __setstate__(self, st):
            self.__dict__.update(st)

    class Data:
        def __init__(self, **kw):
            self.__dict__["_store"] = GlobalStorage(self, **kw)

        def __getstate__(self):
            return self.__dict__.copy()

        def __setstate__(self, st):
            self.__dict__.update(st)

    class Graph:
        def __init__(self, W, info):
            self.W, self.N, self.info = W, W.shape[0], info

    GlobalStorage.__module__, GlobalStorage.__qualname__ = "torch_geometric.data.storage", "GlobalStorage"
    Data.__module__, Data.__qualname__ = "torch_geometric.data.data", "Data"
    Graph.__module__, Graph.__qualname__ = "pygsp.graphs.graph", "Graph"
    tgdd.Data, tgds.GlobalStorage, pggg.Graph = Data, GlobalStorage, Graph
    mods = {"torch_geometric": types.ModuleType("torch_geometric"), "torch_geometric.data": types.ModuleType("torch_geometric.data"),
            "torch_geometric.data.data": tgdd, "torch_geometric.data.storage": tgds, "pygsp": types.ModuleType("pygsp"),
            "pygsp.graphs": types.ModuleType("pygsp.graphs"), "pygsp.graphs.graph": pggg}
    sys.modules.update(mods)
    return Data, Graph, list(mods)


def test_reference_cache_reads_without_pyg_or_pygsp(tmp_path):
    import fitgnn_b200.cache as fc
    if fc._have_reference_packages():
        pytest.skip("torch_geometric / pygsp are installed here: the real classes un-pickle")
    Data, Graph, names = _install_fakes()
    try:
        g = torch.Generator().manual_seed(0)
        subs = [Data(x=torch.randn(4, 3, generator=g), edge_index=torch.tensor([[0, 1, 2], [1, 0, 3]]),
                     mask=torch.tensor([True, True, False, False]), orig_idx=torch.tensor([5, 9, 2, 7]),
                     y=torch.zeros(4, dtype=torch.long)) for _ in range(3)]
        want_x = [s._store._mapping["x"].clone() for s in subs]
        cand = [Graph(sp.lil_matrix(np.eye(3)), {"orig_idx": [0, 2, 3]}), Graph(sp.lil_matrix(np.eye(1)), {"orig_idx": [1]})]
        C = sp.csc_matrix(np.array([[2 ** -0.5, 2 ** -0.5, 0.0], [0.0, 0.0, 1.0]]))
        fc.save_reference_cache(str(tmp_path), "cora", "variation_neighborhoods", 0.3, "node_cls", subs, candidate=cand,
                                C_list=[C], Gc_list=cand[:1], extra_node=True)
    finally:
        for n_ in names:
            del sys.modules[n_]
    with pytest.raises(ModuleNotFoundError, match="stand_ins"):
        fc.load_reference_cache(str(tmp_path), "cora", "variation_neighborhoods", 0.3, extra_node=True, stand_ins=False)
    c = fc.load_reference_cache(str(tmp_path), "cora", "variation_neighborhoods", 0.3, extra_node=True)  # stand-ins by default here
    assert len(c.subgraph_list) == 3 and not c.graph_level
    for s, x in zip(c.subgraph_list, want_x):
        assert type(s).__module__ == "torch_geometric.data.data" and torch.equal(s.x, x)
        assert s.edge_index.tolist() == [[0, 1, 2], [1, 0, 3]] and s.mask.tolist() == [True, True, False, False]
        assert s.orig_idx.tolist() == [5, 9, 2, 7] and not hasattr(s, "no_such_attribute")
        assert set(s.keys()) == {"x", "edge_index", "mask", "orig_idx", "y"}
    part, comps, C_list = fc.partition_from_cache(c, 4)
    assert [list(x) for x in comps] == [[0, 2, 3], [1]] and C_list[1] is None and C_list[0].shape == (2, 3)
    assert part.k == 3 and part.part.tolist() == [0, 2, 0, 1] and np.allclose(part.cweight, [2 ** -0.5, 1.0, 2 ** -0.5, 1.0])
