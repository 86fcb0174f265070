"""The reference's on-disk preprocessing cache: read it, turn it into packs, write it.

Layout written by `save()` /root/reference/main.py:131-172 and read at main.py:270-275, :361-367 and inference.py:543-548:

    ./dataset/<name>/saved/<coarsening_method>/<ratio>_<node_type>_<graph_type>_subgraph_list.pt   torch.save
                                              ..._candidate.pkl  ..._C_list.pkl  ..._Gc_list.pkl   pickle   (node_cls)
                                              ..._Gc_list.pkl  ..._saved_graph_list.pkl            pickle   (graph tasks)
    node_type  d | e | c  = default / extra_node / cluster_node (cluster_node wins, main.py:117-121, :134-138)
    graph_type full | community                                   (main.py:139-142)

`subgraph_list.pt` and `candidate.pkl` pickle torch_geometric `Data` / pygsp `Graph` objects.  With those packages installed
they un-pickle as themselves; without them (`stand_ins=True`, the default when the import fails) the reader substitutes
state-holding stand-ins for every class of the two packages: a pickled PyG `Data` is `{'_store': GlobalStorage}` whose state
holds the attribute dict `_mapping` (torch_geometric/data/data.py, storage.py `__getstate__`), a pygsp `Graph` is its
`__dict__` — enough for everything downstream, which is duck-typed: a subgraph is anything with
`.x / .edge_index / .mask / .orig_idx`, a candidate anything with `.info['orig_idx']`, a C matrix anything scipy can convert.
"""
from __future__ import annotations

import os
import pickle
from dataclasses import dataclass, field

import numpy as np
import torch

from .coarsen import Partition, partition_from_components

_FILES = ("subgraph_list.pt", "candidate.pkl", "C_list.pkl", "Gc_list.pkl", "saved_graph_list.pkl")


def node_type_of(extra_node: bool = False, cluster_node: bool = False) -> str:
    """main.py:134-138 after arg_correction (main.py:117-121: cluster_node wins when both flags are given)."""
    if cluster_node:
        return "c"
    return "e" if extra_node else "d"


def cache_dir(root: str, dataset: str, coarsening_method: str) -> str:
    return os.path.join(root, "dataset", dataset, "saved", coarsening_method)


def cache_paths(root, dataset, coarsening_method, coarsening_ratio, extra_node=False, cluster_node=False,
                use_community_detection=False) -> dict:
    """{'subgraph_list': path, 'candidate': ..., 'C_list': ..., 'Gc_list': ..., 'saved_graph_list': ...}"""
    graph_type = "community" if use_community_detection else "full"
    stem = f"{coarsening_ratio}_{node_type_of(extra_node, cluster_node)}_{graph_type}_"
    d = cache_dir(root, dataset, coarsening_method)
    return {f.split(".")[0]: os.path.join(d, stem + f) for f in _FILES}


@dataclass
class ReferenceCache:
    """What the reference keeps between runs.  Node tasks: subgraph_list = list[Data]; graph tasks: list (graphs) of
    list[Data], Gc_list = list[Data] (one coarsened graph per kept graph), saved_graph_list = dataset indices kept."""
    subgraph_list: list
    candidate: list | None = None
    C_list: list | None = None
    Gc_list: list | None = None
    saved_graph_list: list | None = None
    paths: dict = field(default_factory=dict)

    @property
    def graph_level(self) -> bool:
        return bool(self.subgraph_list) and isinstance(self.subgraph_list[0], (list, tuple))


class _StandIn:
    """An object of a class whose package is not installed: keeps the pickled state, and answers attribute reads the way a PyG
    `Data` does — through `_store._mapping` (the GlobalStorage's attribute dict) — or a storage does (`_mapping`)."""

    def __setstate__(self, state):
        if isinstance(state, tuple) and len(state) == 2 and isinstance(state[1], dict):  # (dict state, slots state)
            self.__dict__.update(state[0] or {})
            self.__dict__.update(state[1])
        elif isinstance(state, dict):
            self.__dict__.update(state)
        else:
            self.__dict__["_state"] = state

    def __getattr__(self, name):
        d = self.__dict__
        store = d.get("_store")
        for mapping in (d.get("_mapping"), store.__dict__.get("_mapping") if store is not None else None):
            if isinstance(mapping, dict) and name in mapping:
                return mapping[name]
        raise AttributeError(name)

    def keys(self):
        store = self.__dict__.get("_store")
        m = self.__dict__.get("_mapping") or (store.__dict__.get("_mapping") if store is not None else None) or {}
        return list(m.keys())


_STAND_IN_PACKAGES = ("torch_geometric", "pygsp")
_stand_in_classes = {}


class _StandInUnpickler(pickle.Unpickler):
    """pickle.Unpickler that resolves classes of the uninstalled reference dependencies to `_StandIn` subclasses."""

    def find_class(self, module, name):
        try:
            return super().find_class(module, name)
        except (ModuleNotFoundError, AttributeError):
            if module.split(".")[0] not in _STAND_IN_PACKAGES:
                raise
            key = (module, name)
            if key not in _stand_in_classes:
                _stand_in_classes[key] = type(name, (_StandIn,), {"__module__": module})
            return _stand_in_classes[key]


class _StandInPickle:
    """the `pickle_module` torch.load takes: Unpickler + load"""
    __name__ = "fitgnn_b200.cache._StandInPickle"
    Unpickler = _StandInUnpickler

    @staticmethod
    def load(f, **kw):
        return _StandInUnpickler(f, **kw).load()


def _have_reference_packages() -> bool:
    """torch_geometric and pygsp importable (or registered in sys.modules, e.g. by a test's stand-in modules)?"""
    import importlib.util
    import sys
    for p in _STAND_IN_PACKAGES:
        if p in sys.modules:
            continue
        try:
            if importlib.util.find_spec(p) is None:
                return False
        except (ImportError, ValueError):
            return False
    return True


def _unpickle(path, loader):
    try:
        return loader(path)
    except ModuleNotFoundError as e:
        raise ModuleNotFoundError(
            f"{path} pickles {e.name} objects (torch_geometric Data / pygsp Graph): un-pickling the reference's cache needs "
            f"that package, or load_reference_cache(..., stand_ins=True)") from e


def load_reference_cache(root, dataset, coarsening_method, coarsening_ratio, extra_node=False, cluster_node=False,
                         use_community_detection=False, stand_ins=None) -> ReferenceCache:
    """Read whatever of the five files exists (subgraph_list.pt is mandatory, as in main.py:270 / :361).
    stand_ins: None = only when torch_geometric / pygsp are not importable; True / False force it."""
    paths = cache_paths(root, dataset, coarsening_method, coarsening_ratio, extra_node, cluster_node, use_community_detection)
    if not os.path.exists(paths["subgraph_list"]):
        raise FileNotFoundError(f"no reference cache at {paths['subgraph_list']}")
    if stand_ins is None:
        stand_ins = not _have_reference_packages()
    kw = {"pickle_module": _StandInPickle} if stand_ins else {}
    out = ReferenceCache(_unpickle(paths["subgraph_list"],
                                   lambda p: torch.load(p, weights_only=False, map_location="cpu", **kw)), paths=paths)
    for key in ("candidate", "C_list", "Gc_list", "saved_graph_list"):
        if os.path.exists(paths[key]):
            with open(paths[key], "rb") as f:
                setattr(out, key, _unpickle(paths[key], lambda p, f=f: (_StandInPickle if stand_ins else pickle).load(f)))
    return out


def save_reference_cache(root, dataset, coarsening_method, coarsening_ratio, task, subgraph_list, candidate=None, C_list=None,
                         Gc_list=None, saved_graph_list=None, extra_node=False, cluster_node=False,
                         use_community_detection=False) -> dict:
    """Write the same files `save()` writes for `task` (main.py:143-171), so the reference's own scripts find them."""
    paths = cache_paths(root, dataset, coarsening_method, coarsening_ratio, extra_node, cluster_node, use_community_detection)
    os.makedirs(os.path.dirname(paths["subgraph_list"]), exist_ok=True)

    def dump(key, obj):
        with open(paths[key], "wb") as f:
            pickle.dump(obj, f)

    if task == "node_cls":
        dump("candidate", candidate); dump("C_list", C_list); dump("Gc_list", Gc_list)
    elif task != "node_reg":  # graph_cls / graph_reg
        dump("Gc_list", Gc_list); dump("saved_graph_list", saved_graph_list)
    torch.save(subgraph_list, paths["subgraph_list"])
    return paths


def partition_from_cache(cache: ReferenceCache, n_nodes: int):
    """-> (Partition, comps, C_list).  The partition vector + C weights of a node-task cache: candidate[i].info['orig_idx'] are the components in
    candidate order (utils.py:144-146) and C_list[i] the coarsening matrix of component i (only components with more than
    one node have one, utils.py:164-166, :352) — the inputs of the device pack builder and of the Gc projection."""
    if cache.candidate is None or cache.C_list is None:
        raise ValueError("partition_from_cache needs candidate.pkl and C_list.pkl (node_cls caches)")
    comps = [np.asarray(h.info["orig_idx"], dtype=np.int64) for h in cache.candidate]
    it = iter(cache.C_list)
    C_list = [next(it) if len(c) > 1 else None for c in comps]
    return partition_from_components(comps, C_list, n_nodes), comps, C_list


def pack_from_reference_cache(cache: ReferenceCache, device="cuda"):
    """Node tasks: (pack, X_packed, node_ids) of the cached subgraph_list (pack.pack_from_subgraph_list).
    Graph tasks: (pack, X_packed, graph_of_sub) with every subgraph of every graph in one pack and graph_of_sub[s] the
    graph a subgraph belongs to — the inputs of infer.graph_level_Gs."""
    from .pack import pack_from_subgraph_list
    if not cache.graph_level:
        return pack_from_subgraph_list(cache.subgraph_list, device)
    flat, graph_of_sub = [], []
    for g, subs in enumerate(cache.subgraph_list):
        flat.extend(subs)
        graph_of_sub.extend([g] * len(subs))
    pack, X, _ = pack_from_subgraph_list(flat, device)
    return pack, X, torch.tensor(graph_of_sub, dtype=torch.long)
