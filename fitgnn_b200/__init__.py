"""fitgnn_b200 — B200-native (sm_100a) implementation of FIT-GNN's data-parallel hot path:
GCN message passing over the coarsening-derived subgraphs Gs and the coarsened-graph projection Gc.

Importing the package loads libfitgnn_b200.so (C ABI in include/fitgnn.h); there is no CPU fallback.
"""
from . import _lib

_lib.lib()  # fail loudly at import if the CUDA library has not been built

from . import autograd, cache, coarsen, coarsen_algo, dist, engine, infer, nn, ops, pack, stream, synth, train  # noqa: E402
from .engine import PackedForward  # noqa: E402
from .nn import (Classify_graph_gc, Classify_graph_gs, Classify_node, GCNConv, Net1, Net2, Regress_graph_gc,  # noqa: E402
                 Regress_graph_gs, Regress_node)
from .pack import Pack, PackStream, build_pack, build_pack_range, build_pack_stream, pack_from_subgraph_list  # noqa: E402
from .stream import StreamedForward  # noqa: E402

__all__ = ["GCNConv", "Classify_node", "Regress_node", "Classify_graph_gc", "Classify_graph_gs", "Regress_graph_gc",
           "Regress_graph_gs", "Net1", "Net2", "Pack", "PackStream", "build_pack", "build_pack_range", "build_pack_stream", "pack_from_subgraph_list", "PackedForward",
           "StreamedForward", "stream", "ops", "coarsen", "infer",
           "synth", "nn", "engine", "pack"]
