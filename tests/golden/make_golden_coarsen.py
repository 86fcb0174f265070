"""Generates tests/golden/coarsen_algo.npz: the UNMODIFIED reference coarsening algorithm
(/root/reference/graph_coarsening/coarsening_utils.py: coarsen :18-182 with contract_variation_linear :530-650,
get_coarsening_matrix :212-254, coarsen_matrix :201-205, graph_utils.zero_diag) run behind oracle/ref_shims.py on seeded
connected graphs, methods 'variation_neighborhoods' (the reference's default, utils.py:159), 'variation_cliques' and 'variation_edges'
(contract_variation_edges :483-527, matching_greedy :931-989).

The reference obtains its spectral basis from scipy's eigsh with a RANDOM start vector and tol = 1e-5 (:84-89); the contraction
order depends on it — two calls of the reference on the same graph differ (recorded below as `own_eigsh_diff`).  The fixture
therefore passes (Uk, lk) in through coarsen's own arguments (:26-27, :78-83) so that everything after the eigen-solver is
pinned bit-exactly.  Run in the authoring container only:  python tests/golden/make_golden_coarsen.py"""
import os
import sys
import tempfile

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_shims  # noqa: E402

ref_shims.install()
_cwd = os.getcwd()
os.chdir(tempfile.mkdtemp())
sys.path.insert(0, ref_shims.REFERENCE_ROOT)
from graph_coarsening.coarsening_utils import coarsen  # noqa: E402  (reference)
import pygsp  # noqa: E402  (the shim)
os.chdir(_cwd)
from tests.golden.make_golden import synth_graph  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
K = 10
CASES = [("n60_r50", 11, 60, 0.5, 4.0), ("n300_r30", 3, 300, 0.3, 4.0), ("n300_r50", 3, 300, 0.5, 4.0), ("n300_r70", 3, 300, 0.7, 4.0),
         ("n800_r60", 5, 800, 0.6, 4.0), ("n800_r90", 5, 800, 0.9, 4.0),
         ("n400_d10_r60", 9, 400, 0.6, 10.0)]  # denser: hundreds of triangles, i.e. induced edges inside the candidate sets
EDGE_CASES = [("ve_n300_r30", "variation_edges", 3, 300, 0.3, 4.0), ("ve_n300_r60", "variation_edges", 3, 300, 0.6, 4.0),
              ("ve_n800_r70", "variation_edges", 5, 800, 0.7, 4.0), ("ve_n400_d10_r50", "variation_edges", 9, 400, 0.5, 10.0),
              ("vc_n300_r40", "variation_cliques", 3, 300, 0.4, 4.0), ("vc_n400_d10_r60", "variation_cliques", 9, 400, 0.6, 10.0),
              ("vc_n800_r80", "variation_cliques", 5, 800, 0.8, 4.0)]


def main():
    out = {"cases": np.array([c[0] for c in CASES]), "K": np.int64(K), "edge_cases": np.array([c[0] for c in EDGE_CASES]),
           "edge_methods": np.array([c[1] for c in EDGE_CASES])}
    for name, method, seed, n_main, r, avg_deg in [(c[0], "variation_neighborhoods") + c[1:] for c in CASES] + EDGE_CASES:
        n, ei = synth_graph(seed, n_main, [], avg_deg=avg_deg)
        W = sp.coo_matrix((np.ones(ei.shape[1]), (ei[0], ei[1])), shape=(n, n)).tocsr()
        W.data[:] = 1.0  # simple graph
        G = pygsp.graphs.Graph(W=W)
        offset = 2 * max(G.dw)
        T = offset * sp.eye(G.N, format="csc") - G.L
        lk, Uk = spla.eigsh(T, k=K, which="LM", tol=1e-5, v0=np.random.default_rng(seed).standard_normal(G.N))
        lk = (offset - lk)[::-1].copy()
        Uk = Uk[:, ::-1].copy()
        C, Gc, maps = coarsen(G, K=K, r=r, method=method, Uk=Uk.copy(), lk=lk.copy())
        C = sp.csc_matrix(C)
        C2, _, _ = coarsen(G, K=K, r=r, method=method)  # the reference's own eigsh (random start vector)
        C2 = sp.csc_matrix(C2)
        diff = -1 if C2.shape != C.shape else int((C2.indices != C.indices).sum())
        Wc = sp.coo_matrix(Gc.W)
        coo = W.tocoo()
        out.update({f"{name}_n": np.int64(n), f"{name}_r": np.float64(r), f"{name}_row": coo.row.astype(np.int64),
                    f"{name}_col": coo.col.astype(np.int64), f"{name}_lk": lk, f"{name}_Uk": Uk,
                    f"{name}_C_indices": C.indices.astype(np.int64), f"{name}_C_data": C.data, f"{name}_C_rows": np.int64(C.shape[0]),
                    f"{name}_levels": np.int64(len(maps)), f"{name}_own_eigsh_diff": np.int64(diff),
                    f"{name}_Wc_row": Wc.row.astype(np.int64), f"{name}_Wc_col": Wc.col.astype(np.int64), f"{name}_Wc_val": Wc.data})
        print(name, "n", n, "->", C.shape[0], "levels", len(maps), "partition entries that differ with the reference's own eigsh:", diff)
    np.savez_compressed(os.path.join(OUT, "coarsen_algo.npz"), **out)


if __name__ == "__main__":
    main()
