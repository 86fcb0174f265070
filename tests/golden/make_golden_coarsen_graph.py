"""Seeded connected test graph shared by the coarsening tests (no reference needed): the preferential-attachment generator of
tests/golden/make_golden.py without its reference imports."""
import numpy as np


def synth_graph(seed, n_main, avg_deg=4.0):
    rng = np.random.default_rng(seed)
    edges = set()
    targets = [0]
    for v in range(1, n_main):
        m = max(1, int(rng.poisson(avg_deg / 2)))
        for u in rng.choice(targets, size=min(m, len(targets)), replace=True):
            if u != v:
                edges.add((min(u, v), max(u, v)))
        targets.extend([v] * m)
        targets.append(int(rng.integers(0, v + 1)))
    und = np.array(sorted(edges), dtype=np.int64)
    und = rng.permutation(n_main)[und]
    ei = np.concatenate([und, und[:, ::-1]], 0)
    return n_main, np.ascontiguousarray(ei[rng.permutation(len(ei))].T)
