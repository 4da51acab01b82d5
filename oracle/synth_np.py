"""TEST INFRASTRUCTURE — host (numpy) generator of synthetic negative-binomial count matrices.

Small-case twin of the device generator (alona_b200/csrc/ab_synth.cu): same model
(SURVEY.md §8(d)), different random stream.  Used by tests/ and oracle/make_golden.py to
make seeded inputs that both the oracle and the CUDA path consume.  Not product code.

Model: C clusters; gene base abundance p_g ~ LogNormal(0, 2) normalised; cluster profile
p_{g|c} = p_g * exp(N(0,1)) on a random 10 % of genes (else x1), renormalised; cell depth
L_i ~ LogNormal(log Lbar, 0.35); counts ~ NB(mean L_i * p_{g|c(i)}, inverse dispersion r=5)
drawn as Poisson(Gamma(r, mu/r)).
"""
import numpy as np
import scipy.sparse as sp


def nb_counts(n_cells, n_genes, lbar=2500.0, n_clusters=30, r=5.0, seed=0,
              drop_empty=True):
    """Returns (csr int32 cells x genes, cluster id per cell).

    Every cell is guaranteed S_c >= 1 and (drop_empty) all-zero genes are removed, which
    is what the reference's filters guarantee upstream of the path
    (alona/cell.py:64-81,141-173).
    """
    rng = np.random.RandomState(seed)
    p = np.exp(rng.normal(0.0, 2.0, n_genes))
    p /= p.sum()
    prof = np.tile(p, (n_clusters, 1))
    for c in range(n_clusters):
        sel = rng.rand(n_genes) < 0.10
        prof[c, sel] *= np.exp(rng.normal(0.0, 1.0, sel.sum()))
        prof[c] /= prof[c].sum()
    cl = rng.randint(0, n_clusters, n_cells)
    depth = np.exp(rng.normal(np.log(lbar), 0.35, n_cells))
    rows = []
    chunk = 2048
    for a in range(0, n_cells, chunk):
        b = min(n_cells, a + chunk)
        mu = depth[a:b, None] * prof[cl[a:b]]
        lam = rng.gamma(r, mu / r)
        x = rng.poisson(lam).astype(np.int32)
        rows.append(sp.csr_matrix(x))
    m = sp.vstack(rows).tocsr()
    # every cell must have at least one count
    empty = np.where(np.asarray(m.sum(axis=1)).ravel() == 0)[0]
    if empty.size:
        m = m.tolil()
        for i in empty:
            m[i, rng.randint(0, n_genes)] = 1
        m = m.tocsr()
    if drop_empty:
        keep = np.asarray((m != 0).sum(axis=0)).ravel() > 0
        m = m[:, keep].tocsr()
    m.sort_indices()
    m = m.astype(np.int32)
    return m, cl


def gaussian_mixture_embedding(n, d=50, n_clusters=30, seed=2, dtype=np.float64):
    """C4-style embedding: 30-cluster Gaussian mixture in d dims, per-dim scale
    linspace(2, 0.5), centres N(0, 6^2) * linspace(3, 0.3)  (SURVEY.md §8(d))."""
    rng = np.random.RandomState(seed)
    centres = rng.normal(0.0, 6.0, (n_clusters, d)) * np.linspace(3.0, 0.3, d)
    scale = np.linspace(2.0, 0.5, d)
    cl = rng.randint(0, n_clusters, n)
    x = centres[cl] + rng.normal(0.0, 1.0, (n, d)) * scale
    return np.ascontiguousarray(x.astype(dtype))
