"""Synthetic negative-binomial count matrices for the bench and tests:
cluster profiles are drawn on the host (tiny), counts are generated on the device by
ab_synth_nb_* as a pure function of (seed, cell, gene) so every sharding sees the same matrix."""
import numpy as np

WORKLOADS = {
    # name: cells, genes, hvg_n, pca_n, k, lbar (depth giving the target genes-detected/cell)
    'c1': dict(cells=2700, genes=32738, hvg_n=1000, pca_n=40, k=10, lbar=1061.0),
    'c2': dict(cells=68579, genes=32738, hvg_n=2000, pca_n=50, k=20, lbar=599.0),
    'c3': dict(cells=1306127, genes=27998, hvg_n=2000, pca_n=50, k=20, lbar=3289.5),
    'c5': dict(cells=4000000, genes=30000, hvg_n=3000, pca_n=100, k=30, lbar=2191.8),
}


def make_profiles(n_genes, n_clusters=30, seed=20251019):
    """p_g ~ LogNormal(0,2); cluster profile = p_g * exp(N(0,1)) on a random 10 % of genes;
    rows renormalised.  Returns float32 [n_clusters, n_genes]."""
    rng = np.random.RandomState(seed)
    p = np.exp(rng.normal(0.0, 2.0, n_genes))
    p /= p.sum()
    prof = np.tile(p, (n_clusters, 1))
    for c in range(n_clusters):
        sel = rng.rand(n_genes) < 0.10
        prof[c, sel] *= np.exp(rng.normal(0.0, 1.0, int(sel.sum())))
        prof[c] /= prof[c].sum()
    return np.ascontiguousarray(prof.astype(np.float32))
