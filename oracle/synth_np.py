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


# ----------------------------------------------------------------------------------------
# numpy twin of the DEVICE generator (alona_b200/csrc/ab_synth.cu): same counter-based hash, same cell draw,
# same NB CDF inversion, so the CPU arm of bench.py can build "the first n cells of the workload matrix" without
# loading the CUDA library.  Identical to the device matrix up to the last-bit differences between the host's and
# the device's float32 log1p/expm1/exp (an entry changes only if its uniform falls within that rounding of a CDF step).
# ----------------------------------------------------------------------------------------
_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _mix64(x):
    x = x.astype(np.uint64, copy=True)
    x ^= x >> np.uint64(30)
    x *= np.uint64(0xbf58476d1ce4e5b9)
    x ^= x >> np.uint64(27)
    x *= np.uint64(0x94d049bb133111eb)
    x ^= x >> np.uint64(31)
    return x


def device_twin_counts(prof, cell_begin, n_cells, lbar, depth_sigma=0.35, r_disp=5.0, seed=20251019, chunk=256):
    """Rows [cell_begin, cell_begin + n_cells) of the matrix ab_synth_nb_* generates from `prof`
    (float32 [clusters, genes], alona_b200.synth.make_profiles).  Returns scipy CSR int32 (cells x genes)."""
    prof = np.ascontiguousarray(prof, dtype=np.float32)
    n_clusters, n_genes = prof.shape
    old = np.seterr(over='ignore')
    try:
        cells = np.arange(cell_begin, cell_begin + n_cells, dtype=np.uint64)
        h = _mix64(np.uint64(seed) * np.uint64(0x9e3779b97f4a7c15) + cells * np.uint64(0xd1b54a32d192ed03) + np.uint64(0x1234567))
        key = _mix64(h ^ np.uint64(0xabcdef12345))
        cluster = (h % np.uint64(n_clusters)).astype(np.int64)
        h2 = _mix64(h + np.uint64(0x632be59bd9b4e019))
        u1 = ((h2 >> np.uint64(11)).astype(np.float64) + 0.5) * (1.0 / 9007199254740992.0)
        u2 = ((_mix64(h2) >> np.uint64(11)).astype(np.float64) + 0.5) * (1.0 / 9007199254740992.0)
        z = np.sqrt(-2.0 * np.log(u1)) * np.cos(2.0 * np.pi * u2)
        depth = np.exp(np.log(lbar) + depth_sigma * z).astype(np.float32)
        r = np.float32(r_disp)
        genes = np.arange(n_genes, dtype=np.uint64) * np.uint64(0x9e3779b97f4a7c15)
        rows = []
        for a in range(0, n_cells, chunk):
            b = min(n_cells, a + chunk)
            hh = _mix64(key[a:b, None] + genes[None, :])
            u = (hh >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)
            mu = depth[a:b, None] * prof[cluster[a:b]]
            lg = np.log1p(mu / r, dtype=np.float32)
            qnz = -np.expm1(-r * lg, dtype=np.float32)
            cnt = np.zeros(mu.shape, dtype=np.int32)
            ii, jj = np.nonzero(u < qnz.astype(np.float64))
            uu = u[ii, jj]
            m = mu[ii, jj]
            p = np.exp(-r * lg[ii, jj], dtype=np.float32)
            ratio = (m / (m + r)).astype(np.float32)
            cdf = np.zeros(uu.shape, dtype=np.float64)
            c = np.zeros(uu.shape, dtype=np.int32)
            live = np.ones(uu.shape, dtype=bool)
            step = 0
            while live.any() and step < 100000:
                idx = np.nonzero(live)[0]
                pp = p[idx]
                pp = pp * (np.float32(step) + r)
                pp = pp / np.float32(step + 1)
                pp = pp * ratio[idx]
                p[idx] = pp
                c[idx] = step + 1
                cdf[idx] += pp.astype(np.float64)
                live[idx] = ~(uu[idx] < cdf[idx])
                step += 1
            cnt[ii, jj] = c
            empty = np.nonzero(cnt.sum(axis=1) == 0)[0]
            for e in empty:                                  # every cell must have S_c >= 1
                cnt[e, int(key[a + e] % np.uint64(n_genes))] = 1
            rows.append(sp.csr_matrix(cnt))
        out = sp.vstack(rows).tocsr()
        out.sort_indices()
        return out.astype(np.int32)
    finally:
        np.seterr(**old)
