"""TEST INFRASTRUCTURE — mints tests/golden/*.npz by running the UNMODIFIED reference.

Run in the development container only (needs /root/reference):
    PYTHONDONTWRITEBYTECODE=1 python -m oracle.make_golden
The reference ships no golden vectors, so these fixtures — outputs of the
reference itself on seeded synthetic inputs — are what pins the oracle and the CUDA path.
Each fixture stores the input (cell-major CSR counts) and every stage output of
alona/cell.py:224-255, alona/hvg.py:133-167, alona/clustering.py:82-113,179-240.
"""
import os
import sys
import warnings

import numpy as np
import pandas as pd

from . import synth_np, ref_shim, alona_port

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden')

PIPELINE_CASES = {
    # name: (cells, genes, lbar, clusters, data_seed, hvg_n, pca_n, k, prune, seed)
    'pipe_small': (600, 3000, 3000.0, 8, 1, 300, 12, 8, 0.067, 42),
    'pipe_medium': (1500, 6000, 1500.0, 12, 7, 500, 30, 10, 0.067, 42),
    'pipe_k20': (1200, 4000, 1200.0, 10, 11, 400, 20, 20, 0.067, 3),
}


def _counts_df(m):
    genes = ['g%d' % i for i in range(m.shape[1])]
    cells = ['c%d' % i for i in range(m.shape[0])]
    return pd.DataFrame(m.T.toarray().astype(np.int64), index=genes, columns=cells)


def make_pipeline_case(name, spec):
    n, g, lbar, ncl, dseed, hvg_n, pca_n, k, prune, seed = spec
    m, _ = synth_np.nb_counts(n, g, lbar=lbar, n_clusters=ncl, seed=dseed)
    df = _counts_df(m)
    ref = ref_shim.run_reference(df, hvg_n=hvg_n, pca_n=pca_n, nn_k=k, prune_snn=prune, seed=seed)
    dn = ref['data_norm']
    gene_pos = {gname: i for i, gname in enumerate(df.index)}
    hvg_ranked = np.array([gene_pos[x] for x in ref['hvg']], dtype=np.int32)
    _, counts = alona_port.snn(ref['nn_idx'], k, prune, return_counts=True)
    np.random.seed(seed)
    v0 = np.random.randn(m.shape[0])
    np.savez_compressed(
        os.path.join(OUT, name + '.npz'),
        indptr=m.indptr.astype(np.int64), indices=m.indices.astype(np.int32),
        data=m.data.astype(np.int32), shape=np.array(m.shape, dtype=np.int64),
        hvg_n=hvg_n, pca_n=pca_n, k=k, prune=prune, seed=seed,
        libsize=np.asarray(df.sum(axis=0).values, dtype=np.int64),
        gene_mean=dn.mean(axis=1).values, gene_var=dn.var(axis=1).values,
        hvg_ranked=hvg_ranked,
        pca_components=ref['pca_components'], s=ref['lanczos']['s'],
        U=ref['lanczos']['U'],
        steps=ref['lanczos']['steps'], nmult=ref['lanczos']['nmult'], v0=v0,
        nn_idx=ref['nn_idx'].astype(np.int32),
        snn_graph=ref['snn_graph'].astype(np.int32), snn_overlap=counts.astype(np.int16))
    print(name, m.shape, 'nnz', m.nnz, 'steps', ref['lanczos']['steps'], 'nmult',
          ref['lanczos']['nmult'], 'edges', ref['snn_graph'].shape[0])


# BASELINE.json configs[0] (C1: 2,700 x 32,738, hvg 1000, 40 PCs, k 10) and C1-shaped variants that drive the
# IRLBA work dimensions the bench shapes run (m_b = nval + 20: 60 / 70 / 95 / 120).  One input
# matrix (stored once, compactly) + one output file per variant.
C1_INPUT = dict(cells=2700, genes=32738, lbar=1061.0, clusters=30, data_seed=20251019)
C1_VARIANTS = {
    # name: (hvg_n, pca_n, k, prune, seed)
    'c1_full': (1000, 40, 10, 0.067, 42),            # BASELINE.json configs[0]; m_b = 60
    'c1_h2000_d50_k20': (2000, 50, 20, 0.067, 42),   # C2/C3 parameters; m_b = 70
    'c1_h2000_d75_k20': (2000, 75, 20, 0.067, 42),   # m_b = 95
    'c1_h3000_d100_k30': (3000, 100, 30, 0.067, 42),  # C5 parameters; m_b = 120
}


def make_c1_cases():
    spec = C1_INPUT
    m, _ = synth_np.nb_counts(spec['cells'], spec['genes'], lbar=spec['lbar'], n_clusters=spec['clusters'],
                              seed=spec['data_seed'])
    assert m.shape[1] < 65536 and m.data.max() < 65536
    np.savez_compressed(os.path.join(OUT, 'c1_input.npz'), indptr=m.indptr.astype(np.int64),
                        indices=m.indices.astype(np.uint16), data=m.data.astype(np.uint16),
                        shape=np.array(m.shape, dtype=np.int64))
    print('c1_input', m.shape, 'nnz', m.nnz, 'genes detected per cell', m.nnz / m.shape[0])
    df = _counts_df(m)
    gene_pos = {gname: i for i, gname in enumerate(df.index)}
    for name, (hvg_n, pca_n, k, prune, seed) in C1_VARIANTS.items():
        ref = ref_shim.run_reference(df, hvg_n=hvg_n, pca_n=pca_n, nn_k=k, prune_snn=prune, seed=seed)
        dn = ref['data_norm']
        hvg_ranked = np.array([gene_pos[x] for x in ref['hvg']], dtype=np.int32)
        _, counts = alona_port.snn(ref['nn_idx'], k, prune, return_counts=True)
        np.random.seed(seed)
        v0 = np.random.randn(m.shape[0])
        np.savez_compressed(
            os.path.join(OUT, name + '.npz'), input='c1_input',
            hvg_n=hvg_n, pca_n=pca_n, k=k, prune=prune, seed=seed,
            libsize=np.asarray(df.sum(axis=0).values, dtype=np.int64),
            gene_mean=dn.mean(axis=1).values, gene_var=dn.var(axis=1).values, hvg_ranked=hvg_ranked,
            pca_components=ref['pca_components'], s=ref['lanczos']['s'],
            steps=ref['lanczos']['steps'], nmult=ref['lanczos']['nmult'], v0=v0,
            nn_idx=ref['nn_idx'].astype(np.int32), snn_graph=ref['snn_graph'].astype(np.int32),
            snn_overlap=counts.astype(np.int16))
        print(name, 'steps', ref['lanczos']['steps'], 'nmult', ref['lanczos']['nmult'], 'edges',
              ref['snn_graph'].shape[0])


def make_pca_regular_case(name, n, g, lbar, seed, hvg_n, pca_n):
    """N4 fixture: AlonaClustering.PCA with --pca regular (alona/clustering.py:114-125: gene-centred exact SVD)."""
    C, _ = ref_shim.load()
    import tempfile
    m, _ = synth_np.nb_counts(n, g, lbar=lbar, n_clusters=6, seed=seed)
    df = _counts_df(m)
    o = C.AlonaClustering()
    with tempfile.TemporaryDirectory() as wd:
        os.makedirs(wd + '/csvs')
        o.params = {'qc_auto': False, 'hvg_method': 'seurat', 'hvg_n': hvg_n, 'pca': 'regular', 'pca_n': pca_n, 'seed': 42,
                    'output_directory': wd, 'loglevel': 'regular'}
        o.low_quality_cells = []
        o.data_ERCC = pd.DataFrame()
        o.anno = None
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            o.data_norm = o.normalize(df, '', 'raw', False)
            o.find_variable_genes()
            o.PCA(wd + '/csvs/pca.csv')
    gene_pos = {gname: i for i, gname in enumerate(df.index)}
    np.savez_compressed(os.path.join(OUT, name + '.npz'),
                        indptr=m.indptr.astype(np.int64), indices=m.indices.astype(np.int32), data=m.data.astype(np.int32),
                        shape=np.array(m.shape, dtype=np.int64), hvg_n=hvg_n, pca_n=pca_n,
                        hvg_ranked=np.array([gene_pos[x] for x in o.hvg], dtype=np.int32),
                        pca_components=np.asarray(o.pca_components.values))
    print(name, m.shape, 'pca regular', o.pca_components.shape)


def make_hvg_methods_case(name, n, g, lbar, seed, hvg_n, n_ercc=48):
    """N4 fixtures: AlonaHighlyVariableGenes.find() of the live reference with hvg_method Brennecke2013
    (alona/hvg.py:65-131) and scran (:169-225) on one input; the ERCC spikes are extra count rows normalised on their
    own, as AlonaCell.load_data does (alona/cell.py:397,410)."""
    C, _ = ref_shim.load()
    import alona.hvg as H
    m, _ = synth_np.nb_counts(n, g, lbar=lbar, n_clusters=6, seed=seed)
    rng = np.random.RandomState(seed + 1000)
    # spikes: the same amount in every cell, abundances over four decades, Poisson noise
    conc = np.exp(rng.uniform(np.log(0.3), np.log(300.0), n_ercc))
    ercc = rng.poisson(conc[None, :] * np.exp(rng.normal(0.0, 0.2, (m.shape[0], 1)))).astype(np.int64)
    ercc[:, ercc.sum(axis=0) == 0] = 1
    df = _counts_df(m)
    df_ercc = pd.DataFrame(ercc.T, index=['ERCC-%05d' % i for i in range(n_ercc)], columns=df.columns)
    o = C.AlonaClustering()
    o.params = {'qc_auto': False, 'loglevel': 'regular'}
    o.low_quality_cells = []
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        dn = o.normalize(df, '', 'raw', False)
        dn_ercc = o.normalize(df_ercc, '', 'raw', False)
        gene_pos = {gname: i for i, gname in enumerate(df.index)}
        out = {}
        for method in ('Brennecke2013', 'scran'):
            f = H.AlonaHighlyVariableGenes(hvg_method=method, hvg_n=hvg_n, data_norm=dn, data_ERCC=dn_ercc)
            out[method] = np.array([gene_pos[x] for x in f.find()], dtype=np.int32)
    import scipy.sparse as sp
    e = sp.csr_matrix(ercc.astype(np.int32))
    np.savez_compressed(os.path.join(OUT, name + '.npz'),
                        indptr=m.indptr.astype(np.int64), indices=m.indices.astype(np.int32), data=m.data.astype(np.int32),
                        shape=np.array(m.shape, dtype=np.int64), ercc_indptr=e.indptr.astype(np.int64),
                        ercc_indices=e.indices.astype(np.int32), ercc_data=e.data.astype(np.int32),
                        ercc_shape=np.array(e.shape, dtype=np.int64), hvg_n=hvg_n,
                        hvg_brennecke=out['Brennecke2013'], hvg_scran=out['scran'])
    print(name, m.shape, 'brennecke', out['Brennecke2013'].size, 'scran', out['scran'].size)


def make_knn_case(name, n, d, k, seed, dup=0, quantise=None):
    emb = synth_np.gaussian_mixture_embedding(n, d=d, seed=seed)
    if quantise:
        emb = np.round(emb * quantise) / quantise      # many exactly-equal distances
    if dup:
        emb[-dup:] = emb[:dup]                          # exact duplicate points
    nn = ref_shim.reference_knn(emb, k)
    edges = ref_shim.reference_snn(nn, k, 0.067)
    _, counts = alona_port.snn(nn, k, 0.067, return_counts=True)
    np.savez_compressed(os.path.join(OUT, name + '.npz'), emb=emb, k=k,
                        nn_idx=nn.astype(np.int32), snn_graph=edges.astype(np.int32),
                        snn_overlap=counts.astype(np.int16), prune=0.067)
    print(name, emb.shape, 'k', k, 'edges', edges.shape[0])


def make_qc_case(name, n, g, lbar, seed, minreads, thresholds=(0.05, 4.0)):
    """N3 fixtures: remove_empty / remove_mito / simple_filters of the live reference (alona/cell.py:64-94,
    141-173) and the per-cell QC expressions of cell.py:306-312 evaluated with pandas on the same frame."""
    m, _ = synth_np.nb_counts(n, g, lbar=lbar, n_clusters=4, seed=seed, drop_empty=False)
    m = m.tolil()
    m[5, :] = 0
    m[:, 7] = 0
    m[:, 9] = 0
    m = m.tocsr()
    m.eliminate_zeros()
    genes = np.array(['g%d' % i for i in range(g)], dtype=object)
    genes[::17] = np.array(['mt-x%d' % i for i in range(len(genes[::17]))], dtype=object)     # '^mt-' genes (cell.py:88)
    genes[3::29] = np.array(['MT-y%d' % i for i in range(len(genes[3::29]))], dtype=object)   # case=False
    cells = np.array(['c%d' % i for i in range(n)], dtype=object)
    df = pd.DataFrame(m.T.toarray().astype(np.int64), index=genes, columns=cells)
    gpos = {x: i for i, x in enumerate(genes)}
    cpos = {x: i for i, x in enumerate(cells)}
    rec = dict(indptr=m.indptr.astype(np.int64), indices=m.indices.astype(np.int32), data=m.data.astype(np.int32),
               shape=np.array(m.shape, dtype=np.int64), genes=genes.astype(str), minreads=minreads,
               thresholds=np.array(thresholds, dtype=np.float64))
    # cell.py:306-312 (the MinCovDet step that consumes them is outside the scope)
    reads_per_cell = df.sum(axis=0)
    rec['reads_per_cell'] = reads_per_cell.values.astype(np.int64)
    rec['no_genes_det'] = np.sum(df > 0, axis=0).values.astype(np.int64)
    data_mt = df[df.index.str.contains('^mt-', regex=True, case=False)]
    with np.errstate(invalid='ignore', divide='ignore'):
        rec['perc_mt'] = (data_mt.sum(axis=0) / reads_per_cell * 100).values
    for remove_mito in ('no', 'yes'):
        for ti, thr in enumerate(thresholds):
            out = ref_shim.run_reference_filters(df, minreads, thr, remove_mito=remove_mito)
            for step, frame in out.items():
                key = '%s_mito%s_t%d' % (step, remove_mito, ti)
                rec[key + '_genes'] = np.array([gpos[x] for x in frame.index], dtype=np.int32)
                rec[key + '_cells'] = np.array([cpos[x] for x in frame.columns], dtype=np.int32)
    np.savez_compressed(os.path.join(OUT, name + '.npz'), **rec)
    print(name, m.shape, 'nnz', m.nnz, 'kept', rec['after_simple_filters_mitono_t0_cells'].size, 'cells',
          rec['after_simple_filters_mitono_t0_genes'].size, 'genes')


def make_cluster_exp_case(name, n, g, lbar, ncl, seed):
    """N2 fixtures: the reference's own normalize() output (alona/cell.py:241-243) reduced per cluster with the
    numpy functions mean_exp / median_exp hand to groupby (alona/celltypes.py:49,61; `groupby(axis=1)` itself is
    gone from the installed pandas), and the residual variance of find_markers.py:96-97 with coef = cluster means."""
    m, labels = synth_np.nb_counts(n, g, lbar=lbar, n_clusters=ncl, seed=seed)
    labels = np.asarray(labels)[:m.shape[0]].astype(np.int64)
    labels[:3] = ncl                                    # a tiny extra cluster (odd size)
    df = _counts_df(m)
    ref = ref_shim.run_reference(df, hvg_n=10, pca_n=2, nn_k=2, stages=('normalize',))
    dn = ref['data_norm']                               # genes x cells float64 DataFrame
    ids = np.unique(labels)
    mean = np.stack([np.mean(dn.values[:, labels == c], axis=1) for c in ids])
    med = np.stack([np.median(dn.values[:, labels == c], axis=1) for c in ids])
    fitted = mean[np.searchsorted(ids, labels)].T       # dm_full.dot(coef), genes x cells
    sigma2 = ((dn.values - fitted) ** 2).sum(axis=1) / (dn.shape[1] - len(ids))
    np.savez_compressed(os.path.join(OUT, name + '.npz'),
                        indptr=m.indptr.astype(np.int64), indices=m.indices.astype(np.int32),
                        data=m.data.astype(np.int32), shape=np.array(m.shape, dtype=np.int64),
                        labels=labels, cluster_ids=ids, mean=mean, median=med, sigma2=sigma2)
    print(name, m.shape, 'clusters', len(ids), 'nonzero medians', int((med > 0).sum()), 'of', med.size)


def main():
    if not ref_shim.available():
        sys.exit('reference tree not present; fixtures can only be minted in the dev container')
    os.makedirs(OUT, exist_ok=True)
    warnings.simplefilter('ignore')
    if '--pcareg' in sys.argv:
        make_pca_regular_case('pcareg_1200x3000', 1200, 3000, 1500.0, 31, 400, 20)
        return
    if '--hvg' in sys.argv:                     # only the N4 selector fixtures
        make_hvg_methods_case('hvgm_900x2500', 900, 2500, 2500.0, 21, 300)
        make_hvg_methods_case('hvgm_1500x4000', 1500, 4000, 900.0, 23, 5000)
        return
    if '--c1' in sys.argv:                      # only the C1 family (minutes of the live reference)
        make_c1_cases()
        return
    for name, spec in PIPELINE_CASES.items():
        make_pipeline_case(name, spec)
    make_knn_case('knn_mix_3000x50_k20', 3000, 50, 20, seed=2)
    make_knn_case('knn_mix_2000x50_k10', 2000, 50, 10, seed=5)
    make_knn_case('knn_mix_1500x100_k30', 1500, 100, 30, seed=9)
    make_knn_case('knn_mix_1000x50_k100', 1000, 50, 100, seed=4)
    make_knn_case('knn_dup_800x50_k10', 800, 50, 10, seed=6, dup=5)
    make_qc_case('qc_square_400', 400, 400, 8.0, 3, 3)
    make_cluster_exp_case('clexp_700x900', 700, 900, 700.0, 5, 13)
    make_cluster_exp_case('clexp_401x500', 401, 500, 150.0, 3, 17)
    make_qc_case('qc_rect_700x300', 700, 300, 10.0, 5, 4)
    make_c1_cases()
    make_hvg_methods_case('hvgm_900x2500', 900, 2500, 2500.0, 21, 300)
    make_hvg_methods_case('hvgm_1500x4000', 1500, 4000, 900.0, 23, 5000)
    make_pca_regular_case('pcareg_1200x3000', 1200, 3000, 1500.0, 31, 400, 20)


if __name__ == '__main__':
    main()
