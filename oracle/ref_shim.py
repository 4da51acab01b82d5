"""TEST INFRASTRUCTURE — drives the UNMODIFIED reference (alona v0.3.5, /root/reference).

Only usable in the development container: /root/reference does not exist on the GPU box.
Nothing under tests/ (-m gpu), smoke() or bench.py imports this; it is used by
oracle/make_golden.py to mint tests/golden/ fixtures and by the "not gpu" tests that pin
the oracle port against the live reference when it is present.

Shims:
  * empty stub modules for the seven packages the reference imports at module top but never
    calls on the hot path (matplotlib, seaborn, umap, leidenalg, igraph, patsy, statsmodels);
  * numpy.float = float  (alona/irlbpy/irlb.py:85 uses the alias removed in numpy 1.24) and numpy.asfarray
    (alona/stats.py:15, removed in numpy 2.0; only the Brennecke selector reaches it).
No reference source is edited or copied.
"""
import os
import sys
import types
import tempfile

import numpy as np
import pandas as pd

REFERENCE_ROOT = os.environ.get('ALONA_REFERENCE_ROOT', '/root/reference')


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, 'alona'))


class _Any:
    def __getattr__(self, k):
        return _Any()

    def __call__(self, *a, **k):
        return _Any()


_loaded = None


def load():
    """Imports alona.clustering from the reference tree with the shims applied."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError('reference tree not present at %s' % REFERENCE_ROOT)
    for name in ['matplotlib', 'matplotlib.pyplot', 'matplotlib.patches',
                 'matplotlib.transforms', 'seaborn', 'umap', 'leidenalg', 'igraph', 'patsy',
                 'statsmodels', 'statsmodels.api']:
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__getattr__ = lambda k: _Any()
            sys.modules[name] = m
    if not hasattr(np, 'float'):
        np.float = float
    if not hasattr(np, 'asfarray'):             # alona/stats.py:15 (p_adjust_bh), removed in numpy 2.0
        np.asfarray = lambda a, dtype=np.float64: np.asarray(a, dtype=dtype)
    sys.dont_write_bytecode = True
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import alona.clustering as C
    import alona.irlbpy as I
    _loaded = (C, I)
    return _loaded


def run_reference(counts_df, hvg_n, pca_n, nn_k, prune_snn=0.067, seed=42, workdir=None,
                  stages=('normalize', 'hvg', 'pca', 'knn', 'snn')):
    """Runs normalize -> find_variable_genes -> PCA -> knn -> snn of the reference.

    counts_df: pandas DataFrame genes x cells, int64 (the layout alona/cell.py:381-383 loads).
    Returns a dict of every intermediate the stages expose.
    """
    import warnings
    C, _ = load()
    own = workdir is None
    if own:
        tmp = tempfile.TemporaryDirectory()
        workdir = tmp.name
    os.makedirs(os.path.join(workdir, 'csvs'), exist_ok=True)
    snn_csv = os.path.join(workdir, 'csvs', 'snn_graph.csv')
    if os.path.exists(snn_csv):
        os.remove(snn_csv)          # cache is not keyed on k/prune (clustering.py:197-201)
    o = C.AlonaClustering()
    o.params = {'qc_auto': False, 'hvg_method': 'seurat', 'hvg_n': hvg_n, 'pca': 'irlb',
                'pca_n': pca_n, 'seed': seed, 'nn_k': nn_k, 'prune_snn': prune_snn,
                'output_directory': workdir, 'loglevel': 'regular'}
    o.low_quality_cells = []
    o.data_ERCC = pd.DataFrame()
    o.anno = None
    out = {}
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        o.data_norm = o.normalize(counts_df, '', 'raw', False)
        out['data_norm'] = o.data_norm
        if 'hvg' in stages:
            o.find_variable_genes()
            out['hvg'] = np.array(o.hvg)
        if 'pca' in stages:
            # the PCA method does not expose steps/nmult; re-run lanczos the same way
            o.PCA(os.path.join(workdir, 'csvs', 'pca.csv'))
            out['pca_components'] = o.pca_components.values.copy()
            sliced = o.data_norm[o.data_norm.index.isin(o.hvg)]
            import alona.irlbpy
            lanc = alona.irlbpy.lanczos(sliced, nval=pca_n, maxit=1000, seed=seed)
            assert np.array_equal(np.dot(lanc.V, np.diag(lanc.s)), out['pca_components'])
            out['lanczos'] = {'U': lanc.U, 's': lanc.s, 'V': lanc.V, 'steps': lanc.steps,
                              'nmult': lanc.nmult}
        if 'knn' in stages:
            o.knn(nn_k)
            out['nn_idx'] = o.nn_idx.copy()
        if 'snn' in stages:
            o.snn(nn_k, prune_snn)
            out['snn_graph'] = o.snn_graph.values.copy()
    if own:
        tmp.cleanup()
    return out


def run_reference_filters(counts_df, minreads, minexpgenes, remove_mito='no'):
    """remove_empty -> remove_mito -> simple_filters of the reference (alona/cell.py:64-94, 141-173) on a
    genes x cells int64 DataFrame; returns the surviving frame after each step."""
    import warnings
    C, _ = load()
    o = C.AlonaClustering()
    o.params = {'dataformat': 'raw', 'minreads': minreads, 'minexpgenes': float(minexpgenes),
                'remove_mito': remove_mito, 'loglevel': 'regular'}
    o.data = counts_df.copy()
    out = {}
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        o.remove_empty()
        out['after_remove_empty'] = o.data.copy()
        o.remove_mito()
        out['after_remove_mito'] = o.data.copy()
        o.simple_filters()
        out['after_simple_filters'] = o.data.copy()
    return out


def reference_knn(emb, k):
    """alona/clustering.py:179-188 on a given embedding (1-based indices)."""
    C, _ = load()
    o = C.AlonaClustering()
    o.pca_components = pd.DataFrame(emb)
    o.knn(k)
    return o.nn_idx


def reference_snn(nn_idx, k, prune_snn, workdir=None):
    """alona/clustering.py:190-240 on a given 1-based nn_idx; returns the edge array."""
    C, _ = load()
    own = workdir is None
    if own:
        tmp = tempfile.TemporaryDirectory()
        workdir = tmp.name
    os.makedirs(os.path.join(workdir, 'csvs'), exist_ok=True)
    p = os.path.join(workdir, 'csvs', 'snn_graph.csv')
    if os.path.exists(p):
        os.remove(p)
    o = C.AlonaClustering()
    o.params = {'output_directory': workdir}
    o.nn_idx = nn_idx
    o.snn(k, prune_snn)
    g = o.snn_graph.values.copy()
    if own:
        tmp.cleanup()
    return g
