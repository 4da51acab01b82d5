"""GPU tier: the reference-facing Python API (same class / method / attribute names as
alona/cell.py, alona/hvg.py, alona/clustering.py) against fixtures minted from the reference."""
import os

import numpy as np
import pandas as pd
import pytest

from conftest import load_golden, golden_csr, PIPE_CASES
from test_gpu_stages import check_free_running_against_reference
from oracle import restate as R

pytestmark = pytest.mark.gpu


def _params(g, wd):
    return {'qc_auto': False, 'hvg_method': 'seurat', 'hvg_n': int(g['hvg_n']), 'pca': 'irlb',
            'pca_n': int(g['pca_n']), 'seed': int(g['seed']), 'nn_k': int(g['k']),
            'prune_snn': float(g['prune']), 'output_directory': wd, 'leiden_res': 0.8}


@pytest.mark.parametrize('case', PIPE_CASES + ['c1_full'])
def test_alona_clustering_dropin(case, tmp_path):
    from alona_b200.clustering import AlonaClustering
    g = load_golden(case)
    m = golden_csr(g)
    genes = ['g%d' % i for i in range(m.shape[1])]
    cells = ['c%d' % i for i in range(m.shape[0])]
    df = pd.DataFrame(m.T.toarray().astype(np.int64), index=genes, columns=cells)   # reference layout
    wd = str(tmp_path)
    os.makedirs(wd + '/csvs', exist_ok=True)
    o = AlonaClustering()
    o.params = _params(g, wd)
    o.low_quality_cells = []
    o.data_ERCC = pd.DataFrame()
    o.data_norm = o.normalize(df, '', 'raw', False)
    assert o.data_norm.shape == (m.shape[1], m.shape[0])
    o.find_variable_genes()
    assert set(o.hvg.tolist()) == set(np.array(genes)[g['hvg_ranked']].tolist())
    o.PCA(wd + '/csvs/pca.csv')
    assert list(o.pca_components.index) == cells
    err = R.column_rel_err(g['pca_components'], o.pca_components.values)
    assert err.max() < 1e-9
    o.knn(int(g['k']))
    assert o.nn_idx.dtype == np.int64 and o.nn_idx.min() == 1                       # 1-based like the reference
    o.snn(int(g['k']), float(g['prune']))
    assert list(o.snn_graph.columns) == ['source_node', 'target_node']
    # no escape: the run's own chain is exact, rows that differ from the reference's must be near ties there
    import torch
    out = {'hvg_ranked': o._hvg_finder.ranked_idx, 'irlb': {'steps': o.report['pca']['steps'], 'nmult': o.report['pca']['nmult']},
           'emb_all': o._emb_dev, 'knn_all': o._knn_dev,
           'snn': {'src': torch.from_numpy(o.snn_graph.values[:, 0].copy()), 'dst': torch.from_numpy(o.snn_graph.values[:, 1].copy()),
                   'overlap': torch.from_numpy(o.snn_overlap)}}
    diff_rows = check_free_running_against_reference(out, g)
    assert diff_rows.size <= max(1, m.shape[0] // 1000)
    # artefacts (alona/constants.py:24-30)
    for rel in ('/csvs/highly_variable_genes.tsv', '/csvs/pca.csv', '/csvs/snn_graph.csv'):
        assert os.path.exists(wd + rel), rel
    back = pd.read_csv(wd + '/csvs/snn_graph.csv', header=None).values
    assert np.array_equal(back, o.snn_graph.values)
    # the device-formatted file is byte-identical to what the reference writes (clustering.py:237-238)
    import io
    ref_bytes = io.StringIO()
    o.snn_graph.to_csv(ref_bytes, header=None, index=None)
    assert open(wd + '/csvs/snn_graph.csv', 'rb').read() == ref_bytes.getvalue().encode()
    hv = pd.read_csv(wd + '/csvs/highly_variable_genes.tsv', header=None, sep='\t')[0].values
    assert np.array_equal(hv, o.hvg)
    # cache rule: an existing snn_graph.csv is loaded, not recomputed (clustering.py:197-201)
    o2 = AlonaClustering(); o2.params = o.params; o2.nn_idx = None
    o2.snn(int(g['k']), float(g['prune']))
    assert np.array_equal(o2.snn_graph.values, back)


def test_data_norm_dataframe_matches_reference_values():
    from alona_b200.clustering import AlonaClustering
    g = load_golden('pipe_small')
    m = golden_csr(g)
    o = AlonaClustering()
    o.params = {'qc_auto': False}
    dn = o.normalize(pd.DataFrame(m.T.toarray().astype(np.int64)), '', 'raw', False)
    dense = dn.to_dataframe().values                       # genes x cells
    ref = np.zeros(dense.shape)
    rows = np.repeat(np.arange(m.shape[0]), np.diff(m.indptr))
    ref[m.indices, rows] = R.normalized_values(m)
    assert np.array_equal(dense == 0, ref == 0)
    assert np.max(np.abs(dense - ref)) < 1e-14
    assert np.allclose(dn.mean(axis=1).values, g['gene_mean'], rtol=1e-12, atol=0)


def test_out_of_scope_options_exit_like_log_error():
    from alona_b200.clustering import AlonaClustering
    from alona_b200.hvg import AlonaHighlyVariableGenes
    o = AlonaClustering()
    o.params = {'qc_auto': False}
    with pytest.raises(SystemExit):
        o.normalize(pd.DataFrame(np.ones((3, 3), dtype=np.int64)), '', 'rpkm', False)
    with pytest.raises(SystemExit):
        AlonaHighlyVariableGenes('scran', 10, None, None).find()             # no ERCC spikes: alona/hvg.py:187-190
    with pytest.raises(SystemExit):
        AlonaHighlyVariableGenes('Chen2016', 10, None, None).find()          # outside SURVEY.md 8(f) N4
    with pytest.raises(SystemExit):
        AlonaHighlyVariableGenes('nonsense', 10, None, None).find()


def test_edges_csv_bytes_match_pandas():
    """ab_edges_csv_*: every digit count (1..10 digits), empty list, one edge."""
    import io
    import torch
    from alona_b200.engine import get_engine
    eng = get_engine()
    rng = np.random.RandomState(1)
    vals = np.concatenate([[0, 9, 10, 99, 100, 999999, 1000000, 2147483647],
                           (10 ** rng.uniform(0, 9.3, 5000)).astype(np.int64)]).astype(np.int32)
    for n in (0, 1, len(vals)):
        src = vals[:n]
        dst = vals[::-1][:n].copy()
        buf = eng.edges_csv_bytes(torch.from_numpy(src).to(eng.device), torch.from_numpy(dst).to(eng.device))
        ref = io.StringIO()
        pd.DataFrame({'source_node': src, 'target_node': dst}).to_csv(ref, header=None, index=None)
        assert bytes(buf.numpy().tobytes()) == ref.getvalue().encode()


# ---- orchestration: cluster() / analysis() (alona/clustering.py:295-311, alona/cell.py:418-435) --------------------
class _FakeGraph:
    """Stand-in for igraph.Graph (python-igraph is not installed here): records what leiden() hands over."""
    last = None

    def __init__(self):
        self.n = 0
        self.vs = {}
        self.edges = None
        _FakeGraph.last = self

    def add_vertices(self, n):
        self.n = n

    def add_edges(self, e):
        self.edges = e


class _FakePartition:
    def __init__(self, membership):
        self.membership = membership


def _install_fake_leiden(monkeypatch, labels_of):
    import sys
    import types
    ig = types.ModuleType('igraph')
    ig.Graph = _FakeGraph
    la = types.ModuleType('leidenalg')
    la.RBERVertexPartition = 'RBER'
    la.ModularityVertexPartition = 'MOD'
    calls = {}

    def find_partition(g, part, n_iterations, resolution_parameter, seed):
        calls.update(part=part, n_iterations=n_iterations, res=resolution_parameter, seed=seed)
        return _FakePartition(labels_of(g))
    la.find_partition = find_partition
    monkeypatch.setitem(sys.modules, 'igraph', ig)
    monkeypatch.setitem(sys.modules, 'leidenalg', la)
    return calls


def _new_pred(g, wd, **extra):
    from alona_b200.celltypes import AlonaCellTypePred
    o = AlonaCellTypePred()
    o.set_params(dict(_params(g, wd), **extra))
    o.low_quality_cells = []
    import pandas as pd
    o.data_ERCC = pd.DataFrame()
    return o


def test_analysis_then_cluster_then_mean_exp(tmp_path, monkeypatch):
    """analysis() = find_variable_genes -> PCA -> cluster() = knn -> snn -> leiden -> leiden_prep, then the
    per-cluster table, with no attribute set by hand (clusters_targets used to be missing)."""
    g = load_golden('pipe_k20')
    m = golden_csr(g)
    n = m.shape[0]
    labels = (np.arange(n) % 7 == 0).astype(int) + (np.arange(n) % 400 == 1).astype(int) * 2   # sizes: big, ~n/7, 3
    calls = _install_fake_leiden(monkeypatch, lambda graph: labels.tolist())
    wd = str(tmp_path)
    o = _new_pred(g, wd, ignore_small_clusters=10)
    df = pd.DataFrame(m.T.toarray().astype(np.int64), index=['g%d' % i for i in range(m.shape[1])],
                      columns=['c%d' % i for i in range(n)])
    o.data_norm = o.normalize(df, '', 'raw', False)
    o.analysis()
    # the hand-off: one [E, 2] integer array, edges identical to the reference's, no tuple list
    fg = _FakeGraph.last
    assert isinstance(fg.edges, np.ndarray) and fg.edges.shape[1] == 2 and fg.edges.flags['C_CONTIGUOUS']
    assert fg.n == n and fg.vs['name'] == list(range(1, n + 1))
    assert np.array_equal(fg.edges, o.snn_graph.values)
    assert calls == {'part': 'RBER', 'n_iterations': 10, 'res': 0.8, 'seed': int(g['seed'])}
    # leiden_prep (clustering.py:242-260)
    ids, counts = np.unique(labels, return_counts=True)
    assert o.n_clusters == int((counts > 10).sum())
    assert np.array_equal(o.clusters_targets, ids[counts > 10])
    cl_file = pd.read_csv(wd + '/csvs/clusters_leiden.csv', header=None, index_col=0)
    assert list(cl_file.index) == list(df.columns) and np.array_equal(cl_file[1].values, labels)
    assert len(o.cluster_colors) == len(o.clusters_targets)
    # downstream consumer works straight away, small cluster dropped
    tab = o.mean_exp()
    assert list(tab.columns) == list(o.clusters_targets) and tab.shape[0] == m.shape[1]
    assert os.path.exists(wd + '/csvs/mean_exp.tsv')
    # cache rule of analysis(): with pca.csv and the embedding file present, HVG/PCA are skipped (cell.py:421-431)
    open(wd + '/csvs/embeddings_tSNE.csv', 'w').write('x\n')
    o2 = _new_pred(g, wd)
    o2.data_norm = o.data_norm
    o2.find_variable_genes = lambda: (_ for _ in ()).throw(AssertionError('HVG must be skipped'))
    os.remove(wd + '/csvs/snn_graph.csv')
    o2.analysis()
    assert np.allclose(o2.pca_components.values, o.pca_components.values, rtol=1e-12, atol=0)
    assert np.array_equal(o2.snn_graph.values, o.snn_graph.values)


def test_cluster_with_custom_clustering_skips_the_graph(tmp_path):
    g = load_golden('pipe_small')
    m = golden_csr(g)
    n = m.shape[0]
    o = _new_pred(g, str(tmp_path))
    cells = ['c%d' % i for i in range(n)]
    o.data_norm = o.normalize(pd.DataFrame(m.T.toarray().astype(np.int64), columns=cells), '', 'raw', False)
    o.preclust = pd.DataFrame({'cell': cells + ['not_in_data'], 'cluster': list(np.arange(n) % 3) + [9]})
    o.cluster()
    assert o.nn_idx is None and o.snn_graph is None
    assert o.n_clusters == 3 and list(o.clusters_targets) == [0, 1, 2]
    o.preclust = pd.DataFrame({'cell': cells[::-1], 'cluster': list(np.arange(n) % 3)})
    with pytest.raises(SystemExit):
        o.cluster()


# ---- N4: the other HVG selectors through the reference's own surface (alona/hvg.py:41-57) ----------------------
@pytest.mark.parametrize('case', ['hvgm_900x2500', 'hvgm_1500x4000'])
def test_hvg_brennecke_and_scran_match_reference(case, tmp_path):
    import scipy.sparse as sp
    from alona_b200.clustering import AlonaClustering
    from alona_b200.hvg import AlonaHighlyVariableGenes
    g = load_golden(case)
    m = golden_csr(g)
    e = sp.csr_matrix((g['ercc_data'], g['ercc_indices'], g['ercc_indptr']), shape=tuple(int(v) for v in g['ercc_shape']))
    genes = np.array(['g%d' % i for i in range(m.shape[1])])
    o = AlonaClustering()
    o.set_params({'output_directory': str(tmp_path), 'qc_auto': False, 'hvg_n': int(g['hvg_n'])})
    o.low_quality_cells = []
    o.data_norm = o.normalize(pd.DataFrame(m.T.toarray().astype(np.int64), index=genes), '', 'raw', False)
    o.data_ERCC = o.normalize(pd.DataFrame(e.T.toarray().astype(np.int64),
                                           index=['ERCC-%05d' % i for i in range(e.shape[1])]), '', 'raw', False)
    f = AlonaHighlyVariableGenes('Brennecke2013', int(g['hvg_n']), o.data_norm, o.data_ERCC)
    assert np.array_equal(f.find(), genes[g['hvg_brennecke']])                  # same genes, same (matrix) order
    f = AlonaHighlyVariableGenes('scran', int(g['hvg_n']), o.data_norm, o.data_ERCC)
    got = f.find()
    assert set(got.tolist()) == set(genes[g['hvg_scran']].tolist())
    assert (got == genes[g['hvg_scran']]).mean() > 0.99
    # and through the mixin: find_variable_genes() -> PCA() keeps the selected genes in original matrix order
    o.params['hvg_method'] = 'Brennecke2013'
    o.params['pca_n'] = 10
    o.params['seed'] = 1
    o.find_variable_genes()
    assert np.array_equal(o.hvg, genes[g['hvg_brennecke']])
    o.PCA(str(tmp_path) + '/csvs/pca.csv')
    assert o.pca_components.shape == (m.shape[0], 10)
    assert o.report['pca']['nnz_hvg'] == int(m[:, g['hvg_brennecke']].nnz)
    with pytest.raises(SystemExit):                                             # scran without spikes: hvg.py:187-190
        AlonaHighlyVariableGenes('scran', 10, o.data_norm, pd.DataFrame()).find()


def test_pca_regular_matches_reference(tmp_path):
    """--pca regular (alona/clustering.py:114-125: gene-centred exact SVD, first pca_n columns of x.v) through the
    implicitly centred device iteration: scores equal to the reference's after sign alignment."""
    from alona_b200.clustering import AlonaClustering
    g = load_golden('pcareg_1200x3000')
    m = golden_csr(g)
    genes = np.array(['g%d' % i for i in range(m.shape[1])])
    o = AlonaClustering()
    o.set_params({'output_directory': str(tmp_path), 'qc_auto': False, 'hvg_n': int(g['hvg_n']), 'pca': 'regular',
                  'pca_n': int(g['pca_n']), 'seed': 42})
    o.low_quality_cells = []
    o.data_ERCC = pd.DataFrame()
    o.data_norm = o.normalize(pd.DataFrame(m.T.toarray().astype(np.int64), index=genes), '', 'raw', False)
    o.find_variable_genes()
    assert set(o.hvg.tolist()) == set(genes[g['hvg_ranked']].tolist())
    o.PCA(str(tmp_path) + '/csvs/pca.csv')
    assert o.report['pca']['not_converged'] == 0
    err = R.column_rel_err(g['pca_components'], o.pca_components.values)
    assert err.max() < 1e-7, err
    # the uncentred default differs: the first component of --pca irlb is dominated by the gene means
    o.params['pca'] = 'irlb'
    o.PCA('')
    assert R.column_rel_err(g['pca_components'][:, :1], o.pca_components.values[:, :1]).max() > 1e-2
