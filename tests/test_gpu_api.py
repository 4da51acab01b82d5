"""GPU tier: the reference-facing Python API (same class / method / attribute names as
alona/cell.py, alona/hvg.py, alona/clustering.py) against fixtures minted from the reference."""
import os

import numpy as np
import pandas as pd
import pytest

from conftest import load_golden, golden_csr, PIPE_CASES
from oracle import restate as R

pytestmark = pytest.mark.gpu


def _params(g, wd):
    return {'qc_auto': False, 'hvg_method': 'seurat', 'hvg_n': int(g['hvg_n']), 'pca': 'irlb',
            'pca_n': int(g['pca_n']), 'seed': int(g['seed']), 'nn_k': int(g['k']),
            'prune_snn': float(g['prune']), 'output_directory': wd, 'leiden_res': 0.8}


@pytest.mark.parametrize('case', PIPE_CASES)
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
    same = np.all(o.nn_idx == g['nn_idx'], axis=1)
    assert same.mean() >= 0.999
    o.snn(int(g['k']), float(g['prune']))
    assert list(o.snn_graph.columns) == ['source_node', 'target_node']
    if same.all():
        assert np.array_equal(o.snn_graph.values, g['snn_graph'])
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
        AlonaHighlyVariableGenes('scran', 10, None, None).find()
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
