"""GPU tier, N2: per-cluster mean / median tables and model statistics from the device CSR."""
import os

import numpy as np
import pandas as pd
import pytest
import torch

from conftest import load_golden, golden_csr, CLEXP_CASES
from oracle import restate as R

pytestmark = pytest.mark.gpu

MEAN_RTOL = 1e-12        # FP64 sums in a different order than numpy's pairwise mean
VALUE_RTOL = 1e-14       # a median is one (or the mean of two) normalised value(s): log2 within a few ulps


@pytest.fixture(scope='module')
def eng():
    from alona_b200.engine import get_engine
    return get_engine()


@pytest.mark.parametrize('case', CLEXP_CASES)
def test_cluster_tables_match_reference(eng, case):
    g = load_golden(case)
    m = golden_csr(g)
    ids = g['cluster_ids']
    dense_id = np.searchsorted(ids, g['labels']).astype(np.int32)
    A = eng.upload_csr(m)
    lib, _, _ = eng.norm_moments(A)
    mom = eng.cluster_moments(A, lib, dense_id, len(ids))
    sizes = np.bincount(dense_id, minlength=len(ids))
    mean = mom['sum'].cpu().numpy() / sizes[:, None]
    assert np.allclose(mean, g['mean'], rtol=MEAN_RTOL, atol=1e-15)
    assert np.array_equal(mom['nnz'].cpu().numpy(), np.stack([np.asarray((m[dense_id == c] > 0).sum(axis=0)).ravel()
                                                              for c in range(len(ids))]))
    med, info = eng.cluster_median(A, lib, dense_id, len(ids), nnz=mom['nnz'])
    med = med.cpu().numpy()
    assert np.array_equal(med == 0, g['median'] == 0)
    assert np.allclose(med, g['median'], rtol=VALUE_RTOL, atol=0)
    assert 0 < info['gathered'] < m.nnz                      # only the pairs that need their values are gathered
    # batching by clusters gives the same table
    med2, info2 = eng.cluster_median(A, lib, dense_id, len(ids), nnz=mom['nnz'], max_batch=1)
    assert torch.equal(torch.from_numpy(med), med2.cpu()) and info2['batch'] < info['batch'] + 1
    # model statistics of find_markers.py:96-97
    n_c = sizes.astype(np.float64)
    rss = mom['sumsq'].cpu().numpy().sum(axis=0) - (n_c[:, None] * mean * mean).sum(axis=0)
    sigma2 = rss / (m.shape[0] - len(ids))
    assert np.allclose(sigma2, g['sigma2'], rtol=1e-9, atol=1e-13)


def test_cluster_median_ignored_cells_and_empty_cluster(eng):
    g = load_golden('clexp_401x500')
    m = golden_csr(g)
    lab = np.searchsorted(g['cluster_ids'], g['labels']).astype(np.int32)
    lab[10:40] = -1                                           # ignored cells
    C = len(g['cluster_ids']) + 1                             # one cluster without any cell
    A = eng.upload_csr(m)
    lib, _, _ = eng.norm_moments(A)
    med, _ = eng.cluster_median(A, lib, lab, C)
    med = med.cpu().numpy()
    keep = lab >= 0
    _, ref, _ = R.cluster_mean_median(m[keep], lab[keep], C, lib=R.libsize(m)[keep])
    assert np.all(np.isnan(med[C - 1])) and np.all(np.isnan(ref[C - 1]))
    assert np.allclose(med[:C - 1], ref[:C - 1], rtol=VALUE_RTOL, atol=0)


@pytest.mark.parametrize('case', CLEXP_CASES)
def test_celltypes_dropin_mean_and_median_exp(case, tmp_path):
    from alona_b200.celltypes import AlonaCellTypePred
    g = load_golden(case)
    m = golden_csr(g)
    genes = ['g%d' % i for i in range(m.shape[1])]
    cells = ['c%d' % i for i in range(m.shape[0])]
    df = pd.DataFrame(m.T.toarray().astype(np.int64), index=genes, columns=cells)
    wd = str(tmp_path)
    os.makedirs(wd + '/csvs', exist_ok=True)
    o = AlonaCellTypePred()
    o.params = {'qc_auto': False, 'output_directory': wd}
    o.low_quality_cells = []
    o.data_norm = o.normalize(df, '', 'raw', False)
    o.leiden_cl = list(g['labels'])
    ids, counts = np.unique(g['labels'], return_counts=True)
    o.clusters_targets = ids[counts > 10]                      # leiden_prep: clusters above --ignore_small_clusters
    o.anno = None
    keep = np.isin(ids, o.clusters_targets)
    mean = o.mean_exp()
    med = o.median_exp(o.data_norm)
    assert list(mean.columns) == list(ids[keep]) and list(mean.index) == genes
    assert np.allclose(mean.values, g['mean'][keep].T, rtol=MEAN_RTOL, atol=1e-15)
    assert np.allclose(med.values, g['median'][keep].T, rtol=VALUE_RTOL, atol=0)
    back = pd.read_csv(wd + '/csvs/median_exp.tsv', sep='\t', index_col=0)
    assert np.allclose(back.values, med.values, rtol=1e-15, atol=0) and list(back.index) == genes
    assert os.path.exists(wd + '/csvs/mean_exp.tsv')
    st = o.cluster_stats()
    assert st['resid_df'] == int(counts[keep].sum()) - int(keep.sum())
    assert np.allclose(st['coef'].values, g['mean'][keep], rtol=MEAN_RTOL, atol=1e-15)
