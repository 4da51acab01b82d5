"""GPU tier, N3: QC metrics and filters on the device CSR against the reference-minted fixtures and the oracle."""
import numpy as np
import pandas as pd
import pytest
import scipy.sparse as sp
import torch

from conftest import load_golden, golden_csr, QC_CASES
from oracle import restate as R

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def eng():
    from alona_b200.engine import get_engine
    return get_engine()


@pytest.mark.parametrize('case', QC_CASES)
def test_qc_metrics_kernel_matches_reference(eng, case):
    g = load_golden(case)
    m = golden_csr(g)
    mt = pd.Index(g['genes']).str.contains('^mt-', regex=True, case=False)
    A = eng.upload_csr(m)
    q = eng.qc_metrics(A, mt.astype(np.uint8))
    assert np.array_equal(q['reads_per_cell'].cpu().numpy(), g['reads_per_cell'])
    assert np.array_equal(q['genes_det'].cpu().numpy(), g['no_genes_det'])
    reads, det, gene_cells, sets = R.qc_metrics(m, [mt])
    assert np.array_equal(q['gene_cells'].cpu().numpy(), gene_cells)
    assert np.array_equal(q['set_reads'][0].cpu().numpy(), sets[0])
    assert int(q['set_reads'][1].abs().sum()) == 0


@pytest.mark.parametrize('case', QC_CASES)
@pytest.mark.parametrize('remove_mito', ['no', 'yes'])
def test_alona_cell_filters_dropin(case, remove_mito):
    """remove_empty -> remove_mito -> simple_filters through the mirror of alona/cell.py, reference frame in."""
    from alona_b200.cell import AlonaCell
    g = load_golden(case)
    m = golden_csr(g)
    genes = g['genes']
    cells = np.array(['c%d' % i for i in range(m.shape[0])])
    df = pd.DataFrame(m.T.toarray().astype(np.int64), index=genes, columns=cells)
    for ti, thr in enumerate(g['thresholds']):
        o = AlonaCell()
        o.params = {'dataformat': 'raw', 'minreads': int(g['minreads']), 'minexpgenes': float(thr),
                    'remove_mito': remove_mito, 'qc_auto': False}
        o.data = df
        for step, fn in (('after_remove_empty', o.remove_empty), ('after_remove_mito', o.remove_mito),
                         ('after_simple_filters', o.simple_filters)):
            fn()
            key = '%s_mito%s_t%d' % (step, remove_mito, ti)
            if not isinstance(o.data, pd.DataFrame):
                assert list(o.data.index) == list(genes[g[key + '_genes']]), key
                assert list(o.data.columns) == list(cells[g[key + '_cells']]), key
        # the surviving counts are the reference's frame, entry for entry
        ref = df.loc[genes[g[key + '_genes']], cells[g[key + '_cells']]]
        assert np.array_equal(o.data.to_dataframe().values, ref.values)
        # and they feed normalize() without leaving the device
        dn = o.normalize(o.data, '', 'raw', False)
        assert dn.shape == ref.shape
        lib = ref.sum(axis=0).values
        assert np.array_equal(dn.libsize.cpu().numpy(), lib)


def test_filter_csr_random_masks_against_scipy(eng):
    rng = np.random.RandomState(4)
    n, G = 20000, 3000
    m = sp.random(n, G, density=0.03, random_state=rng, format='csr', dtype=np.float64)
    m.data = rng.randint(0, 6, size=m.nnz).astype(np.int32)            # explicit zeros included
    m.sort_indices()
    A = eng.upload_csr(m, canonical=False)                             # the filter kernels drop them themselves
    keep_row = rng.rand(n) < 0.7
    keep_gene = rng.rand(G) < 0.5
    gene_map = np.where(keep_gene, np.cumsum(keep_gene) - 1, -1).astype(np.int32)
    out, src = eng.filter_csr(A, keep_row, gene_map, n_genes_out=int(keep_gene.sum()))
    ref = sp.csr_matrix(m)
    ref.eliminate_zeros()
    ref = ref[keep_row][:, keep_gene]
    ref.sort_indices()
    assert np.array_equal(out.rowptr.cpu().numpy(), ref.indptr)
    assert np.array_equal(out.col.cpu().numpy(), ref.indices)
    assert np.array_equal(out.val.cpu().numpy(), ref.data)
    assert np.array_equal(src.cpu().numpy(), np.nonzero(keep_row)[0])
    # no mask at all = drop explicit zeros only; filtering twice changes nothing
    same, _ = eng.filter_csr(out)
    assert torch.equal(same.rowptr, out.rowptr) and torch.equal(same.col, out.col) and torch.equal(same.val, out.val)
    # metrics of the filtered matrix
    q = eng.qc_metrics(out)
    reads, det, gene_cells, _ = R.qc_metrics(ref)
    assert np.array_equal(q['reads_per_cell'].cpu().numpy(), reads)
    assert np.array_equal(q['genes_det'].cpu().numpy(), det)
    assert np.array_equal(q['gene_cells'].cpu().numpy(), gene_cells)
    assert int(q['reads_per_cell'].sum()) == int(ref.data.sum())


def test_qc_edge_cases(eng):
    # empty matrix rows, a single row, everything filtered away
    m = sp.csr_matrix((np.array([3, 0, 2], dtype=np.int32), np.array([0, 1, 4], dtype=np.int32),
                       np.array([0, 0, 2, 3, 3], dtype=np.int64)), shape=(4, 5))
    A = eng.upload_csr(m, canonical=False)
    q = eng.qc_metrics(A)
    assert q['reads_per_cell'].tolist() == [0, 3, 2, 0]
    assert q['genes_det'].tolist() == [0, 1, 1, 0]
    assert q['gene_cells'].tolist() == [1, 0, 0, 0, 1]
    out, src = eng.filter_csr(A, np.array([0, 0, 0, 0], dtype=np.uint8))
    assert out.n_local == 0 and out.rowptr.tolist() == [0] and src.numel() == 0
    out, src = eng.filter_csr(A, np.array([1, 1, 0, 1], dtype=np.uint8), np.array([-1, 0, 1, 2, 3], dtype=np.int32))
    assert out.rowptr.tolist() == [0, 0, 0, 0] and out.col.numel() == 0 and src.tolist() == [0, 1, 3]
