import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a real B200 (run with -m gpu on the GPU box)')


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name + '.npz'))


def golden_csr(g):
    """The fixture's input counts (cell-major CSR).  The C1 family shares one input file, named by `input`."""
    import scipy.sparse as sp
    if 'input' in g.files:
        g = load_golden(str(g['input']))
    shape = tuple(int(x) for x in g['shape'])
    return sp.csr_matrix((g['data'].astype(np.int32), g['indices'].astype(np.int32), g['indptr']), shape=shape)


PIPE_CASES = ['pipe_small', 'pipe_medium', 'pipe_k20']
# BASELINE.json configs[0] in full and C1-shaped variants at the IRLBA work dimensions the named configs run
# (m_b = 60 / 70 / 95 / 120), minted from the live reference by oracle/make_golden.py --c1
C1_CASES = ['c1_full', 'c1_h2000_d50_k20', 'c1_h2000_d75_k20', 'c1_h3000_d100_k30']
QC_CASES = ['qc_square_400', 'qc_rect_700x300']
CLEXP_CASES = ['clexp_700x900', 'clexp_401x500']
KNN_CASES = ['knn_mix_3000x50_k20', 'knn_mix_2000x50_k10', 'knn_mix_1500x100_k30',
             'knn_mix_1000x50_k100']


@pytest.fixture(scope='session')
def has_cuda():
    import torch
    return torch.cuda.is_available()
