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
    import scipy.sparse as sp
    shape = tuple(int(x) for x in g['shape'])
    return sp.csr_matrix((g['data'], g['indices'], g['indptr']), shape=shape)


PIPE_CASES = ['pipe_small', 'pipe_medium', 'pipe_k20']
QC_CASES = ['qc_square_400', 'qc_rect_700x300']
CLEXP_CASES = ['clexp_700x900', 'clexp_401x500']
KNN_CASES = ['knn_mix_3000x50_k20', 'knn_mix_2000x50_k10', 'knn_mix_1500x100_k30',
             'knn_mix_1000x50_k100']


@pytest.fixture(scope='session')
def has_cuda():
    import torch
    return torch.cuda.is_available()
