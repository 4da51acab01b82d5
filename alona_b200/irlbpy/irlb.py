"""lanczos() — drop-in for alona.irlbpy.lanczos (alona/irlbpy/irlb.py:95-258).

Same signature and result object (U, s, V, steps, nmult).  The iteration runs on the device
(ab_irlb): CSR SpMV in both directions, classical Gram-Schmidt panels, one-sided Jacobi SVD of
the small bidiagonal-plus-spike matrix, Ritz updates — all FP64, same restart logic, so the
trajectory follows the reference's.  The start vector is drawn exactly as the reference draws
it (np.random.seed(seed); np.random.randn(n), irlb.py:151-152) and scattered to the shards.
"""
import numpy as np
import scipy.sparse as sp
import torch

from ..engine import get_engine, DeviceCSR


class LanczosResult():
    def __init__(self, **kwargs):
        for key in kwargs:
            setattr(self, key, kwargs[key])


class MatrixShapeException(Exception):
    pass


def _device_matrix(A, eng):
    """Accepts the HVG-restricted device matrix (cells x H DeviceCSR) or a host matrix in the
    reference's orientation (H x N: ndarray, DataFrame or scipy sparse)."""
    if isinstance(A, DeviceCSR):
        return A
    if hasattr(A, 'values') and not sp.issparse(A):
        A = A.values
    if getattr(A, 'ndim', 2) != 2:
        raise MatrixShapeException('alona_b200.lanczos supports 2-D matrices only '
                                   '(the 1-D Hankel path of irlb.py:124-133 is unreachable from alona)')
    At = sp.csr_matrix(A.T) if sp.issparse(A) else sp.csr_matrix(np.asarray(A, dtype=np.float64).T)
    At.sort_indices()
    n_total, H = At.shape
    lo = (n_total * eng.rank) // eng.world
    hi = (n_total * (eng.rank + 1)) // eng.world
    part = At[lo:hi]
    return DeviceCSR(torch.from_numpy(part.indptr.astype(np.int64)).to(eng.device),
                     torch.from_numpy(part.indices.astype(np.int32)).to(eng.device),
                     torch.from_numpy(part.data.astype(np.float64)).to(eng.device), H, lo, n_total)


def lanczos(A, nval, tol=0.0001, maxit=50, center=None, scale=None, L=None, seed=None, device_result=False,
            gene_center=None):
    """Truncated SVD by augmented implicitly-restarted Lanczos bidiagonalisation.

    Returns LanczosResult(U [H x nval], s [nval], V [N x nval], steps, nmult) such that
    A.dot(V) ~= U * s, for A given in the reference's orientation (genes x cells).

    gene_center (extension, device tensor of H doubles): per-gene means for implicit centring, operator
    A - gene_center.1^T (what AlonaClustering.PCA uses for `--pca regular`); not the reference's `center` argument.
    """
    if center is not None or scale is not None or L is not None:
        raise NotImplementedError('center/scale/L are never passed by alona (clustering.py:108-109) '
                                  'and are not supported by alona_b200')
    eng = get_engine()
    AH = _device_matrix(A, eng)
    n_total, H = AH.n_total, AH.n_genes
    if min(H, n_total) < 2:
        raise MatrixShapeException("The input matrix must be at least 2x2.")
    np.random.seed(seed)                               # irlb.py:151
    v0 = np.random.randn(n_total)                      # irlb.py:152
    lo, hi = AH.row_begin, AH.row_begin + AH.n_local
    v0_local = torch.from_numpy(np.ascontiguousarray(v0[lo:hi])).to(eng.device)
    res = eng.irlb(AH, nval, tol=tol, maxit=maxit, v0_local=v0_local, center=gene_center)
    if device_result:
        return res
    V = res['V']
    if eng.world > 1:
        counts = [(n_total * (r + 1)) // eng.world - (n_total * r) // eng.world for r in range(eng.world)]
        V = eng.allgather_rows(V, counts)
    return LanczosResult(U=res['U'].cpu().numpy(), s=res['s'].cpu().numpy(), V=V.cpu().numpy(),
                         steps=res['steps'], nmult=res['nmult'])
