"""TEST INFRASTRUCTURE — plain numpy/C restatements of the hot path's arithmetic.

Where oracle/alona_port.py calls the same third-party engines as the reference, this module
restates what those engines compute in plain numpy (and C for the O(N^2) kNN), working on
the cell-major CSR the CUDA path consumes.  These are the fast checkers the -m gpu parity
tests use at sizes where the dense pandas port is impractical; each is pinned against the
port / the live reference in tests/test_oracle_*.py.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this.

Third-party semantics restated:
  * pandas.cut equal-width bins: edges = linspace(min, max, nb+1), edges[0] -= 0.001*(max-min),
    right-closed, membership by searchsorted(side='left')          (pandas 3.0.2 tile.py)
  * pandas var/std: ddof=1 two-pass, NaN-skipping                   (pandas nanops)
  * scikit-learn 1.9.0 BallTree Euclidean rdist: sequential, non-fused FP64
    d += (a_j-b_j)*(a_j-b_j), j ascending                            (_dist_metrics.pxd.tp:39-49)
  * scipy 1.18.1 csr_matmat output order: per row, reverse of first-encounter order of the
    linked-list accumulator                                          (sparsetools/csr.h)
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


# ----------------------------------------------------------------------------------------
# N3: QC metrics and filters on CSR  (alona/cell.py:64-81, 141-173, 306-312)
# ----------------------------------------------------------------------------------------
def qc_metrics(csr, gene_sets=()):
    """reads per cell = data.sum(axis=0) (cell.py:306), genes detected = sum(data > 0, axis=0) (:307),
    reads on each gene set (numerators of :310-312), per gene the number of cells with a count > 0
    (simple_filters' `sum(x > 0)`, :161)."""
    pos = csr.data > 0
    rows = np.repeat(np.arange(csr.shape[0]), np.diff(csr.indptr))
    reads = np.bincount(rows, weights=np.where(pos, csr.data, 0), minlength=csr.shape[0]).astype(np.int64)
    det = np.bincount(rows[pos], minlength=csr.shape[0]).astype(np.int64)
    gene_cells = np.bincount(csr.indices[pos], minlength=csr.shape[1]).astype(np.int64)
    sets = []
    for mask in gene_sets:
        m = np.asarray(mask, dtype=bool)[csr.indices] & pos
        sets.append(np.bincount(rows[m], weights=csr.data[m], minlength=csr.shape[0]).astype(np.int64))
    return reads, det, gene_cells, sets


def remove_empty_masks(csr):
    """cell.py:64-81, AS WRITTEN: `cells` = zero entries per cell is compared with the number of CELLS
    (`cells == total_cells`, :72) and `genes` = zero entries per gene with the number of GENES (:77) - the two
    totals are swapped relative to the intent, so an all-zero cell is only removed when G == N.  The genes are
    tested on the frame AFTER the cell removal (the frame is re-sliced first, the masks were computed before).
    The drop-in reproduces the reference's behaviour, not its intent."""
    n, g = csr.shape
    _, det, gene_cells, _ = qc_metrics(csr)
    keep_cells = (g - det) != n
    keep_genes = (n - gene_cells) != g
    return keep_cells, keep_genes


def simple_filter_masks(csr, minreads, minexpgenes):
    """cell.py:141-173 (raw counts): keep cells with reads > minreads; THEN, on the remaining cells, keep genes
    expressed in more than `minexpgenes` cells (integral value) or in more than that fraction of the cells.
    The comparisons are evaluated exactly like the reference's (Python ints / numpy float division)."""
    reads, _, _, _ = qc_metrics(csr)
    keep_cells = reads > minreads
    sub = csr[keep_cells]
    _, _, gene_cells, _ = qc_metrics(sub)
    thres = float(minexpgenes)
    if thres > 0:
        if thres.is_integer():
            keep_genes = gene_cells > thres
        else:
            keep_genes = gene_cells / sub.shape[0] > thres
    else:
        keep_genes = np.ones(csr.shape[1], dtype=bool)
    return keep_cells, keep_genes


# ----------------------------------------------------------------------------------------
# N2: per-cluster mean / median of the normalised expression  (alona/celltypes.py:42-68)
# ----------------------------------------------------------------------------------------
def cluster_mean_median(csr, cluster, n_clusters, lib=None):
    """data_norm.groupby(cluster, axis=1).aggregate(np.mean | np.median) restated on the dense normalised matrix
    of a SMALL input: per cluster, np.mean / np.median over its cells, gene by gene (an empty cluster gives NaN).
    (`groupby(axis=1)` no longer exists in the pandas this image ships, 3.0.2, so the reference's own call cannot
    be executed here; this is the same per-group numpy reduction it applied.)  Returns [C][G] arrays and the
    per-cluster residual sum of squares of find_markers' one-way model (find_markers.py:96-97)."""
    x = np.zeros(csr.shape)
    rows = np.repeat(np.arange(csr.shape[0]), np.diff(csr.indptr))
    x[rows, csr.indices] = normalized_values(csr, lib)
    cluster = np.asarray(cluster)
    mean = np.full((n_clusters, csr.shape[1]), np.nan)
    med = np.full((n_clusters, csr.shape[1]), np.nan)
    rss = np.zeros(csr.shape[1])
    for c in range(n_clusters):
        sel = cluster == c
        if sel.any():
            mean[c] = np.mean(x[sel], axis=0)
            med[c] = np.median(x[sel], axis=0)
            rss += ((x[sel] - mean[c]) ** 2).sum(axis=0)
    return mean, med, rss


# ----------------------------------------------------------------------------------------
# normalise + per-gene moments on CSR  (alona/cell.py:241-243, alona/hvg.py:148,157)
# ----------------------------------------------------------------------------------------
def libsize(csr):
    """Exact per-cell totals S_c (cell.py:241)."""
    return np.asarray(csr.astype(np.int64).sum(axis=1)).ravel()


def normalized_values(csr, lib=None):
    """x = log2(fl(fl(c/S)*1e4) + 1) for every stored nonzero, FP64 (cell.py:242-243)."""
    if lib is None:
        lib = libsize(csr)
    rows = np.repeat(np.arange(csr.shape[0]), np.diff(csr.indptr))
    c = csr.data.astype(np.float64)
    return np.log2((c / lib[rows].astype(np.float64)) * 10000 + 1)


def gene_moments(csr, lib=None):
    """Returns (sum_x[G], sum_x2[G], nnz[G]) over ALL cells (zeros contribute 0)."""
    x = normalized_values(csr, lib)
    g = csr.shape[1]
    sx = np.bincount(csr.indices, weights=x, minlength=g)
    sx2 = np.bincount(csr.indices, weights=x * x, minlength=g)
    nz = np.bincount(csr.indices, minlength=g)
    return sx, sx2, nz


def gene_moments_fixed(csr, lib=None):
    """The K1 accumulators of include/alona_b200.h (ab_norm_moments): per gene {lo, hi of sum x in units of 2^-45,
    lo, hi of sum x^2 in units of 2^-46}, value = lo + hi * 2^32, as int64 [G, 4].  Python integers: exact."""
    x = normalized_values(csr, lib)
    fx = np.rint(x * float(1 << 45)).astype(np.uint64)
    fxx = np.rint((x * x) * float(1 << 46)).astype(np.uint64)
    out = np.zeros((csr.shape[1], 4), dtype=np.int64)
    for g, a, b in zip(csr.indices.tolist(), fx.tolist(), fxx.tolist()):
        out[g, 0] += a & 0xFFFFFFFF
        out[g, 1] += a >> 32
        out[g, 2] += b & 0xFFFFFFFF
        out[g, 3] += b >> 32
    return out


def moments_from_fixed(acc):
    """ab_moments_finalize: exact integer recombination, then FP64."""
    sx = np.array([(int(a[1]) << 32) + int(a[0]) for a in acc], dtype=object)
    sxx = np.array([(int(a[3]) << 32) + int(a[2]) for a in acc], dtype=object)
    return (np.array([float(v) / float(1 << 45) for v in sx]), np.array([float(v) / float(1 << 46) for v in sxx]))


def gene_mean_var(csr, lib=None):
    """mean and ddof=1 variance per gene the way pandas computes them (two-pass):
    var = [ sum_nz (x-mean)^2 + (N-nnz)*mean^2 ] / (N-1)."""
    n, g = csr.shape
    x = normalized_values(csr, lib)
    sx = np.bincount(csr.indices, weights=x, minlength=g)
    nz = np.bincount(csr.indices, minlength=g)
    mean = sx / n
    dev = x - mean[csr.indices]
    ss = np.bincount(csr.indices, weights=dev * dev, minlength=g) + (n - nz) * mean * mean
    return mean, ss / (n - 1)


# ----------------------------------------------------------------------------------------
# Seurat-style HVG from per-gene mean/variance  (alona/hvg.py:146-165)
# ----------------------------------------------------------------------------------------
def cut_bins(mean, num_bin=20):
    """pandas.cut(mean, num_bin) bin ids (0..num_bin-1), restated."""
    mn, mx = float(np.min(mean)), float(np.max(mean))
    edges = np.linspace(mn, mx, num_bin + 1)
    edges[0] -= 0.001 * (mx - mn)
    b = np.searchsorted(edges, mean, side='left') - 1
    return b, edges


def hvg_seurat_from_moments(mean, var, hvg_n, num_bin=20):
    """Returns (ranked gene indices [<=hvg_n], z[G]).  Ties broken by gene index; NaN last."""
    with np.errstate(divide='ignore', invalid='ignore'):
        disp = var / mean                                        # hvg.py:157
        b, _ = cut_bins(mean, num_bin)
        z = np.full(mean.shape[0], np.nan)
        for q in range(num_bin):                                 # hvg.py:154
            idx = np.where(b == q)[0]
            if idx.size == 0:
                continue
            d = disp[idx]
            ok = ~np.isnan(d)
            cnt = ok.sum()
            if cnt == 0:
                continue
            mu = d[ok].sum() / cnt
            if cnt > 1:
                sd = np.sqrt(((d[ok] - mu) ** 2).sum() / (cnt - 1))
            else:
                sd = np.nan
            z[idx] = (d - mu) / sd                               # hvg.py:158
    key = np.where(np.isnan(z), -np.inf, z)
    order = np.lexsort((np.arange(z.shape[0]), -key))            # z desc, index asc
    order = order[~np.isnan(z[order])]
    nan_tail = np.where(np.isnan(z))[0]
    order = np.concatenate([order, nan_tail])                    # NaN last (sort_values default)
    return order[:hvg_n], z


def hvg_matrix(csr, hvg_idx, lib=None):
    """A_H: cells x H CSR of normalised values, HVG columns in ORIGINAL gene order
    (alona/clustering.py:104-105).  Returns (csr float64, sorted hvg column ids)."""
    import scipy.sparse as sp
    cols = np.sort(np.asarray(hvg_idx))
    x = normalized_values(csr, lib)
    full = sp.csr_matrix((x, csr.indices, csr.indptr), shape=csr.shape)
    return full[:, cols].tocsr(), cols


# ----------------------------------------------------------------------------------------
# IRLBA on a cell-major CSR, optionally with simulated row shards
# (alona/irlbpy/irlb.py:138-258; sharding per SURVEY.md §8(e))
# ----------------------------------------------------------------------------------------
def lanczos_csr(At, nval, tol=1e-4, maxit=1000, seed=None, nshards=1, v0=None):
    """At: cells x H (CSR or dense) == transpose of the reference's A (H x N).
    Contractions over cells are computed per shard and summed in rank order, the way the
    multi-GPU path all-reduces them.  Returns dict(U,s,V,steps,nmult,sum_panel)."""
    n, m = At.shape
    A_T = At.T.tocsr() if hasattr(At, 'tocsr') else np.ascontiguousarray(At.T)
    bounds = np.linspace(0, n, nshards + 1).astype(np.int64)

    def a_times(v):                    # A.v  : contraction over cells, sharded
        acc = np.zeros(m)
        for r in range(nshards):
            lo, hi = bounds[r], bounds[r + 1]
            acc = acc + np.asarray(A_T[:, lo:hi] @ v[lo:hi]).ravel()
        return acc

    def at_times(w):                   # A^T.w : output over cells, local
        return np.asarray(At @ w).ravel()

    def sharded_dot(P, f):             # P^T f with P n x c
        acc = np.zeros(P.shape[1])
        for r in range(nshards):
            lo, hi = bounds[r], bounds[r + 1]
            acc = acc + P[lo:hi].T @ f[lo:hi]
        return acc

    def sharded_norm(f):
        acc = 0.0
        for r in range(nshards):
            lo, hi = bounds[r], bounds[r + 1]
            acc = acc + float(f[lo:hi] @ f[lo:hi])
        return np.sqrt(acc)

    eps2 = 2 * np.finfo(float).eps
    inv = lambda x: (1.0 / x) if x > eps2 else 0.0
    nu = nval
    work = min(nu + 20, 3 * nu, n)
    V = np.zeros((n, work)); W = np.zeros((m, work)); B = np.zeros((work, work))
    if v0 is None:
        np.random.seed(seed)
        v0 = np.random.randn(n)
    V[:, 0] = v0 / sharded_norm(v0)
    nmult = 0; it = 0; keep = nu; smax = 1.0; sum_panel = 0
    while it < maxit:
        j = keep if it > 0 else 0
        W[:, j] = a_times(V[:, j]); nmult += 1
        if it > 0:
            W[:, j] -= W[:, :j] @ (W[:, :j].T @ W[:, j])
        s = np.linalg.norm(W[:, j]); W[:, j] *= inv(s)
        while j < work:
            F = at_times(W[:, j]); nmult += 1
            F = F - s * V[:, j]
            F = F - V[:, :j + 1] @ sharded_dot(V[:, :j + 1], F)
            sum_panel += j + 1
            fn = sharded_norm(F); F = F * inv(fn)
            if j < work - 1:
                V[:, j + 1] = F; B[j, j] = s; B[j, j + 1] = fn
                W[:, j + 1] = a_times(V[:, j + 1]); nmult += 1
                W[:, j + 1] -= fn * W[:, j]
                W[:, j + 1] -= W[:, :j + 1] @ (W[:, :j + 1].T @ W[:, j + 1])
                s = np.linalg.norm(W[:, j + 1]); W[:, j + 1] *= inv(s)
            else:
                B[j, j] = s
            j += 1
        Ub, sb, Vbt = np.linalg.svd(B)
        R = fn * Ub[work - 1, :]
        smax = sb[0] if it == 0 else max(sb[0], smax)
        conv = int(np.sum(np.abs(R[:nu]) < tol * smax))
        if conv >= nu:
            break
        keep = min(max(conv + nu, keep), work - 3)
        V[:, :keep] = V[:, :work] @ Vbt.T[:, :keep]
        V[:, keep] = F
        B = np.zeros((work, work))
        B[np.arange(keep), np.arange(keep)] = sb[:keep]
        B[:keep, keep] = R[:keep]
        W[:, :keep] = W[:, :work] @ Ub[:, :keep]
        it += 1
    return {'U': W[:, :work] @ Ub[:, :nu], 's': sb[:nu], 'V': V[:, :work] @ Vbt.T[:, :nu],
            'steps': it, 'nmult': nmult, 'sum_panel': sum_panel}


def align_signs(ref, got):
    """Flip columns of `got` so that each has a positive inner product with `ref`."""
    sgn = np.sign(np.sum(ref * got, axis=0))
    sgn[sgn == 0] = 1.0
    return got * sgn


def column_rel_err(ref, got):
    """Per-column ||ref-got|| / ||ref|| after sign alignment."""
    g = align_signs(ref, got)
    return np.linalg.norm(ref - g, axis=0) / np.maximum(np.linalg.norm(ref, axis=0), 1e-300)


# ----------------------------------------------------------------------------------------
# exact kNN, brute force, FP64 sequential non-fused rdist, ties by index  (C, OpenMP)
# ----------------------------------------------------------------------------------------
_knn_lib = None


def build_c(force=False):
    """Compiles oracle/knn_bruteforce.c -> oracle/_build/liboracle_knn.so (gcc)."""
    out_dir = os.path.join(_HERE, '_build')
    os.makedirs(out_dir, exist_ok=True)
    so = os.path.join(out_dir, 'liboracle_knn.so')
    src = os.path.join(_HERE, 'knn_bruteforce.c')
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(['gcc', '-O2', '-fPIC', '-shared', '-fopenmp', '-ffp-contract=off',
                               '-o', so, src, '-lm'])
    return so


def _lib():
    global _knn_lib
    if _knn_lib is None:
        _knn_lib = ctypes.CDLL(build_c())
        _knn_lib.oracle_knn_bruteforce.argtypes = [
            ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int,
            ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
        _knn_lib.oracle_knn_bruteforce.restype = ctypes.c_int
    return _knn_lib


def knn_bruteforce(db, k, queries=None, threads=0):
    """Exact kNN of `queries` (default: db itself) in `db` (float64 N x d).
    Returns (idx int64 [nq,k] 0-based, rdist float64 [nq,k]) sorted by (rdist, idx)."""
    db = np.ascontiguousarray(db, dtype=np.float64)
    q = db if queries is None else np.ascontiguousarray(queries, dtype=np.float64)
    nq, d = q.shape
    idx = np.empty((nq, k), dtype=np.int64)
    rd = np.empty((nq, k), dtype=np.float64)
    rc = _lib().oracle_knn_bruteforce(db.ctypes.data, db.shape[0], q.ctypes.data, nq, d, k,
                                      idx.ctypes.data, rd.ctypes.data, threads)
    if rc != 0:
        raise RuntimeError('oracle_knn_bruteforce failed: %d' % rc)
    return idx, rd


def knn_tie_rows(rdist_kplus1):
    """Rows whose k-th and (k+1)-th reduced distances are exactly equal (ambiguous in the
    reference: BallTree keeps whichever it visited first, sklearn/utils/_heap.pyx:46)."""
    return np.where(rdist_kplus1[:, -1] == rdist_kplus1[:, -2])[0]


# ----------------------------------------------------------------------------------------
# SNN  (alona/clustering.py:204-231 + scipy csr_matmat ordering)
# ----------------------------------------------------------------------------------------
def prune_threshold(k, prune_snn):
    """Smallest integer overlap v in 0..k kept by the reference's test
    v/(k+(k-v)) > prune_snn (clustering.py:225-227), evaluated exactly as Python does."""
    for v in range(0, k + 1):
        if v / (k + (k - v)) > prune_snn:
            return v
    return k + 1


def snn_restated(nn_idx0, k, prune_snn, return_all=False, rows=None):
    """nn_idx0: 0-based (N,k) table with column 0 == the row's own cell.
    Returns (edges [E,2] int64, overlap [E] int64) in the reference's order:
    rows ascending; within row i, REVERSE first-encounter order when scanning
    j in sorted(N(i)) and, for each j, m in sorted(R(j)), R = self-inclusive reverse lists."""
    n = nn_idx0.shape[0]
    vmin = prune_threshold(k, prune_snn)
    nbr = np.sort(nn_idx0, axis=1)
    # reverse lists, sorted by m
    src = np.repeat(np.arange(n), k)
    dst = nn_idx0.ravel()
    order = np.lexsort((src, dst))
    dst_s, src_s = dst[order], src[order]
    start = np.searchsorted(dst_s, np.arange(n + 1))
    e_src, e_dst, e_cnt = [], [], []
    for i in (range(n) if rows is None else rows):          # rows: only these rows of the graph (large inputs)
        seq = np.concatenate([src_s[start[j]:start[j + 1]] for j in nbr[i]])
        uniq, first, cnt = np.unique(seq, return_index=True, return_counts=True)
        o = np.argsort(-first, kind='stable')
        u, c = uniq[o], cnt[o]
        if not return_all:
            keep = c >= vmin
            u, c = u[keep], c[keep]
        e_src.append(np.full(u.shape[0], i, dtype=np.int64)); e_dst.append(u); e_cnt.append(c)
    edges = np.stack([np.concatenate(e_src), np.concatenate(e_dst)], axis=1)
    return edges, np.concatenate(e_cnt)


def snn_overlap_bruteforce(nn_idx0):
    """Dense |N(i) ∩ N(m)| for tiny N (set intersection; independent of the matmul form)."""
    n = nn_idx0.shape[0]
    sets = [set(r.tolist()) for r in nn_idx0]
    out = np.zeros((n, n), dtype=np.int64)
    for i in range(n):
        for m in range(n):
            out[i, m] = len(sets[i] & sets[m])
    return out
