"""CPU tier: the arithmetic behind the kNN certificate (alona_b200/csrc/ab_knn_tc.cu), restated in
numpy on a small embedding.  No GPU, no library: this pins the INEQUALITIES the CUDA path relies on, not its output.

The tensor pass multiplies q_hi = fp16(q) against x_hi + x_lo (22 bits of x) with FP32 accumulation and adds the norm of
x: a score s~(q, x).  The re-rank certifies a query when  sqrt(tau + |q_hi|^2 - eps) - delta > sqrt(D_k),  with
delta = |q - q_hi| and eps = kappa (|q| + R)^2, kappa = 1.25 K 2^-24 + 2^-19.  That is sound iff for EVERY pair
    sqrt(max(s~ + |q_hi|^2 - eps, 0)) - delta  <=  |x - q|  <=  sqrt(s~ + |q_hi|^2 + eps) + delta .
"""
import numpy as np


def _operands(emb):
    mean = emb.mean(axis=0)
    c = emb - mean
    R = float(np.sqrt((c * c).sum(axis=1)).max() * 1.000001)
    sc = 2.0 ** (5 - int(np.floor(np.log2(R))))                 # largest norm into [32, 64)
    xs = c * sc                                                  # FP64 scaled coordinates
    xf = xs.astype(np.float32)
    hi = xf.astype(np.float16)
    lo = (xf - hi.astype(np.float32)).astype(np.float16)
    n2 = (xf.astype(np.float64) ** 2).sum(axis=1)
    n0 = n2.astype(np.float16)
    r1 = n2 - n0.astype(np.float64)
    n1 = r1.astype(np.float16)
    nn2 = (r1 - n1.astype(np.float64)).astype(np.float16)
    return xs, hi, lo, (n0, n1, nn2), n2, R * sc, sc


def test_two_product_scores_bound_the_true_distance():
    rng = np.random.default_rng(7)
    n, d = 600, 50
    centres = rng.normal(0.0, 6.0, size=(8, d))
    emb = centres[rng.integers(0, 8, size=n)] + rng.normal(0.0, 1.0, size=(n, d))
    emb[:, 0] += 55.0                                            # an uncentred leading component, as alona's PCA has
    xs, hi, lo, (n0, n1, nn2), n2, Rs, sc = _operands(emb)
    K = 2 * d + 3
    kt = (K + 31) // 32 * 32
    ksteps16 = (K + 15) // 16 * 16
    kappa = 1.25 * ksteps16 * 2.0 ** -24 + 2.0 ** -19
    assert kt == 128 and ksteps16 == 112
    # the GEMM: A[q] = [q_hi | q_hi | 1 1 1],  B[x] = [-2 x_hi | -2 x_lo | n0 n1 n2], products exact in FP32, FP32 accumulation
    A = np.concatenate([hi, hi, np.ones((n, 3), np.float16)], axis=1).astype(np.float32)
    B = np.concatenate([-2.0 * hi.astype(np.float32), -2.0 * lo.astype(np.float32),
                        np.stack([n0, n1, nn2], axis=1).astype(np.float32)], axis=1)
    s = np.zeros((n, n), np.float32)
    for j in range(A.shape[1]):                                  # one rounding per added product, like a sequential pipe
        s = (s + A[:, j:j + 1] * B[:, j][None, :]).astype(np.float32)
    s = s.astype(np.float64)
    qh = hi.astype(np.float64)
    qh2 = (qh * qh).sum(axis=1)
    delta = np.sqrt(((xs - qh) ** 2).sum(axis=1))                # |q - q_hi|, scaled units
    qnorm = np.sqrt(n2)
    eps = kappa * (qnorm * (1.0 + 2.0 ** -10) + Rs) ** 2         # per query
    true = np.sqrt(((xs[:, None, :] - xs[None, :, :]) ** 2).sum(axis=2))     # |x - q| for every pair, scaled units
    # distance to the perturbed query, as the GEMM sees it
    d_hi2 = s + qh2[:, None]
    exact_hi2 = ((xs[None, :, :] - qh[:, None, :]) ** 2).sum(axis=2)
    assert np.max(np.abs(d_hi2 - exact_hi2) / eps[:, None]) < 0.5            # the guard of ab_knn_tc_run trips at kappa / 2
    lower = np.sqrt(np.maximum(d_hi2 - eps[:, None], 0.0)) - delta[:, None]
    upper = np.sqrt(d_hi2 + eps[:, None]) + delta[:, None]
    assert np.all(lower <= true * (1.0 + 1e-12))
    assert np.all(true <= upper * (1.0 + 1e-12))
    # what the dropped product costs: delta is ~2^-12 |q|, far below the neighbour distances it is compared with
    assert np.max(delta / qnorm) < 2.0 ** -11
    k = 20
    dk = np.sort(true, axis=1)[:, k]                             # k-th neighbour (column 0 is the query itself)
    assert np.median(2.0 * delta * dk / dk ** 2) < 2e-3


def test_certificate_excludes_every_unlisted_point():
    """Lists of the m smallest GROUP minima (groups of 8 consecutive points), threshold tau = the m-th: when the
    certificate holds, no point outside the listed groups - and no point of a listed group whose score is at or above
    tau - can be among the k nearest."""
    rng = np.random.default_rng(11)
    n, d, k, m, grp = 640, 24, 10, 32, 8
    emb = rng.normal(0.0, 1.0, size=(n, d)) * np.linspace(4.0, 0.5, d)[None, :]
    emb[:, 0] += 30.0
    xs, hi, lo, norms, n2, Rs, sc = _operands(emb)
    K = 2 * d + 3
    kappa = 1.25 * ((K + 15) // 16 * 16) * 2.0 ** -24 + 2.0 ** -19
    qh = hi.astype(np.float64)
    xt = hi.astype(np.float64) + lo.astype(np.float64)
    s = n2[None, :] - 2.0 * qh @ xt.T                            # (exact arithmetic stands in for the FP32 pipe here)
    qh2 = (qh * qh).sum(axis=1)
    delta = np.sqrt(((xs - qh) ** 2).sum(axis=1))
    eps = kappa * (np.sqrt(n2) * (1.0 + 2.0 ** -10) + Rs) ** 2
    true2 = ((xs[:, None, :] - xs[None, :, :]) ** 2).sum(axis=2)
    gmin = s.reshape(n, n // grp, grp).min(axis=2)
    certified = 0
    for q in range(n):
        order = np.argsort(gmin[q], kind='stable')
        tau = gmin[q, order[m - 1]]
        kept = np.zeros(n, bool)
        for g in order[:m]:
            seg = slice(g * grp, (g + 1) * grp)
            kept[seg] = s[q, seg] < tau                          # pair masks keep a superset of this
            kept[g * grp + int(np.argmin(s[q, seg]))] = True
        dk = np.sort(true2[q, kept])[k - 1] if kept.sum() >= k else np.inf
        t2 = tau + qh2[q] - eps[q]
        if t2 > dk and np.sqrt(t2) - delta[q] > np.sqrt(dk):
            certified += 1
            assert np.all(true2[q, ~kept] > dk)                  # nothing left out can enter the top k
    assert certified > 0.9 * n
