"""AlonaHighlyVariableGenes — mirror of alona/hvg.py:30-63 (dispatch), :133-167 (seurat), :65-131 (Brennecke2013)
and :169-225 (scran).

Every selector works from per-gene moments that a kernel produced in one pass over the device-resident counts
(ab_norm_moments / ab_norm_moments_linear); the genes x cells frame of the reference is never built.  Seurat's
statistics run on the device (ab_hvg_seurat); the Brennecke and scran selectors finish with O(G) host statistics
(a two-parameter GLM, a chi-square test, a degree-4 polynomial fit) exactly as the reference does on its per-gene
vectors.  The constructor/`find()` surface is the reference's.
"""
import numpy as np
import pandas as pd

from .engine import get_engine
from .log import log_debug, log_error


def p_adjust_bh(p):
    """Benjamini-Hochberg adjustment (alona/stats.py:12-20)."""
    p = np.asarray(p, dtype=np.float64)
    by_descend = p.argsort()[::-1]
    by_orig = by_descend.argsort()
    steps = float(len(p)) / np.arange(float(len(p)), 0, -1)
    q = np.minimum(1, np.minimum.accumulate(steps * p[by_descend]))
    return q[by_orig]


def gamma_glm_identity(X, y, max_iter=100, tol=0.1 ** 5):
    """The fit hvg_brennecke obtains from alona/glm (glm.py:180-227 with families.Gamma as modified there:
    identity mean, d_inv_link = mu, variance = mu^2): Fisher scoring started at (mean(y), 0, ...), stopped when the
    relative change of the Gamma deviance drops below tol.  With those family functions the working weights are 1
    and the score is -X^T((y - mu)/mu).  Returns the coefficient vector."""
    X = np.asarray(X, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    coef = np.zeros(X.shape[1])
    coef[0] = np.mean(y)
    deviance = np.inf
    for _ in range(max_iter):
        mu = np.dot(X, coef)
        dmu = mu
        var = mu * mu
        dbeta = -np.sum(X * ((y - mu) * (dmu / var)).reshape(-1, 1), axis=0)
        ddbeta = np.dot(X.T, X * (dmu ** 2 / var).reshape(-1, 1))
        coef = coef - np.linalg.solve(ddbeta, dbeta)
        previous = deviance
        deviance = 2 * np.sum((y - mu) / mu - np.log(y / mu))
        if previous != np.inf and np.abs((deviance - previous) / previous) < tol:
            break
    return coef


def mean_var_from_sums(s1, s2, n):
    """Per-gene mean and ddof-1 variance (pandas' mean/var over the cells, zeros included) from sum and sum of squares."""
    s1 = np.asarray(s1, dtype=np.float64)
    s2 = np.asarray(s2, dtype=np.float64)
    mean = s1 / n
    with np.errstate(invalid='ignore', divide='ignore'):
        var = (s2 - s1 * s1 / n) / (n - 1)
    return mean, np.maximum(var, 0.0)


def brennecke_select(ysum, ysumsq, m, hvg_n, fdr=0.1, minBiolDisp=0.5):
    """alona/hvg.py:88-131 on the per-gene sums of the linear-scale values; returns (gene indices, diagnostics)."""
    import scipy.stats
    meansGenes, varsGenes = mean_var_from_sums(ysum, ysumsq, m)
    meansSp, varsSp = meansGenes, varsGenes                      # find() never passes spikes: norm_ERCC = data_norm (:80-81)
    with np.errstate(invalid='ignore', divide='ignore'):
        cv2Sp = varsSp / meansSp ** 2
    minMeanForFit = np.quantile(meansSp[cv2Sp > 0.3], 0.8)
    useForFit = meansSp >= minMeanForFit
    if np.sum(useForFit) < 20:                                   # :100-104 (`cv2All` is the variance there)
        minMeanForFit = np.quantile(meansGenes[varsGenes > 0.3], 0.8)
        useForFit = meansSp >= minMeanForFit
    X = np.stack([np.ones(int(useForFit.sum())), 1.0 / meansSp[useForFit]], axis=1)
    coef = gamma_glm_identity(X, cv2Sp[useForFit])
    a0, a1 = coef[0], coef[1]
    psia1theta = a1
    minBiolDisp = minBiolDisp ** 2
    cv2th = a0 + minBiolDisp + a0 * minBiolDisp
    testDenom = (meansGenes * psia1theta + (meansGenes ** 2) * cv2th) / (1 + cv2th / m)
    with np.errstate(invalid='ignore', divide='ignore'):
        q = varsGenes * (m - 1) / testDenom
    p = 1 - scipy.stats.chi2.cdf(q, m - 1)
    padj = p_adjust_bh(p)
    return np.nonzero(padj < fdr)[0][:hvg_n], {'a0': a0, 'a1': a1, 'pvalue': p, 'padj': padj}


def scran_select(means_tech, vars_tech, gsum, gsumsq, n, hvg_n):
    """alona/hvg.py:196-225: degree-4 polynomial fit of log(var) on mean over the spikes, technical variance predicted
    for every gene and subtracted; returns (gene indices sorted by the biological remainder, the remainders)."""
    from sklearn.linear_model import LinearRegression
    from sklearn.preprocessing import PolynomialFeatures
    to_fit = np.log(np.asarray(vars_tech, dtype=np.float64))
    order = sorted(zip(np.asarray(means_tech, dtype=np.float64), to_fit))
    x = np.array([a for a, _ in order])
    y = np.array([b for _, b in order])
    poly_reg = PolynomialFeatures(degree=4)
    pol_reg = LinearRegression()
    pol_reg.fit(poly_reg.fit_transform(x.reshape(-1, 1)), y)
    bio_means, vars_bio_total = mean_var_from_sums(gsum, gsumsq, n)
    vars_pred = pol_reg.predict(poly_reg.fit_transform(bio_means.reshape(-1, 1)))
    vars_bio_bio = pd.Series(vars_bio_total - vars_pred, index=np.arange(bio_means.size))
    vars_bio_bio = vars_bio_bio.sort_values(ascending=False)
    return vars_bio_bio.head(hvg_n).index.values, vars_bio_bio


class AlonaHighlyVariableGenes():
    """HVG class (same constructor as alona/hvg.py:35-39)."""

    def __init__(self, hvg_method, hvg_n, data_norm, data_ERCC=None):
        self.hvg_method = hvg_method
        self.hvg_n = hvg_n
        self.data_norm = data_norm
        self.data_ERCC = data_ERCC
        self.top_hvg = None
        self.stats = None

    def find(self):
        """ Finds HVG. Returns an array of HVG (alona/hvg.py:41-57). """
        if self.hvg_method == 'seurat':
            return self.hvg_seurat()
        if self.hvg_method == 'Brennecke2013':
            return self.hvg_brennecke()
        if self.hvg_method == 'scran':
            return self.hvg_scran()
        if self.hvg_method in ('Chen2016', 'M3Drop_smartseq2', 'M3Drop_UMI'):
            # alona/hvg.py:227-519 are outside SURVEY.md 8(f) N4 (which names :65-131 and :169-225)
            log_error('hvg method "%s" is outside the scope of alona_b200 (seurat, Brennecke2013, scran).' % self.hvg_method)
        log_error('Unknown hvg method specified.')

    def _finish_by_name(self, idx):
        """Common tail: remembers the column numbering of the selected genes (original gene order, clustering.py:104-105)
        for PCA() and returns the labels in the selector's order."""
        import torch
        dn = self.data_norm
        idx = np.asarray(idx, dtype=np.int64)
        g2h = np.full(dn.shape[0], -1, dtype=np.int32)
        cols = np.sort(idx)
        g2h[cols] = np.arange(cols.size, dtype=np.int32)
        self.gene2hvg = torch.from_numpy(g2h).to(dn.engine.device)
        self.ranked_idx = torch.from_numpy(idx.astype(np.int32)).to(dn.engine.device)
        self.stats = {'finite_z': int(idx.size), 'selected': int(idx.size), 'ties_at_cut': 0}
        labels = np.asarray(dn.index)[idx]
        self.top_hvg = pd.Series(np.zeros(idx.size), index=labels)
        return np.array(labels)

    def hvg_brennecke(self, norm_ERCC=None, fdr=0.1, minBiolDisp=0.5):
        """Brennecke et al. (2013) as alona/hvg.py:65-131 runs it: on the linear scale y = 2^x - 1, spikes = the genes
        themselves (find() never passes norm_ERCC), technical fit a0 + a1/mean by the Gamma GLM of alona/glm, chi-square
        test of every gene's variance against the fit, BH-adjusted p < fdr, first hvg_n genes in matrix order."""
        log_debug('Entering hvg_brennecke()')
        if norm_ERCC is not None:
            log_error('hvg_brennecke: explicit spike matrices are outside the scope of alona_b200.')
        dn = self.data_norm
        ys, yq = dn.engine.norm_moments_linear(dn.counts, dn.libsize)
        idx, self.brennecke = brennecke_select(ys.cpu().numpy(), yq.cpu().numpy(), dn.shape[1], self.hvg_n, fdr, minBiolDisp)
        log_debug('Finishing hvg_brennecke()')
        return self._finish_by_name(idx)

    def hvg_scran(self):
        """scran's variance decomposition as alona/hvg.py:169-225 runs it; `data_ERCC` is the NormalizedMatrix that
        normalize() returns for the spike rows (alona/cell.py:410) or a host frame of normalised spike values."""
        log_debug('Entering hvg_scran()')
        ercc = self.data_ERCC
        if ercc is None or (hasattr(ercc, 'empty') and ercc.empty) or ercc.shape[0] == 0:
            log_error('Running "--hvg scran" requires ERCC spikes in the dataset. these should begin with ERCC- '
                      'followed by numbers.')
        dn = self.data_norm
        if isinstance(ercc, pd.DataFrame):
            means_tech, vars_tech = ercc.mean(axis=1).values, ercc.var(axis=1).values
        else:
            means_tech, vars_tech = mean_var_from_sums(ercc.gsum.cpu().numpy(), ercc.gsumsq.cpu().numpy(), ercc.shape[1])
        idx, vb = scran_select(means_tech, vars_tech, dn.gsum.cpu().numpy(), dn.gsumsq.cpu().numpy(), dn.shape[1], self.hvg_n)
        self.scran = {'vars_bio': vb}
        log_debug('Finishing hvg_scran()')
        return self._finish_by_name(idx)

    @staticmethod
    def _exp_mean(mat):
        return mat.mean(axis=1)

    def hvg_seurat(self):
        """Ranked array of the top `hvg_n` gene labels (alona/hvg.py:133-167); 20 equal-width
        mean bins, dispersion = var/mean, per-bin z-score."""
        log_debug('Entering hvg_seurat()')
        dn = self.data_norm
        eng = get_engine()
        ranked, z, gene2hvg, stats = eng.hvg_seurat(dn.gsum, dn.gsumsq, dn.counts.n_total, self.hvg_n, nbins=20)
        self.ranked_idx = ranked
        self.z = z
        self.gene2hvg = gene2hvg
        st = stats.cpu().numpy()
        self.stats = {'finite_z': int(st[0]), 'selected': int(st[1]), 'ties_at_cut': int(st[2])}
        idx = ranked.cpu().numpy()
        self.top_hvg = pd.Series(z.cpu().numpy()[idx], index=np.asarray(dn.index)[idx])
        log_debug('Finishing hvg_seurat()')
        return np.array(self.top_hvg.index)
