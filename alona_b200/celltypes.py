"""AlonaCellTypePred — the per-cluster expression summaries of the drop-in (mirror of
alona/celltypes.py:42-68), computed from the device-resident counts instead of the dense
genes x cells `data_norm` frame.

`median_exp(data_norm)` and `mean_exp()` keep the reference's names, arguments, return value and artefacts
(`csvs/median_exp.tsv`, `csvs/mean_exp.tsv`, alona/constants.py:32-33): a genes x clusters DataFrame whose columns
are the cluster ids of `self.clusters_targets`, preceded by the `desc` annotation column when `self.anno` is a
DataFrame.  `cluster_stats()` additionally returns what the one-way linear model of find_markers needs
(alona/find_markers.py:75-111): cluster means (its coefficients), cluster sizes and the per-gene residual variance.
The cell-type scoring and the pairwise t-tests that consume these tables are unchanged CPU code outside the path.
"""
import numpy as np
import pandas as pd

from .clustering import AlonaClustering
from .constants import OUTPUT
from .log import log_debug


class AlonaCellTypePred(AlonaClustering):
    """Cell type prediction methods (only the expression summaries are in scope; alona/celltypes.py:24)."""

    def __init__(self):
        self.markers = None
        self.marker_freq = None
        self.res_pred = None
        self.anno = getattr(self, 'anno', None)
        super().__init__()

    # one pass over the counts per table; `data_norm` is the NormalizedMatrix handle normalize() returned
    def _cluster_tables(self, data_norm, want_median):
        eng = data_norm.engine
        clust = np.asarray(self.leiden_cl)
        ids = np.unique(clust)
        lut = {c: i for i, c in enumerate(ids)}
        lo = data_norm.counts.row_begin
        local = np.array([lut[c] for c in clust[lo:lo + data_norm.counts.n_local]], dtype=np.int32)
        sizes = np.array([(clust == c).sum() for c in ids], dtype=np.int64)
        mom = eng.cluster_moments(data_norm.counts, data_norm.libsize, local, len(ids))
        med = None
        if want_median:
            med, _ = eng.cluster_median(data_norm.counts, data_norm.libsize, local, len(ids), nnz=mom['nnz'])
        return ids, sizes, mom, med

    def _finish(self, table, ids, data_norm, fn):
        ret = pd.DataFrame(table.T, index=data_norm.index, columns=ids)
        ret = ret.iloc[:, ret.columns.isin(self.clusters_targets)]
        if isinstance(self.anno, pd.DataFrame):
            ret = pd.concat([self.anno['desc'], ret], axis=1)
        self._write(lambda path: ret.to_csv(path, header=True, sep='\t'), fn)       # rank 0 only
        return ret

    def median_exp(self, data_norm):
        """ Represent each cluster with median gene expression (alona/celltypes.py:42-54). """
        log_debug('median_exp() Computing median expression per cluster')
        ids, _, _, med = self._cluster_tables(data_norm, True)
        ret = self._finish(med.cpu().numpy(), ids, data_norm, OUTPUT['FILENAME_MEDIAN_EXP'])
        log_debug('median_exp() finished')
        return ret

    def mean_exp(self):
        """ Represent each cluster with mean gene expression (alona/celltypes.py:56-68). """
        log_debug('mean_exp() Computing mean expression per cluster')
        ids, sizes, mom, _ = self._cluster_tables(self.data_norm, False)
        mean = mom['sum'].cpu().numpy() / sizes[:, None]
        ret = self._finish(mean, ids, self.data_norm, OUTPUT['FILENAME_MEAN_EXP'])
        log_debug('mean_exp() finished')
        return ret

    def cluster_stats(self):
        """Sufficient statistics of find_markers' model `~ 0 + C(cl)` on the clusters of `clusters_targets`
        (alona/find_markers.py:80-97): coefficients (= cluster means, clusters x genes), cluster sizes, residual
        degrees of freedom and the residual variance per gene, sigma2 = (sum x^2 - sum_c n_c mean_c^2)/(n - C)."""
        ids, sizes, mom, _ = self._cluster_tables(self.data_norm, False)
        keep = np.isin(ids, self.clusters_targets)
        s = mom['sum'].cpu().numpy()[keep]
        q = mom['sumsq'].cpu().numpy()[keep]
        n_c = sizes[keep].astype(np.float64)
        coef = s / n_c[:, None]
        resid_df = int(n_c.sum()) - int(keep.sum())
        rss = q.sum(axis=0) - (n_c[:, None] * coef * coef).sum(axis=0)
        return {'clusters': ids[keep], 'coef': pd.DataFrame(coef, index=ids[keep], columns=self.data_norm.index),
                'sizes': sizes[keep], 'resid_df': resid_df,
                'sigma2': pd.Series(rss / resid_df, index=self.data_norm.index)}
