"""AlonaHighlyVariableGenes — mirror of alona/hvg.py:30-63,133-167 (Seurat flavour only).

The statistics run on the device from the per-gene moments the normalisation kernel already
produced (ab_hvg_seurat); the constructor/`find()` surface is the reference's.
"""
import numpy as np

from .engine import get_engine
from .log import log_debug, log_error


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
        if self.hvg_method in ('Brennecke2013', 'scran', 'Chen2016', 'M3Drop_smartseq2', 'M3Drop_UMI'):
            # alona/hvg.py:65-131,169-519: ERCC/GLM based selectors, not on the B200 hot path
            log_error('hvg method "%s" is outside the scope of alona_b200 (seurat only).' % self.hvg_method)
        log_error('Unknown hvg method specified.')

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
        import pandas as pd
        self.top_hvg = pd.Series(z.cpu().numpy()[idx], index=np.asarray(dn.index)[idx])
        log_debug('Finishing hvg_seurat()')
        return np.array(self.top_hvg.index)
