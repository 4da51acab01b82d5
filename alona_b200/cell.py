"""AlonaCell — the normalisation stage of the drop-in (mirror of alona/cell.py:224-255).

`normalize()` keeps the reference's signature and semantics for raw UMI counts
(library-size scaling x 10,000, log2(x+1), alona/cell.py:241-243) but never builds the dense
genes x cells float64 frame: it returns a NormalizedMatrix, a device-resident handle holding
the cell-major CSR counts, the exact library sizes and the per-gene moments that the fused
kernel (ab_norm_moments) produced in the same pass.  `.to_dataframe()` materialises the
reference's DataFrame on demand for small inputs (downstream consumers).
"""
import os

import numpy as np
import pandas as pd
import scipy.sparse as sp
import torch

from .engine import get_engine, DeviceCSR
from .log import log_debug, log_error


class CountMatrix:
    """Raw counts in the layout the GPU path consumes: scipy CSR, rows = cells, cols = genes."""

    def __init__(self, csr, genes=None, cells=None):
        csr = sp.csr_matrix(csr)
        csr.sort_indices()
        self.csr = csr
        self.genes = np.asarray(genes if genes is not None else np.arange(csr.shape[1]))
        self.cells = np.asarray(cells if cells is not None else np.arange(csr.shape[0]))

    @classmethod
    def from_dataframe(cls, df):
        """df: genes x cells integer DataFrame, the reference's in-memory layout (cell.py:381-383)."""
        vals = df.values
        if not np.issubdtype(vals.dtype, np.integer):
            if np.any(vals != np.floor(vals)):
                log_error('Non-count values detected in data matrix while data type is set to "raw".')
            vals = vals.astype(np.int64)
        return cls(sp.csr_matrix(vals.T), np.asarray(df.index), np.asarray(df.columns))


class NormalizedMatrix:
    """What `self.data_norm` is in the drop-in: genes x cells, log2-normalised, on the device."""

    def __init__(self, engine, counts, libsize, gsum, gsumsq, genes, cells):
        self.engine = engine
        self.counts = counts            # DeviceCSR (this rank's cell rows)
        self.libsize = libsize          # int64 [n_local]
        self.gsum, self.gsumsq = gsum, gsumsq   # float64 [G], already summed over all ranks
        self.index = pd.Index(genes)    # genes   (DataFrame.index of the reference)
        self.columns = pd.Index(cells)  # cells   (DataFrame.columns of the reference)

    @property
    def shape(self):
        return (self.counts.n_genes, self.counts.n_total)

    def mean(self, axis=1):
        assert axis == 1
        return pd.Series((self.gsum / self.counts.n_total).cpu().numpy(), index=self.index)

    def values_csr(self):
        """Host scipy CSR (cells x genes) of the normalised values of this shard."""
        c = self.counts
        eng = self.engine
        g2h = torch.arange(c.n_genes, dtype=torch.int32, device=eng.device)
        full = eng.hvg_extract(c, self.libsize, g2h, c.n_genes)
        return sp.csr_matrix((full.val.cpu().numpy(), full.col.cpu().numpy(), full.rowptr.cpu().numpy()),
                             shape=(c.n_local, c.n_genes))

    def to_dataframe(self, max_bytes=8 << 30):
        """Dense genes x cells float64 DataFrame — the reference's `data_norm` (cell.py:243)."""
        g, n = self.shape
        if g * n * 8 > max_bytes:
            raise MemoryError('dense data_norm would need %.1f GB' % (g * n * 8 / 2 ** 30))
        if self.engine.world > 1:
            raise RuntimeError('to_dataframe() is a single-GPU convenience')
        dense = self.values_csr().T.toarray()
        return pd.DataFrame(dense, index=self.index, columns=self.columns)


class AlonaCell:
    """Cell data class (only the normalisation stage is in scope; alona/cell.py:30)."""

    def __init__(self):
        self.data = None
        self.data_norm = None
        self.data_ERCC = None
        self.low_quality_cells = None
        self.params = getattr(self, 'params', None)

    def get_wd(self):
        return self.params['output_directory']

    def normalize(self, data, fn_out='', input_type='raw', mrnafull=False):
        """Normalizes gene expression values (alona/cell.py:224-255).

        data: genes x cells DataFrame (reference layout), a CountMatrix, or a scipy sparse
        genes x cells matrix.  Returns a NormalizedMatrix.
        """
        log_debug('Inside normalize()')
        remove_low_quality = self.params['qc_auto']
        if mrnafull or input_type != 'raw':
            # cell.py:244-250 (RPKM / pre-normalised input) is outside the B200 hot path
            log_error('alona_b200 accelerates --dataformat raw without --mrnafull only '
                      '(alona/cell.py:240-243); other input types are out of scope.')
        if isinstance(data, pd.DataFrame):
            if remove_low_quality and self.low_quality_cells is not None and len(self.low_quality_cells):
                data = data.drop(self.low_quality_cells, axis=1, errors='ignore')     # cell.py:232-239
            cm = CountMatrix.from_dataframe(data)
        elif isinstance(data, CountMatrix):
            cm = data
            if remove_low_quality and self.low_quality_cells is not None and len(self.low_quality_cells):
                keep = ~np.isin(cm.cells, np.asarray(self.low_quality_cells))
                cm = CountMatrix(cm.csr[keep], cm.genes, cm.cells[keep])
        elif sp.issparse(data):
            cm = CountMatrix(sp.csr_matrix(data.T))
        else:
            log_error('normalize(): unsupported input type %s' % type(data))
        if cm.csr.shape[0] and np.any(np.diff(cm.csr.indptr) == 0):
            log_error('normalize(): cells without any count must be removed first (alona/cell.py:64-81).')
        eng = get_engine()
        n_total = cm.csr.shape[0]
        lo = (n_total * eng.rank) // eng.world
        hi = (n_total * (eng.rank + 1)) // eng.world
        shard = eng.upload_csr(cm.csr[lo:hi], row_begin=lo, n_total=n_total)
        libsize, gsum, gsumsq = eng.norm_moments(shard)
        out = NormalizedMatrix(eng, shard, libsize, gsum, gsumsq, cm.genes, cm.cells)
        log_debug('Finished normalize()')
        return out
