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

from .alonabase import AlonaBase
from .engine import get_engine, DeviceCSR
from .log import log_debug, log_error, log_info


class CountMatrix:
    """Raw counts in the layout the GPU path consumes: scipy CSR, rows = cells, cols = genes."""

    def __init__(self, csr, genes=None, cells=None):
        csr = sp.csr_matrix(csr)
        if csr.nnz and csr.data.min() < 0:
            log_error('Negative values detected in the count matrix.')
        # canonical CSR: the kernels take "one strictly positive entry per (cell, gene), columns ascending" for
        # granted (an explicitly stored zero would otherwise pass the empty-cell checks)
        csr.sum_duplicates()
        csr.eliminate_zeros()
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


class DeviceCounts:
    """`self.data` of the drop-in once a filter has touched it: the raw counts as a device-resident cell-major
    CSR plus the gene / cell labels the reference keeps on the DataFrame axes.  `.shape` is (genes, cells) like
    the reference's frame; `.to_dataframe()` rebuilds that frame for small inputs."""

    def __init__(self, engine, csr, genes, cells):
        self.engine = engine
        self.csr = csr                  # DeviceCSR, rows = cells
        self.index = pd.Index(genes)
        self.columns = pd.Index(cells)

    @classmethod
    def from_any(cls, data):
        if isinstance(data, DeviceCounts):
            return data
        if isinstance(data, pd.DataFrame):
            cm = CountMatrix.from_dataframe(data)
        elif isinstance(data, CountMatrix):
            cm = data
        elif sp.issparse(data):
            cm = CountMatrix(sp.csr_matrix(data.T))
        else:
            log_error('unsupported count container %s' % type(data))
        eng = get_engine()
        n = cm.csr.shape[0]
        lo, hi = (n * eng.rank) // eng.world, (n * (eng.rank + 1)) // eng.world       # = pipeline.shard_bounds
        return cls(eng, eng.upload_csr(cm.csr[lo:hi], row_begin=lo, n_total=n), cm.genes, cm.cells)

    @property
    def shape(self):
        return (self.csr.n_genes, self.csr.n_total)

    @property
    def local_columns(self):
        """Labels of the cells this rank holds (all of them on one GPU)."""
        return self.columns[self.csr.row_begin:self.csr.row_begin + self.csr.n_local]

    def to_dataframe(self, max_bytes=8 << 30):
        g, n = self.shape
        if g * n * 8 > max_bytes:
            raise MemoryError('dense count frame would need %.1f GB' % (g * n * 8 / 2 ** 30))
        if self.engine.world > 1:
            raise RuntimeError('to_dataframe() is a single-GPU convenience')
        c = self.csr
        m = sp.csr_matrix((c.val.cpu().numpy().astype(np.int64), c.col.cpu().numpy(), c.rowptr.cpu().numpy()),
                          shape=(n, g))
        return pd.DataFrame(m.T.toarray(), index=self.index, columns=self.columns)

    def gsum(self, x):
        """Sum of a per-rank host integer over the ranks (decisions that lead to collectives must agree)."""
        if self.engine.world == 1:
            return int(x)
        import torch.distributed as dist
        parts = [None] * self.engine.world
        dist.all_gather_object(parts, int(x))
        return int(sum(parts))

    def subset(self, keep_cells=None, keep_genes=None):
        """Rows (cells) and columns (genes) selected by boolean masks, compacted on the device."""
        g, n = self.shape
        gene_map = None
        genes = self.index
        if keep_genes is not None:
            keep_genes = np.asarray(keep_genes, dtype=bool)
            gene_map = np.where(keep_genes, np.cumsum(keep_genes) - 1, -1).astype(np.int32)
            genes = self.index[keep_genes]
        eng = self.engine
        out, _ = eng.filter_csr(self.csr, keep_cells, gene_map, n_genes_out=len(genes))
        if eng.world == 1:
            cells = self.columns if keep_cells is None else self.columns[np.asarray(keep_cells, dtype=bool)]
            return DeviceCounts(eng, out, genes, cells)
        # keep_cells is this rank's mask: the global mask decides the labels, the kept rows are re-balanced
        import torch.distributed as dist
        from .pipeline import reshard_csr
        local = np.ones(self.csr.n_local, dtype=bool) if keep_cells is None else np.asarray(keep_cells, dtype=bool)
        masks = [None] * eng.world
        dist.all_gather_object(masks, local)
        cells = self.columns[np.concatenate(masks)]
        rowptr, col, val, row_begin, n_total = reshard_csr(out.rowptr, out.col, out.val, eng.rank, eng.world)
        return DeviceCounts(eng, DeviceCSR(rowptr, col, val, len(genes), row_begin, n_total), genes, cells)


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


class AlonaCell(AlonaBase):
    """Cell data class (only the normalisation stage is in scope; alona/cell.py:30)."""

    def __init__(self):
        self.data = None
        self.data_norm = None
        self.data_ERCC = None
        self.low_quality_cells = None
        super().__init__()

    # ---- N3: upstream filters ----------------------------------------------------------
    def qc_metrics(self, rRNA_genes=()):
        """Per-cell QC metrics of find_low_quality_cells (alona/cell.py:306-312) as a DataFrame indexed by cell:
        reads_per_cell, no_genes_det, perc_rRNA, perc_mt (the MinCovDet outlier model that consumes them, :313-329,
        is outside the accelerated path)."""
        self.data = DeviceCounts.from_any(self.data)
        d = self.data
        is_mt = d.index.astype(str).str.contains('^mt-', regex=True, case=False)          # cell.py:309
        is_r = d.index.isin(list(rRNA_genes))                                              # cell.py:308
        flags = is_mt.astype(np.uint8) | (is_r.astype(np.uint8) << 1)
        q = d.engine.qc_metrics(d.csr, flags)
        reads = q['reads_per_cell'].cpu().numpy()
        sets = q['set_reads'].cpu().numpy()
        with np.errstate(invalid='ignore', divide='ignore'):
            return pd.DataFrame({'reads_per_cell': reads, 'no_genes_det': q['genes_det'].cpu().numpy(),
                                 'perc_rRNA': sets[1] / reads * 100, 'perc_mt': sets[0] / reads * 100},
                                index=d.local_columns)       # this rank's cells

    def remove_empty(self):
        """Removes empty cells and genes - with the reference's own comparisons (alona/cell.py:64-81): the zero
        count of a cell is compared with the number of CELLS and that of a gene with the number of GENES (the
        totals are swapped in the reference, :72/:77; the drop-in reproduces what it does)."""
        self.data = DeviceCounts.from_any(self.data)
        d = self.data
        total_genes, total_cells = d.shape
        q = d.engine.qc_metrics(d.csr)
        cells = total_genes - q['genes_det'].cpu().numpy().astype(np.int64)       # zero entries per cell
        genes = total_cells - q['gene_cells'].cpu().numpy()                       # zero entries per gene
        keep_cells = keep_genes = None
        n_empty = d.gsum(np.sum(cells == total_cells))
        if n_empty > 0:
            log_info('%s empty cells will be removed' % n_empty)
            keep_cells = np.logical_not(cells == total_cells)
        if np.sum(genes == total_genes) > 0:
            log_info('%s empty genes will be removed' % np.sum(genes == total_genes))
            keep_genes = np.logical_not(genes == total_genes)
        if keep_cells is not None or keep_genes is not None:
            self.data = d.subset(keep_cells, keep_genes)
        log_info('Current dimensions: %s' % str(self.data.shape))

    def remove_mito(self):
        """Remove mitochondrial genes (alona/cell.py:83-94)."""
        if self.params['remove_mito'] == 'yes':
            self.data = DeviceCounts.from_any(self.data)
            mt_count = self.data.index.astype(str).str.contains('^mt-', regex=True, case=False)
            if np.sum(mt_count) > 0:
                log_info('detected and removed %s mitochondrial gene(s)' % np.sum(mt_count))
                self.data = self.data.subset(None, np.logical_not(mt_count))

    def simple_filters(self):
        """Removes cells with too few reads (`--minreads`) and underexpressed genes (`--minexpgenes`, a fraction
        of the cells if it has a fractional part, else a number of cells) - alona/cell.py:141-173."""
        self.data = DeviceCounts.from_any(self.data)
        if self.params['dataformat'] == 'raw':
            min_reads = self.params['minreads']
            d = self.data
            cell_counts = d.engine.qc_metrics(d.csr)['reads_per_cell'].cpu().numpy()
            log_info('Keeping %s out of %s cells' % (d.gsum(np.sum(cell_counts > min_reads)), d.shape[1]))
            self.data = d.subset(cell_counts > min_reads, None)
            if self.data.shape[1] < 100:
                log_error(msg='After removing cells with < %s reads, less than 100 cells remain. Please adjust '
                              'filtering stringency with --minreads' % min_reads)
        if self.params['minexpgenes'] > 0:
            log_debug('Filtering genes based on --minexpgenes')
            thres = float(self.params['minexpgenes'])
            d = self.data
            counts = d.engine.qc_metrics(d.csr)['gene_cells'].cpu().numpy()
            if thres.is_integer():
                genes_expressed = counts
            else:
                genes_expressed = counts / d.shape[1]                             # sum(x > 0)/len(x), cell.py:168
            log_info('Removing %s genes.' % '{0:,g}'.format(np.sum(genes_expressed <= thres)))
            self.data = d.subset(None, genes_expressed > thres)

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
        elif isinstance(data, DeviceCounts):
            if remove_low_quality and self.low_quality_cells is not None and len(self.low_quality_cells):
                data = data.subset(~np.isin(np.asarray(data.local_columns), np.asarray(self.low_quality_cells)), None)
            eng = data.engine
            if int((data.csr.rowptr[1:] == data.csr.rowptr[:-1]).sum().item()):
                log_error('normalize(): cells without any count must be removed first (alona/cell.py:64-81).')
            libsize, gsum, gsumsq = eng.norm_moments(data.csr)
            log_debug('Finished normalize()')
            return NormalizedMatrix(eng, data.csr, libsize, gsum, gsumsq, np.asarray(data.index), np.asarray(data.columns))
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
