"""AlonaClustering — drop-in mirror of the clustering front end of alona/clustering.py:45-311.

Same method names, arguments, side-effect attributes (`hvg`, `pca_components`, `nn_idx`,
`snn_graph`), artefact files (alona/constants.py:24-30) and cache-skip rules; the arithmetic
runs in hand-written sm_100a kernels behind the C-ABI (include/alona_b200.h):

    find_variable_genes()  -> ab_hvg_seurat                     (alona/clustering.py:64-80)
    PCA(out_path)          -> ab_hvg_extract_*, ab_irlb         (alona/clustering.py:82-113,126)
    knn(inp_k)             -> ab_knn                            (alona/clustering.py:179-188)
    snn(k, prune_snn)      -> ab_snn_count / ab_snn_fill        (alona/clustering.py:190-240)
    cluster()              -> knn, snn, leiden                  (alona/clustering.py:295-311)

`leiden()` is the consumer of the edge list and is unchanged (it needs leidenalg/igraph).
"""
import os

import numpy as np
import pandas as pd
import torch

from . import irlbpy
from .cell import AlonaCell
from .constants import OUTPUT
from .engine import get_engine
from .hvg import AlonaHighlyVariableGenes
from .log import log_debug, log_error, log_info, log_warning
from .pipeline import prune_threshold, shard_bounds


class AlonaClustering(AlonaCell):
    """Clustering class."""

    def __init__(self):
        self.hvg = None
        self.pca_components = None
        self.embeddings = None
        self.nn_idx = None
        self.snn_graph = None
        self.leiden_cl = None
        self.anno = None
        self.preclust = None
        self.report = {}               # parity / diagnostics counters of the last run
        super().__init__()

    # -- HVG ------------------------------------------------------------------------------
    def find_variable_genes(self):
        log_debug('Entering find_variable_genes()')
        finder = AlonaHighlyVariableGenes(hvg_method=self.params['hvg_method'], hvg_n=self.params['hvg_n'],
                                          data_norm=self.data_norm, data_ERCC=self.data_ERCC)
        self.hvg = finder.find()
        self._hvg_finder = finder
        self.report['hvg'] = finder.stats
        if isinstance(self.anno, pd.DataFrame):
            w = pd.DataFrame(self.hvg, index=self.hvg, columns=['gene']).merge(self.anno)
        else:
            w = pd.DataFrame(self.hvg)
        self._write(lambda fn: w.to_csv(fn, header=False, index=False, sep='\t'), OUTPUT['FILENAME_HVG'])
        log_debug('Finished find_variable_genes()')

    # -- PCA ------------------------------------------------------------------------------
    def PCA(self, out_path):
        """--pca irlb: truncated SVD of the (uncentred) HVG x cells log-expression matrix by IRLBA,
        scores = V.diag(s) (alona/clustering.py:103-113).
        --pca regular (alona/clustering.py:114-125): the reference centres every gene, takes the full LAPACK SVD and
        keeps the first pca_n columns of x.v = u.diag(d).  Here the same leading components come from the same
        device iteration run on the IMPLICITLY centred operator (gene means subtracted inside the products, the
        centred matrix is never formed) to a residual tolerance of 1e-10, i.e. the centred exact SVD to ~1e-8 per
        component; component signs are LAPACK's convention in the reference and arbitrary here (as for irlb)."""
        log_debug('Running PCA...')
        if self.params['pca'] not in ('irlb', 'regular'):
            log_error('--pca must be irlb or regular.')
        eng = get_engine()
        dn = self.data_norm
        n_comp = self.params['pca_n']
        finder = getattr(self, '_hvg_finder', None)
        if finder is not None and finder.gene2hvg is not None and np.array_equal(self.hvg, finder.top_hvg.index):
            gene2hvg = finder.gene2hvg
        else:
            # HVGs supplied by name: keep ORIGINAL gene order (clustering.py:104-105)
            mask = np.asarray(dn.index.isin(self.hvg))
            g2h = np.where(mask, np.cumsum(mask) - 1, -1).astype(np.int32)
            gene2hvg = torch.from_numpy(g2h).to(eng.device)
        n_hvg = int((gene2hvg >= 0).sum().item())
        self._AH = eng.hvg_extract(dn.counts, dn.libsize, gene2hvg, n_hvg)
        if self.params['pca'] == 'regular':
            # per-gene means of the HVG columns (original gene order), from the moments of the normalisation pass
            sel = torch.nonzero(gene2hvg >= 0).flatten()
            center = (dn.gsum[sel] / dn.counts.n_total).contiguous()
            lanc = irlbpy.lanczos(self._AH, nval=n_comp, tol=1e-10, maxit=1000, seed=self.params['seed'],
                                  device_result=True, gene_center=center)
        else:
            lanc = irlbpy.lanczos(self._AH, nval=n_comp, maxit=1000, seed=self.params['seed'], device_result=True)
        self._lanczos = lanc
        self.report['pca'] = {'steps': lanc['steps'], 'nmult': lanc['nmult'], 'nnz_hvg': self._AH.nnz,
                              'not_converged': lanc['not_converged']}
        bounds = shard_bounds(dn.counts.n_total, eng.world)
        rows = [bounds[r + 1] - bounds[r] for r in range(eng.world)]
        self._emb_dev = eng.allgather_rows(lanc['scores'], rows)
        self.pca_components = pd.DataFrame(self._emb_dev.cpu().numpy(), index=dn.columns)
        if out_path:
            self._write(lambda fn: self.pca_components.to_csv(path_or_buf=fn, sep=',', header=None), None, out_path)
        log_debug('Finished PCA')

    # -- kNN ------------------------------------------------------------------------------
    def knn(self, inp_k, filename=''):
        """Exact k nearest neighbours of every cell among all cells, self included; sets
        `self.nn_idx` (N,k) int64, 1-BASED, column 0 = the cell itself (clustering.py:184-187)."""
        log_debug('Performing Nearest Neighbour Search')
        eng = get_engine()
        emb = getattr(self, '_emb_dev', None)
        if emb is None or emb.shape[0] != len(self.pca_components):
            emb = torch.from_numpy(np.ascontiguousarray(np.asarray(self.pca_components, dtype=np.float64))).to(eng.device)
        n = emb.shape[0]
        bounds = shard_bounds(n, eng.world)
        lo, hi = bounds[eng.rank], bounds[eng.rank + 1]
        idx, rdist, stats = eng.knn(emb.contiguous(), int(inp_k), lo, hi, mode=int(self.params.get('knn_mode', 0)))
        rows = [bounds[r + 1] - bounds[r] for r in range(eng.world)]
        self._knn_dev = eng.allgather_rows(idx, rows)
        if eng.world > 1:
            counters = stats[:4].clone()                     # tie / fallback / duplicate / certified rows: sum over ranks
            eng.allreduce_i64(counters)
            stats = torch.cat([counters, stats[4:]])
        st = stats.cpu().numpy()
        self.report['knn'] = {'tie_rows': int(st[0]), 'exact_fallback_rows': int(st[1]), 'duplicate_rows': int(st[2]),
                              'certified_rows': int(st[3])}
        self.nn_idx = self._knn_dev.cpu().numpy().astype(np.int64) + 1
        log_debug('Finished NNS')

    # -- SNN ------------------------------------------------------------------------------
    def snn(self, k, prune_snn):
        """Shared-nearest-neighbour graph: keep (i,m) iff |N(i) n N(m)|/(2k-|N(i) n N(m)|) > prune_snn;
        sets `self.snn_graph` (DataFrame source_node,target_node; 0-based) and writes snn_graph.csv
        (clustering.py:190-240)."""
        log_debug('Computing SNN graph...')
        snn_path = self.get_wd() + OUTPUT['FILENAME_SNN_GRAPH'] if self._has_wd() else None
        if snn_path and os.path.exists(snn_path):
            log_debug('Loading SNN from file...')                  # same cache rule as clustering.py:197-201
            self.snn_graph = pd.read_csv(snn_path, header=None)
            return
        eng = get_engine()
        knn_dev = getattr(self, '_knn_dev', None)
        if knn_dev is None or knn_dev.shape != tuple(self.nn_idx.shape):
            knn_dev = torch.from_numpy((np.asarray(self.nn_idx) - 1).astype(np.int32)).to(eng.device).contiguous()
        n = knn_dev.shape[0]
        bounds = shard_bounds(n, eng.world)
        lo, hi = bounds[eng.rank], bounds[eng.rank + 1]
        vmin = prune_threshold(int(k), prune_snn)
        res = eng.snn(knn_dev, vmin, lo, hi, want_overlap=True)
        self._snn_dev = res
        counts = None
        if eng.world > 1:
            import torch.distributed as dist
            cnt = torch.tensor([res['src'].numel()], dtype=torch.int64, device=eng.device)
            allc = [torch.zeros_like(cnt) for _ in range(eng.world)]
            dist.all_gather(allc, cnt)
            counts = [int(c.item()) for c in allc]
            src = eng.allgather_rows(res['src'], counts)
            dst = eng.allgather_rows(res['dst'], counts)
            ov = eng.allgather_rows(res['overlap'], counts)
        else:
            src, dst, ov = res['src'], res['dst'], res['overlap']
        self.snn_overlap = ov.cpu().numpy()
        nonzeros, dup_rows = res['nonzeros_unpruned'], res['dup_rows']
        if eng.world > 1:
            t = torch.tensor([nonzeros, dup_rows], dtype=torch.int64, device=eng.device)
            eng.allreduce_i64(t)
            nonzeros, dup_rows = int(t[0].item()), int(t[1].item())
        self.report['snn'] = {'edges': int(src.numel()), 'v_min': vmin, 'max_walk': res['max_walk'],
                              'duplicate_rows': dup_rows, 'nonzeros_unpruned': nonzeros}
        # the reference's messages (clustering.py:231-235)
        pruned_count = nonzeros - int(src.numel())
        perc_pruned = (pruned_count / max(1, nonzeros)) * 100
        log_debug('%.2f%% (n=%s) of links pruned' % (perc_pruned, '{:,}'.format(pruned_count)))
        if perc_pruned > 80:
            log_warning('more than 80% of the edges were pruned')
        res = dict(res, dup_rows=dup_rows)
        if res['dup_rows']:
            log_warning('%d cells have an exact duplicate ranked before themselves; the reference keys such rows '
                        'on the duplicate (clustering.py:204) - reported, not emulated.' % res['dup_rows'])
        df = pd.DataFrame({'source_node': src.cpu().numpy(), 'target_node': dst.cpu().numpy()})
        if snn_path:
            # same bytes as df.to_csv(fn, header=None, index=None) (clustering.py:237-238), formatted on the device
            self._write(lambda fn: eng.write_edges_csv(src.contiguous(), dst.contiguous(), fn), None, snn_path)
        self.snn_graph = df
        log_debug('Done computing SNN.')

    # -- orchestration ----------------------------------------------------------------------
    def cluster(self):
        """kNN -> SNN -> Leiden (alona/clustering.py:295-311, de-novo branch)."""
        if isinstance(self.preclust, pd.DataFrame):
            # clustering.py:297-305: a pre-made clustering replaces kNN/SNN/Leiden; the cells must already be in
            # the order of the clustering file (the device-resident data_norm is not re-indexed)
            t = self.preclust.cell.isin(self.data_norm.columns)
            self.preclust = self.preclust.loc[t, :]
            if self.data_norm.shape[1] != self.preclust.shape[0]:
                log_error('Number of cells mismatch (data_norm and preclust)')
            if not np.array_equal(np.asarray(self.preclust['cell']), np.asarray(self.data_norm.columns)):
                log_error('--custom_clustering must list the cells in the order of the data matrix.')
            self.leiden_cl = list(self.preclust['cluster'])
            self.leiden_prep()
            return
        k = self.params['nn_k']
        self.knn(k, filename=(self.get_wd() + OUTPUT['FILENAME_KNN_map']) if self._has_wd() else '')
        self.snn(k, self.params['prune_snn'])
        self.leiden()

    def snn_edge_array(self):
        """The SNN edge list as ONE contiguous [E, 2] integer array (what igraph's Graph(edges=...) consumes
        directly): the hand-off to the unchanged Leiden step without the per-edge Python tuples of
        alona/clustering.py:277-280 (tens of millions of edges at the 1.3M-cell configuration)."""
        return np.ascontiguousarray(self.snn_graph.values)

    def leiden_prep(self):
        """ Post-clustering stuff (alona/clustering.py:242-260): clusters file, cluster sizes, clusters_targets. """
        idx = self.data_norm.columns
        self._write(lambda fn: pd.DataFrame(self.leiden_cl, index=idx).to_csv(fn, header=False, index=True),
                    OUTPUT['FILENAME_CLUSTERS_LEIDEN'])
        cl, counts = np.unique(self.leiden_cl, return_counts=True)
        ignore_clusters = self.params['ignore_small_clusters']
        self.n_clusters = np.sum(counts > ignore_clusters)
        n = len(set(self.leiden_cl))
        log_info('there are %s cell clusters (n=%s are OK)' % (n, self.n_clusters))
        self.clusters_targets = cl[counts > ignore_clusters]
        log_debug(('cluster', 'cells'))
        for item in zip(cl, counts):
            log_debug(item)
        if not getattr(self, 'cluster_colors', None):
            # plot colours (the reference draws them with utils.uniqueColors; plots are outside the scope)
            import colorsys
            m = max(1, len(self.clusters_targets))
            self.cluster_colors = ['#%02x%02x%02x' % tuple(int(255 * c) for c in colorsys.hsv_to_rgb(i / m, 0.75, 0.9))
                                   for i in range(len(self.clusters_targets))]

    def leiden(self):
        """Cluster the SNN graph using the Leiden algorithm: the unchanged leidenalg step of the reference
        (alona/clustering.py:262-293), fed from the edge ARRAY instead of a list of tuples."""
        log_debug('Running leiden clustering...')
        try:
            import igraph as ig
            import leidenalg
        except ImportError:
            log_warning('leidenalg/igraph are not installed; the SNN edge list is the hand-off point.')
            return
        res = self.params['leiden_res']
        seed = self.params['seed']
        edges = self.snn_edge_array()
        nn = np.unique(edges[:, 0]).size                              # len(set(source_node)), clustering.py:274
        g = ig.Graph()
        g.add_vertices(nn)
        g.vs['name'] = list(range(1, nn + 1))
        g.add_edges(edges)
        if self.params == 'ModularityVertexPartition':               # (sic, clustering.py:281: never true)
            part = leidenalg.ModularityVertexPartition
        else:
            part = leidenalg.RBERVertexPartition
        cl = leidenalg.find_partition(g, part, n_iterations=10, resolution_parameter=res, seed=seed)
        self.leiden_cl = cl.membership
        self.leiden_prep()
        log_debug('Leiden has finished.')

    def analysis(self):
        """The hot-path part of AlonaCell.analysis (alona/cell.py:418-435) incl. its cache rule."""
        pca_path = self.get_wd() + OUTPUT['FILENAME_PCA']
        emb_path = self.get_wd() + OUTPUT['FILENAME_EMBEDDING_PREFIX'] + self.params.get('embedding', 'tSNE') + '.csv'
        if os.path.exists(emb_path) and os.path.exists(pca_path):
            log_debug('Loading embeddings from file')
            self.pca_components = pd.read_csv(pca_path, header=None, index_col=0)
            self._emb_dev = None
        else:
            self.find_variable_genes()
            self.PCA(pca_path)
        self.cluster()

    # -- helpers ------------------------------------------------------------------------------
    def _has_wd(self):
        return bool(self.params and self.params.get('output_directory'))

    def _write(self, writer, rel, path=None):
        """Artefacts are written by rank 0 only, into <output_directory>/..."""
        if path is None:
            if not self._has_wd():
                return
            path = self.get_wd() + rel
        if get_engine().rank != 0:
            return
        os.makedirs(os.path.dirname(path), exist_ok=True)
        writer(path)
