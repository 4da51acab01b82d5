"""Device-level stage functions: torch tensors for memory and streams, every bit of arithmetic
in libalona_b200.so (hand-written sm_100a kernels) through the C-ABI.

One Engine per process / GPU.  PyTorch is plumbing only (allocation, streams,
torch.distributed bootstrap); there is no torch compute on the hot path and no CPU fallback.
"""
import ctypes
import os

import numpy as np
import torch

from . import _lib

SCALE = 10000.0          # alona/cell.py:242


class DeviceCSR:
    """Cell-major CSR shard on the device: rows = this rank's cells."""

    def __init__(self, rowptr, col, val, n_genes, row_begin=0, n_total=None):
        self.rowptr, self.col, self.val = rowptr, col, val
        self.n_local = int(rowptr.numel() - 1)
        self.n_genes = int(n_genes)
        self.row_begin = int(row_begin)
        self.n_total = int(n_total if n_total is not None else self.n_local)

    @property
    def nnz(self):
        return int(self.col.numel())


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


class Engine:
    def __init__(self, device=None, rank=0, world=1):
        if not torch.cuda.is_available():
            raise RuntimeError('alona_b200 needs a B200 GPU (sm_100a); there is no CPU fallback')
        self.lib = _lib.load()
        if device is None:
            device = torch.cuda.current_device()
        self.device_index = int(device)
        self.device = torch.device('cuda', self.device_index)
        self.rank, self.world = int(rank), int(world)
        h = ctypes.c_void_p()
        _lib.check(self.lib.ab_ctx_create(ctypes.byref(h), self.device_index, self.rank, self.world), 'ab_ctx_create')
        self.ctx = h
        self._ws = None
        self.peer = False

    def close(self):
        if self.ctx is not None:
            self.lib.ab_ctx_destroy(self.ctx)
            self.ctx = None

    # -- plumbing ---------------------------------------------------------------------
    def stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def workspace(self, nbytes):
        nbytes = int(nbytes)
        if self._ws is None or self._ws.numel() < nbytes:
            self._ws = None
            self._ws = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        return self._ws

    def launch_count(self):
        return int(self.lib.ab_launch_count(self.ctx))

    # -- measurement ------------------------------------------------------------------
    def prof_set_level(self, level):
        """0 = off, 1 = stage-level kernels, 2 = every IRLBA launch too (include/alona_b200.h)."""
        _lib.check(self.lib.ab_prof_set_level(self.ctx, int(level)), 'ab_prof_set_level')

    def prof_read(self):
        """{slot name: (device ms since the last read, event pairs)} for the slots that recorded."""
        out = {}
        for slot in range(16):
            name = self.lib.ab_prof_slot_name(slot).decode()
            if not name:
                continue
            ms = ctypes.c_double(0.0)
            cnt = ctypes.c_int64(0)
            _lib.check(self.lib.ab_prof_read(self.ctx, slot, ctypes.byref(ms), ctypes.byref(cnt)), 'ab_prof_read')
            if cnt.value:
                out[name] = (ms.value, cnt.value)
        return out

    def empty(self, shape, dtype):
        return torch.empty(shape, dtype=dtype, device=self.device)

    def zeros(self, shape, dtype):
        return torch.zeros(shape, dtype=dtype, device=self.device)

    def init_nccl(self):
        """Joins the world's NCCL communicator; the unique id travels over torch.distributed."""
        if self.world == 1:
            return
        import torch.distributed as dist
        path = nccl_library_path().encode()
        buf = ctypes.create_string_buffer(128)
        if self.rank == 0:
            _lib.check(self.lib.ab_nccl_unique_id(path, buf), 'ab_nccl_unique_id')
        obj = [buf.raw if self.rank == 0 else None]
        dist.broadcast_object_list(obj, src=0)
        idbuf = ctypes.create_string_buffer(obj[0], 128)
        _lib.check(self.lib.ab_ctx_init_nccl(self.ctx, path, idbuf), 'ab_ctx_init_nccl')
        self.init_peer()

    def init_peer(self):
        """Maps every rank's peer-exchange buffer (cudaIpc handles travel over torch.distributed) so that ab_irlb
        fuses its small all-reduces into its kernels as one-shot NVLink pushes.  All ranks agree on the outcome: if
        any rank cannot map its peers (no NVLink/P2P between the devices), every rank stays on ncclAllReduce and says
        so once on stderr."""
        import sys
        import torch.distributed as dist
        if self.world == 1 or self.world > 8 or os.environ.get('AB_NO_PEER'):
            return
        nb = int(self.lib.ab_peer_handle_bytes())
        mine = ctypes.create_string_buffer(nb)
        ok = self.lib.ab_peer_alloc(self.ctx, mine) == 0
        handles = [None] * self.world
        dist.all_gather_object(handles, mine.raw if ok else None)
        ok = ok and all(h is not None for h in handles)
        if ok:
            allh = ctypes.create_string_buffer(b''.join(handles), nb * self.world)
            ok = self.lib.ab_peer_open(self.ctx, allh) == 0
        flags = [None] * self.world
        dist.all_gather_object(flags, bool(ok))
        self.peer = all(flags)
        if not self.peer:
            os.environ['AB_IRLB_NCCL'] = '1'
            if self.rank == 0:
                sys.stderr.write('alona_b200: peer exchange buffers unavailable (%s); IRLBA uses ncclAllReduce\n'
                                 % _lib.last_error())

    def allreduce_f64(self, t):
        _lib.check(self.lib.ab_allreduce_f64(self.ctx, _ptr(t), t.numel(), self.stream()), 'ab_allreduce_f64')

    def allreduce_i64(self, t):
        _lib.check(self.lib.ab_allreduce_i64(self.ctx, _ptr(t), t.numel(), self.stream()), 'ab_allreduce_i64')

    def allgather_rows(self, local, counts):
        """All-gather of row blocks with per-rank row counts `counts` (list); returns [sum(counts), ...]."""
        if self.world == 1:
            return local
        from .pipeline import pack_rows, unpack_rows
        mx = max(counts)
        pad = pack_rows(local, mx)
        out = self.empty((self.world * mx,) + tuple(local.shape[1:]), local.dtype)
        nbytes = pad.numel() * pad.element_size()
        _lib.check(self.lib.ab_allgather_bytes(self.ctx, _ptr(pad), _ptr(out), nbytes, self.stream()),
                   'ab_allgather_bytes')
        return unpack_rows(out, counts)

    # -- uploads ----------------------------------------------------------------------
    def upload_csr(self, csr, row_begin=0, n_total=None, non_blocking=False, canonical=True):
        """scipy/numpy CSR (cells x genes, ascending column ids) -> DeviceCSR.  scipy matrices are brought to
        canonical form first (duplicates summed, explicit zeros dropped, columns sorted); negative counts are
        refused.  canonical=False uploads the arrays as they are (tests of the kernels' own guards)."""
        import scipy.sparse as sp
        if canonical and sp.issparse(csr):
            if csr.nnz and csr.data.min() < 0:
                raise ValueError('alona_b200: negative values in the count matrix')
            if not csr.has_canonical_format or (csr.nnz and csr.data.min() == 0):
                csr = sp.csr_matrix(csr, copy=True)
                csr.sum_duplicates()
                csr.eliminate_zeros()
                csr.sort_indices()
        indptr = np.ascontiguousarray(csr.indptr, dtype=np.int64)
        indices = np.ascontiguousarray(csr.indices, dtype=np.int32)
        data = np.ascontiguousarray(csr.data, dtype=np.int32)
        return DeviceCSR(torch.from_numpy(indptr).to(self.device, non_blocking=non_blocking),
                         torch.from_numpy(indices).to(self.device, non_blocking=non_blocking),
                         torch.from_numpy(data).to(self.device, non_blocking=non_blocking),
                         csr.shape[1], row_begin, n_total)

    # -- N3: QC metrics and filters (alona/cell.py:64-81, 141-173, 306-312) -----------------
    def qc_metrics(self, A, gene_flags=None, allreduce=True):
        """One pass over the counts: reads per cell, genes detected per cell, reads on up to two gene
        sets per cell, and per gene the number of cells expressing it (summed over ranks)."""
        n, G = A.n_local, A.n_genes
        reads = self.empty((n,), torch.int64)
        det = self.empty((n,), torch.int32)
        flags = None
        sets = None
        if gene_flags is not None:
            flags = torch.as_tensor(np.ascontiguousarray(gene_flags, dtype=np.uint8)).to(self.device)
            sets = self.empty((2, n), torch.int64)
        gene_cells = self.zeros((G,), torch.int64)
        _lib.check(self.lib.ab_qc_metrics(self.ctx, _ptr(A.rowptr), _ptr(A.col), _ptr(A.val), n, G, _ptr(flags),
                                          _ptr(reads), _ptr(det), _ptr(sets), _ptr(gene_cells), self.stream()),
                   'ab_qc_metrics')
        if allreduce and self.world > 1:
            self.allreduce_i64(gene_cells)
        return {'reads_per_cell': reads, 'genes_det': det, 'set_reads': sets, 'gene_cells': gene_cells}

    def filter_csr(self, A, keep_row=None, gene_map=None, n_genes_out=None):
        """Compacts the shard: rows with keep_row != 0, genes with gene_map >= 0 (renumbered by gene_map).
        Returns (DeviceCSR with row_begin = 0 / n_total = kept rows of THIS shard, source row of every kept row)."""
        n = A.n_local
        kr = None if keep_row is None else torch.as_tensor(np.ascontiguousarray(keep_row, dtype=np.uint8)).to(self.device)
        gm = None if gene_map is None else torch.as_tensor(np.ascontiguousarray(gene_map, dtype=np.int32)).to(self.device)
        ws = self.workspace(self.lib.ab_filter_ws_bytes(n))
        rows, nnz = ctypes.c_int64(0), ctypes.c_int64(0)
        _lib.check(self.lib.ab_filter_count(self.ctx, _ptr(A.rowptr), _ptr(A.col), _ptr(A.val), n, _ptr(kr), _ptr(gm),
                                            ctypes.byref(rows), ctypes.byref(nnz), _ptr(ws), ws.numel(), self.stream()),
                   'ab_filter_count')
        rowptr = self.empty((rows.value + 1,), torch.int64)
        col = self.empty((nnz.value,), torch.int32)
        val = self.empty((nnz.value,), torch.int32)
        src = self.empty((rows.value,), torch.int64)
        _lib.check(self.lib.ab_filter_fill(self.ctx, _ptr(A.rowptr), _ptr(A.col), _ptr(A.val), n, _ptr(kr), _ptr(gm),
                                           _ptr(rowptr), _ptr(col), _ptr(val), _ptr(src), _ptr(ws), ws.numel(),
                                           self.stream()), 'ab_filter_fill')
        G = int(n_genes_out) if n_genes_out is not None else (A.n_genes if gene_map is None else int(np.max(gene_map)) + 1)
        return DeviceCSR(rowptr, col, val, G, 0, rows.value), src

    # -- N2: per-cluster summaries (alona/celltypes.py:42-68, find_markers.py:75-111) ---------
    def cluster_moments(self, A, libsize, cluster, n_clusters, want_sumsq=True, allreduce=True):
        """[C][G] tables: sum of normalised values, sum of squares, number of nonzeros per (cluster, gene)."""
        cl = torch.as_tensor(np.ascontiguousarray(cluster, dtype=np.int32)).to(self.device)
        C, G = int(n_clusters), A.n_genes
        sm = self.zeros((C, G), torch.float64)
        sq = self.zeros((C, G), torch.float64) if want_sumsq else None
        nz = self.zeros((C, G), torch.int64)
        _lib.check(self.lib.ab_cluster_moments(self.ctx, _ptr(A.rowptr), _ptr(A.col), _ptr(A.val), _ptr(libsize), A.n_local,
                                               G, SCALE, _ptr(cl), C, _ptr(sm), _ptr(sq), _ptr(nz), self.stream()),
                   'ab_cluster_moments')
        if allreduce and self.world > 1:
            self.allreduce_f64(sm)
            if sq is not None:
                self.allreduce_f64(sq)
            self.allreduce_i64(nz)
        return {'sum': sm, 'sumsq': sq, 'nnz': nz, 'cluster_dev': cl}

    def cluster_median(self, A, libsize, cluster, n_clusters, nnz=None, max_batch=1 << 29):
        """[C][G] medians of the normalised values per (cluster, gene); single shard."""
        if self.world > 1:
            raise RuntimeError('cluster_median needs every cell of a cluster on one device')
        cluster = np.ascontiguousarray(cluster, dtype=np.int32)
        C, G = int(n_clusters), A.n_genes
        cl = torch.from_numpy(cluster).to(self.device)
        if nnz is None:
            nnz = self.cluster_moments(A, libsize, cluster, C, want_sumsq=False)['nnz']
        sizes = np.bincount(cluster[cluster >= 0], minlength=C)[:C].astype(np.int64)
        h_sizes = (ctypes.c_int64 * C)(*sizes.tolist())
        per = (ctypes.c_int64 * C)()
        ws = self.workspace(self.lib.ab_cluster_median_ws_bytes(G, C, 0))
        _lib.check(self.lib.ab_cluster_median_count(self.ctx, _ptr(nnz), h_sizes, G, C, per, _ptr(ws), ws.numel(),
                                                    self.stream()), 'ab_cluster_median_count')
        per = np.array(list(per), dtype=np.int64)
        batch = int(max(per.max() if C else 0, min(int(per.sum()), int(max_batch))))
        ws = self.workspace(self.lib.ab_cluster_median_ws_bytes(G, C, batch))
        med = self.empty((C, G), torch.float64)
        _lib.check(self.lib.ab_cluster_median_fill(self.ctx, _ptr(A.rowptr), _ptr(A.col), _ptr(A.val), _ptr(libsize),
                                                   A.n_local, G, SCALE, _ptr(cl), C, _ptr(nnz), h_sizes, batch, _ptr(med),
                                                   _ptr(ws), ws.numel(), self.stream()), 'ab_cluster_median_fill')
        return med, {'gathered': int(per.sum()), 'batch': batch}

    # -- K1 ---------------------------------------------------------------------------
    def norm_moments(self, A, acc=None, allreduce=True, finalize=True, libsize_out=None):
        """Library sizes + per-gene sum x / sum x^2 of the normalised values.  `acc` ([G, 4] int64 fixed-point
        accumulators, include/alona_b200.h) lets the caller chain row tiles; with finalize=False the raw accumulators
        are returned instead of the FP64 sums."""
        lib = libsize_out if libsize_out is not None else self.empty((A.n_local,), torch.int64)
        if acc is None:
            acc = self.zeros((A.n_genes, 4), torch.int64)
        _lib.check(self.lib.ab_norm_moments(self.ctx, _ptr(A.rowptr), _ptr(A.col), _ptr(A.val), A.n_local,
                                            A.n_genes, SCALE, _ptr(lib), _ptr(acc), self.stream()), 'ab_norm_moments')
        if not finalize:
            return lib, acc
        if allreduce and self.world > 1:
            self.allreduce_i64(acc)
        gsum, gsumsq = self.moments_finalize(acc)
        return lib, gsum, gsumsq

    def moments_finalize(self, acc):
        g = acc.shape[0]
        gsum = self.empty((g,), torch.float64)
        gsumsq = self.empty((g,), torch.float64)
        _lib.check(self.lib.ab_moments_finalize(self.ctx, _ptr(acc), g, _ptr(gsum), _ptr(gsumsq), self.stream()),
                   'ab_moments_finalize')
        return gsum, gsumsq

    def norm_moments_linear(self, A, libsize, allreduce=True):
        """Per-gene sum / sum of squares of y = 2^x - 1 (the linear-scale values hvg_brennecke works on)."""
        ys = self.zeros((A.n_genes,), torch.float64)
        yq = self.zeros((A.n_genes,), torch.float64)
        _lib.check(self.lib.ab_norm_moments_linear(self.ctx, _ptr(A.rowptr), _ptr(A.col), _ptr(A.val), _ptr(libsize),
                                                   A.n_local, A.n_genes, SCALE, _ptr(ys), _ptr(yq), self.stream()),
                   'ab_norm_moments_linear')
        if allreduce and self.world > 1:
            both = torch.cat([ys, yq])
            self.allreduce_f64(both)
            ys, yq = both[:A.n_genes].contiguous(), both[A.n_genes:].contiguous()
        return ys, yq

    # -- K2 ---------------------------------------------------------------------------
    def hvg_seurat(self, gsum, gsumsq, n_total, hvg_n, nbins=20):
        g = gsum.numel()
        ranked = self.empty((min(hvg_n, g),), torch.int32)
        z = self.empty((g,), torch.float64)
        gene2hvg = self.empty((g,), torch.int32)
        stats = self.empty((4,), torch.int32)
        _lib.check(self.lib.ab_hvg_seurat(self.ctx, _ptr(gsum), _ptr(gsumsq), int(n_total), g, nbins, int(hvg_n),
                                          _ptr(ranked), _ptr(z), _ptr(gene2hvg), _ptr(stats), self.stream()),
                   'ab_hvg_seurat')
        return ranked, z, gene2hvg, stats

    # -- K3 ---------------------------------------------------------------------------
    def hvg_extract(self, A, libsize, gene2hvg, n_hvg):
        ws = self.workspace(self.lib.ab_hvg_extract_ws_bytes(A.n_local))
        rowptr_h = self.empty((A.n_local + 1,), torch.int64)
        nnz = ctypes.c_int64(0)
        _lib.check(self.lib.ab_hvg_extract_count(self.ctx, _ptr(A.rowptr), _ptr(A.col), A.n_local, _ptr(gene2hvg),
                                                 _ptr(rowptr_h), ctypes.byref(nnz), _ptr(ws), ws.numel(),
                                                 self.stream()), 'ab_hvg_extract_count')
        col_h = self.empty((nnz.value,), torch.int32)
        val_h = self.empty((nnz.value,), torch.float64)
        _lib.check(self.lib.ab_hvg_extract_fill(self.ctx, _ptr(A.rowptr), _ptr(A.col), _ptr(A.val), _ptr(libsize),
                                                A.n_local, _ptr(gene2hvg), SCALE, _ptr(rowptr_h), _ptr(col_h),
                                                _ptr(val_h), self.stream()), 'ab_hvg_extract_fill')
        return DeviceCSR(rowptr_h, col_h, val_h, n_hvg, A.row_begin, A.n_total)

    # -- K4-K8 ------------------------------------------------------------------------
    def irlb(self, AH, nval, tol=1e-4, maxit=1000, v0_local=None, want_v=True, want_u=True, center=None):
        n, H = AH.n_local, AH.n_genes
        ws = self.workspace(self.lib.ab_irlb_ws_bytes_compact(n, H, nval, int(AH.val.numel())))
        scores = self.empty((n, nval), torch.float64)
        v = self.empty((n, nval), torch.float64) if want_v else None
        s = self.empty((nval,), torch.float64)
        u = self.empty((H, nval), torch.float64) if want_u else None
        info = (ctypes.c_int64 * 4)()
        _lib.check(self.lib.ab_irlb(self.ctx, _ptr(AH.rowptr), _ptr(AH.col), _ptr(AH.val), n, AH.n_total, H, int(nval),
                                    float(tol), int(maxit), _ptr(v0_local), _ptr(center), _ptr(scores), _ptr(v), _ptr(s), _ptr(u),
                                    info, _ptr(ws), ws.numel(), self.stream()), 'ab_irlb')
        return {'scores': scores, 'V': v, 's': s, 'U': u, 'steps': int(info[0]), 'nmult': int(info[1]),
                'sum_panel': int(info[2]), 'not_converged': int(info[3]) & 1, 'compact': bool(int(info[3]) & 2),
                'fused_allreduce': bool(int(info[3]) & 4)}

    # -- K9 ---------------------------------------------------------------------------
    def knn(self, emb_all, k, q_begin=0, q_end=None, mode=0):
        n, d = emb_all.shape
        if q_end is None:
            q_end = n
        nq = q_end - q_begin
        assert emb_all.dtype == torch.float64 and emb_all.is_contiguous()
        ws = self.workspace(self.lib.ab_knn_ws_bytes(n, d, nq, k))
        idx = self.empty((nq, k), torch.int32)
        rd = self.empty((nq, k), torch.float64)
        stats = self.empty((8,), torch.int64)
        _lib.check(self.lib.ab_knn(self.ctx, _ptr(emb_all), n, d, q_begin, q_end, int(k), int(mode), _ptr(idx),
                                   _ptr(rd), _ptr(stats), _ptr(ws), ws.numel(), self.stream()), 'ab_knn')
        return idx, rd, stats

    # -- K10 --------------------------------------------------------------------------
    def snn(self, knn_all, v_min, row_begin=0, row_end=None, want_overlap=True):
        n, k = knn_all.shape
        if row_end is None:
            row_end = n
        nrows = row_end - row_begin
        assert knn_all.dtype == torch.int32 and knn_all.is_contiguous()
        ws = self.workspace(self.lib.ab_snn_ws_bytes(n, k, nrows))
        rowptr = self.empty((nrows + 1,), torch.int64)
        edges = ctypes.c_int64(0)
        stats = (ctypes.c_int64 * 8)()
        _lib.check(self.lib.ab_snn_count(self.ctx, _ptr(knn_all), n, k, row_begin, row_end, int(v_min), _ptr(rowptr),
                                         ctypes.byref(edges), stats, _ptr(ws), ws.numel(), self.stream()),
                   'ab_snn_count')
        src = self.empty((edges.value,), torch.int32)
        dst = self.empty((edges.value,), torch.int32)
        ov = self.empty((edges.value,), torch.uint8) if want_overlap else None
        _lib.check(self.lib.ab_snn_fill(self.ctx, _ptr(knn_all), n, k, row_begin, row_end, int(v_min), _ptr(rowptr),
                                        _ptr(src), _ptr(dst), _ptr(ov), _ptr(ws), ws.numel(), self.stream()),
                   'ab_snn_fill')
        return {'rowptr': rowptr, 'src': src, 'dst': dst, 'overlap': ov,
                'big_rows': int(stats[0]), 'max_walk': int(stats[1]), 'dup_rows': int(stats[2]),
                'dense_rows': int(stats[3]), 'sum_walk': int(stats[4]), 'nonzeros_unpruned': int(stats[5]),
                'class_rows': [int(stats[6]), int(stats[7])]}

    # -- hand-off ---------------------------------------------------------------------
    def edges_csv_bytes(self, src, dst):
        """The edge list as the bytes alona writes to snn_graph.csv (clustering.py:237-238), formatted on the
        device; returns a pinned host uint8 tensor."""
        n = int(src.numel())
        assert src.dtype == torch.int32 and dst.dtype == torch.int32 and int(dst.numel()) == n
        ws = self.workspace(self.lib.ab_edges_csv_ws_bytes(n))
        off = self.empty((n + 1,), torch.int64)
        nbytes = ctypes.c_int64(0)
        _lib.check(self.lib.ab_edges_csv_count(self.ctx, _ptr(src), _ptr(dst), n, _ptr(off), ctypes.byref(nbytes), _ptr(ws),
                                               ws.numel(), self.stream()), 'ab_edges_csv_count')
        out = self.empty((nbytes.value,), torch.uint8)
        _lib.check(self.lib.ab_edges_csv_fill(self.ctx, _ptr(src), _ptr(dst), n, _ptr(off), _ptr(out), self.stream()),
                   'ab_edges_csv_fill')
        host = torch.empty((nbytes.value,), dtype=torch.uint8, pin_memory=True)
        host.copy_(out, non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        return host

    def write_edges_csv(self, src, dst, path):
        buf = self.edges_csv_bytes(src, dst)
        with open(path, 'wb') as fh:
            fh.write(memoryview(buf.numpy()))

    # -- synthetic data ---------------------------------------------------------------
    def synth_counts(self, prof, cell_begin, n_local, lbar, n_total=None, depth_sigma=0.35, r_disp=5.0,
                     seed=20251019):
        ncl, g = prof.shape
        assert prof.dtype == torch.float32 and prof.is_contiguous()
        ws = self.workspace(self.lib.ab_synth_ws_bytes(n_local))
        rowptr = self.empty((n_local + 1,), torch.int64)
        nnz = ctypes.c_int64(0)
        _lib.check(self.lib.ab_synth_nb_count(self.ctx, _ptr(prof), ncl, g, int(cell_begin), int(n_local), float(lbar),
                                              float(depth_sigma), float(r_disp), int(seed), _ptr(rowptr),
                                              ctypes.byref(nnz), _ptr(ws), ws.numel(), self.stream()),
                   'ab_synth_nb_count')
        col = self.empty((nnz.value,), torch.int32)
        val = self.empty((nnz.value,), torch.int32)
        _lib.check(self.lib.ab_synth_nb_fill(self.ctx, _ptr(prof), ncl, g, int(cell_begin), int(n_local), float(lbar),
                                             float(depth_sigma), float(r_disp), int(seed), _ptr(rowptr), _ptr(col),
                                             _ptr(val), self.stream()), 'ab_synth_nb_fill')
        return DeviceCSR(rowptr, col, val, g, cell_begin, n_total if n_total is not None else n_local)


def nccl_library_path():
    """The torch-bundled NCCL (nvidia-nccl-cu12 wheel), else the system one."""
    try:
        import nvidia.nccl
        base = list(nvidia.nccl.__path__)[0]
        p = os.path.join(base, 'lib', 'libnccl.so.2')
        if os.path.exists(p):
            return p
    except Exception:
        pass
    return 'libnccl.so.2'


_engine = None


def get_engine():
    """Process-wide engine, honouring torch.distributed if it has been initialised."""
    global _engine
    if _engine is None:
        rank, world = 0, 1
        try:
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized():
                rank, world = dist.get_rank(), dist.get_world_size()
        except Exception:
            pass
        _engine = Engine(rank=rank, world=world)
        if world > 1:
            _engine.init_nccl()
    return _engine
