"""The device-resident front end: normalise -> HVG -> PCA (IRLBA) -> kNN -> SNN on one rank's
shard of cells, with the exchanges SURVEY.md §8(e) lists (moments all-reduce inside
norm_moments, IRLBA all-reduces inside ab_irlb, all-gather of the embedding and of the kNN
table here).  Order of the stages = alona/cell.py:408,432-435 and alona/clustering.py:307-310.
"""
import numpy as np
import torch


def prune_threshold(k, prune_snn):
    """Smallest integer overlap kept by the reference's test  v/(k+(k-v)) > prune_snn
    (alona/clustering.py:225-227), evaluated with Python floats exactly as the reference does."""
    for v in range(0, k + 1):
        if v / (k + (k - v)) > prune_snn:
            return v
    return k + 1


def shard_bounds(n_total, world):
    return [(n_total * r) // world for r in range(world + 1)]


def pack_rows(local, n_max):
    """Pads a rank's row block to n_max rows (zeros) so that an equal-sized all-gather can carry
    ragged shards.  Works on any torch tensor (device or host)."""
    if local.shape[0] == n_max:
        return local.contiguous()
    pad = local.new_zeros((n_max,) + tuple(local.shape[1:]))
    pad[:local.shape[0]] = local
    return pad


def unpack_rows(gathered, counts):
    """Inverse of pack_rows on the gathered [world * max(counts), ...] tensor: drops the padding and
    returns the rows in rank order (= global row order, shards are contiguous row ranges)."""
    n_max = max(counts)
    if all(c == n_max for c in counts):
        return gathered
    import torch as _torch
    return _torch.cat([gathered[r * n_max: r * n_max + c] for r, c in enumerate(counts)], dim=0)


def reshard_plan(counts, world):
    """Row exchange that turns ragged contiguous shards (rank r holds `counts[r]` consecutive rows of the global
    order) into the balanced shards of shard_bounds(): plan[r][t] = (first local row, number of rows) rank r sends
    to rank t.  Pure host logic (tested over gloo)."""
    total = int(sum(counts))
    bounds = shard_bounds(total, world)
    first = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    plan = []
    for r in range(world):
        row = []
        for t in range(world):
            lo = max(int(first[r]), bounds[t])
            hi = min(int(first[r + 1]), bounds[t + 1])
            row.append((lo - int(first[r]), hi - lo) if hi > lo else (0, 0))
        plan.append(row)
    return plan, bounds


def reshard_csr(rowptr, col, val, rank, world):
    """After a filter every rank holds a ragged number of rows; downstream stages expect shard_bounds().  Moves
    whole rows between ranks with torch.distributed all-to-all (plumbing, no arithmetic): row lengths, column ids and
    counts travel as three variable-split exchanges; rows arrive in source-rank order = global order.
    Works on device (NCCL) or host (gloo) tensors.  Returns (rowptr, col, val, row_begin, n_total)."""
    import torch.distributed as dist
    n_local = int(rowptr.numel() - 1)
    counts = [None] * world
    dist.all_gather_object(counts, n_local)
    plan, bounds = reshard_plan(counts, world)
    mine = plan[rank]
    rp = rowptr.cpu()
    send_rows = [c for (_, c) in mine]
    send_nnz = [int(rp[s + c] - rp[s]) if c else 0 for (s, c) in mine]
    recv_rows = [plan[r][rank][1] for r in range(world)]
    sizes = [None] * world
    dist.all_gather_object(sizes, send_nnz)
    recv_nnz = [sizes[r][rank] for r in range(world)]
    # what this rank sends is one contiguous run of rows (targets are consecutive ranges of the global order)
    sent = [(s, c) for (s, c) in mine if c]
    s0 = sent[0][0] if sent else 0
    s1 = sent[-1][0] + sent[-1][1] if sent else 0
    lens = (rowptr[1:] - rowptr[:-1])[s0:s1].contiguous()
    a, b = int(rp[s0]), int(rp[s1])
    out_lens = lens.new_empty((sum(recv_rows),))
    out_col = col.new_empty((sum(recv_nnz),))
    out_val = val.new_empty((sum(recv_nnz),))
    dist.all_to_all_single(out_lens, lens, recv_rows, send_rows)
    dist.all_to_all_single(out_col, col[a:b].contiguous(), recv_nnz, send_nnz)
    dist.all_to_all_single(out_val, val[a:b].contiguous(), recv_nnz, send_nnz)
    new_rowptr = rowptr.new_zeros((out_lens.numel() + 1,))
    new_rowptr[1:] = torch.cumsum(out_lens, 0)
    return new_rowptr, out_col, out_val, bounds[rank], bounds[world]


def start_vector(n_total, seed):
    """The reference's start vector (alona/irlbpy/irlb.py:151-152)."""
    np.random.seed(seed)
    return np.random.randn(n_total)


class StageTimer:
    """CUDA-event timer per stage (events on the current stream)."""

    def __init__(self, enabled=True, verbose=False):
        self.enabled = enabled
        self.verbose = verbose
        self.marks = []

    def mark(self, name):
        if self.enabled and self.verbose:
            import sys
            torch.cuda.synchronize()
            sys.stderr.write('[frontend] reached %s\n' % name)
            sys.stderr.flush()
        if self.enabled:
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            self.marks.append((name, e))

    def result(self):
        if not self.enabled or len(self.marks) < 2:
            return {}
        torch.cuda.synchronize()
        out = {}
        for (_, a), (name, b) in zip(self.marks[:-1], self.marks[1:]):
            out[name] = out.get(name, 0.0) + a.elapsed_time(b)
        return out


def upload_and_normalize(eng, h_rowptr, h_col, h_val, n_genes, row_begin, n_total, tiles=16):
    """Host (pinned) CSR shard -> device, with K1 running on row tile i while tile i+1 is still crossing PCIe: the
    library sizes and the fixed-point moment accumulators (chainable across tiles, include/alona_b200.h) are complete
    a kernel's length after the last byte lands.  Returns (DeviceCSR, libsize, gsum, gsumsq)."""
    from .engine import DeviceCSR
    dev = eng.device
    n = int(h_rowptr.numel() - 1)
    nnz = int(h_col.numel())
    rowptr = h_rowptr.to(dev, non_blocking=True)
    col = torch.empty(nnz, dtype=torch.int32, device=dev)
    val = torch.empty(nnz, dtype=torch.int32, device=dev)
    libsize = torch.empty(n, dtype=torch.int64, device=dev)
    acc = torch.zeros((n_genes, 4), dtype=torch.int64, device=dev)
    compute = torch.cuda.current_stream(dev)
    copy = torch.cuda.Stream(device=dev)
    copy.wait_stream(compute)
    rp = h_rowptr                                              # host view: tile boundaries in nonzeros
    tiles = max(1, min(int(tiles), n))
    bounds = [(n * t) // tiles for t in range(tiles + 1)]
    for t in range(tiles):
        r0, r1 = bounds[t], bounds[t + 1]
        a, b = int(rp[r0]), int(rp[r1])
        with torch.cuda.stream(copy):
            col[a:b].copy_(h_col[a:b], non_blocking=True)
            val[a:b].copy_(h_val[a:b], non_blocking=True)
            landed = torch.cuda.Event()
            landed.record(copy)
        compute.wait_event(landed)
        tile = DeviceCSR((rowptr[r0:r1 + 1] - rowptr[r0]).contiguous(), col[a:b], val[a:b], n_genes)
        lib_t, _ = eng.norm_moments(tile, acc=acc, finalize=False, libsize_out=libsize[r0:r1])
    if eng.world > 1:
        eng.allreduce_i64(acc)
    gsum, gsumsq = eng.moments_finalize(acc)
    return DeviceCSR(rowptr, col, val, n_genes, row_begin, n_total), libsize, gsum, gsumsq


def frontend(eng, counts, hvg_n, pca_n, nn_k, prune_snn=0.067, seed=42, knn_mode=0, v0=None,
             timer=None, want_overlap=False, tol=1e-4, maxit=1000, normalized=None):
    """Runs the whole path on `counts` (DeviceCSR shard).  Everything returned is on the device.
    Returns a dict with libsize, gsum, gsumsq, hvg_ranked, z, gene2hvg, AH, irlb (dict),
    emb_all [N x d], knn_idx (this shard, 0-based), knn_all, snn (dict) and statistics."""
    t = timer or StageTimer(False)
    bounds = shard_bounds(counts.n_total, eng.world)
    rows = [bounds[r + 1] - bounds[r] for r in range(eng.world)]
    t.mark('start')
    if normalized is not None:                     # (libsize, gsum, gsumsq) from upload_and_normalize
        libsize, gsum, gsumsq = normalized
    else:
        libsize, gsum, gsumsq = eng.norm_moments(counts)
    t.mark('normalize')
    ranked, z, gene2hvg, hstats = eng.hvg_seurat(gsum, gsumsq, counts.n_total, hvg_n)
    n_hvg = min(int(hvg_n), counts.n_genes)
    t.mark('hvg')
    AH = eng.hvg_extract(counts, libsize, gene2hvg, n_hvg)
    t.mark('hvg_extract')
    if v0 is None:
        v0 = start_vector(counts.n_total, seed)
    lo, hi = counts.row_begin, counts.row_begin + counts.n_local
    v0_local = torch.from_numpy(np.ascontiguousarray(v0[lo:hi])).to(eng.device)
    ir = eng.irlb(AH, pca_n, tol=tol, maxit=maxit, v0_local=v0_local, want_v=False, want_u=False)
    t.mark('pca')
    emb_all = eng.allgather_rows(ir['scores'], rows)
    idx, rdist, kstats = eng.knn(emb_all, nn_k, q_begin=lo, q_end=hi, mode=knn_mode)
    t.mark('knn')
    knn_all = eng.allgather_rows(idx, rows)
    vmin = prune_threshold(nn_k, prune_snn)
    sn = eng.snn(knn_all, vmin, row_begin=lo, row_end=hi, want_overlap=want_overlap)
    t.mark('snn')
    return {'libsize': libsize, 'gsum': gsum, 'gsumsq': gsumsq, 'hvg_ranked': ranked, 'z': z,
            'gene2hvg': gene2hvg, 'hvg_stats': hstats, 'AH': AH, 'irlb': ir, 'emb_all': emb_all,
            'knn_idx': idx, 'knn_rdist': rdist, 'knn_stats': kstats, 'knn_all': knn_all, 'snn': sn,
            'v_min': vmin}
