"""N>1 host-side logic on CPU: two processes over torch.distributed (gloo, 127.0.0.1).

What shards and how: cells are split into contiguous row ranges; the exchanges are
an all-reduce of the per-gene moments, an all-gather of the (ragged) embedding / kNN row blocks and,
inside IRLBA, all-reduces of H- and (j+1)-vectors.  The GPU path does these through the C-ABI
(ab_allreduce_f64 / ab_allgather_bytes over NCCL); here the same host helpers (shard_bounds,
pack_rows / unpack_rows, start-vector slicing, prune threshold) run over gloo with the oracle's
numpy arithmetic standing in for the kernels, and the result must equal the single-process one.
"""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from alona_b200.pipeline import shard_bounds, pack_rows, unpack_rows, start_vector, prune_threshold
        from oracle import restate as R
        from oracle import synth_np
        m, _ = synth_np.nb_counts(1001, 3000, lbar=400.0, seed=5, drop_empty=False)      # ragged: 1001 = 500 + 501
        n = m.shape[0]
        b = shard_bounds(n, world)
        lo, hi = b[rank], b[rank + 1]
        rows = [b[r + 1] - b[r] for r in range(world)]
        shard = m[lo:hi]

        # K1: per-gene moments of the shard, summed across ranks (= ab_allreduce_f64 of gsum|gsumsq)
        sx, sx2, _ = R.gene_moments(shard)
        both = torch.from_numpy(np.concatenate([sx, sx2]))
        dist.all_reduce(both)
        fx, fx2, _ = R.gene_moments(m)
        assert np.allclose(both.numpy()[:m.shape[1]], fx, rtol=1e-13, atol=0)
        assert np.allclose(both.numpy()[m.shape[1]:], fx2, rtol=1e-13, atol=0)
        assert np.array_equal(R.libsize(shard), R.libsize(m)[lo:hi])                     # library sizes are local
        # the product all-reduces the fixed-point accumulators as int64 (ab_allreduce_i64): exactly additive, so the
        # sharded sum has the SAME BITS as the single-shard one, for any world size
        acc = torch.from_numpy(R.gene_moments_fixed(shard))
        dist.all_reduce(acc)
        assert np.array_equal(acc.numpy(), R.gene_moments_fixed(m))
        gx, gx2 = R.moments_from_fixed(acc.numpy())
        ok = fx > 0
        assert np.max(np.abs(gx[ok] - fx[ok]) / fx[ok]) < 1e-13 and np.max(np.abs(gx2[ok] - fx2[ok]) / fx2[ok]) < 1e-12

        # K2 is replicated: every rank must select the same genes from the reduced moments
        mean = both.numpy()[:m.shape[1]] / n
        with np.errstate(invalid='ignore', divide='ignore'):
            var = (both.numpy()[m.shape[1]:] - n * mean * mean) / (n - 1)
            ranked = R.hvg_seurat_from_moments(mean, var, 200)[0]
        mine = torch.from_numpy(np.asarray(ranked, dtype=np.int64))
        other = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(other, mine)
        assert all(torch.equal(o, mine) for o in other)

        # start vector: every rank draws the reference's vector and keeps its slice (irlb.py:151-152)
        v0 = start_vector(n, 42)
        parts = [torch.empty(max(rows), dtype=torch.float64) for _ in range(world)]
        dist.all_gather(parts, pack_rows(torch.from_numpy(v0[lo:hi].copy()), max(rows)))
        assert np.array_equal(unpack_rows(torch.cat(parts), rows).numpy(), v0)

        # K9 exchange: ragged all-gather of the embedding rows (pack -> equal-sized all-gather -> unpack)
        rng = np.random.RandomState(3)
        emb = rng.randn(n, 7)
        g = [torch.empty((max(rows), 7), dtype=torch.float64) for _ in range(world)]
        dist.all_gather(g, pack_rows(torch.from_numpy(emb[lo:hi].copy()), max(rows)))
        emb_all = unpack_rows(torch.cat(g, dim=0), rows).numpy()
        assert np.array_equal(emb_all, emb)

        # K9/K10: each rank answers its own query rows against the gathered database, the kNN table is
        # gathered the same way and each rank emits its own SNN rows; rank order = row order
        k = 8
        idx_local, _ = R.knn_bruteforce(emb_all, k, queries=emb_all[lo:hi])
        gi = [torch.empty((max(rows), k), dtype=torch.int64) for _ in range(world)]
        dist.all_gather(gi, pack_rows(torch.from_numpy(idx_local.astype(np.int64)), max(rows)))
        knn_all = unpack_rows(torch.cat(gi, dim=0), rows).numpy()
        full_idx, _ = R.knn_bruteforce(emb, k)
        assert np.array_equal(knn_all, full_idx)
        e_full, c_full = R.snn_restated(knn_all, k, 0.067)
        sel = (e_full[:, 0] >= lo) & (e_full[:, 0] < hi)
        np.save(os.path.join(out_dir, 'edges_%d.npy' % rank), e_full[sel])
        dist.barrier()
        if rank == 0:
            cat = np.concatenate([np.load(os.path.join(out_dir, 'edges_%d.npy' % r)) for r in range(world)])
            assert np.array_equal(cat, e_full)                                           # concatenation in rank order
            assert prune_threshold(k, 0.067) == R.prune_threshold(k, 0.067)

        # IRLBA: the restated sharded Lanczos (partial A.v summed over shards) follows the unsharded trajectory
        hv = np.sort(np.asarray(ranked[:120]))
        At, _ = R.hvg_matrix(m, hv)
        r1 = R.lanczos_csr(At, 10, seed=42, nshards=1)
        r2 = R.lanczos_csr(At, 10, seed=42, nshards=world)
        assert (r1['steps'], r1['nmult']) == (r2['steps'], r2['nmult'])
        assert R.column_rel_err(r1['V'] * r1['s'], r2['V'] * r2['s']).max() < 1e-9
    finally:
        dist.destroy_process_group()


def test_two_rank_sharding_over_gloo(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)


def test_shard_bounds_cover_rows_exactly():
    from alona_b200.pipeline import shard_bounds
    for n in (1, 7, 1000, 1306127):
        for w in (1, 2, 4, 8):
            b = shard_bounds(n, w)
            assert b[0] == 0 and b[-1] == n and all(b[i] <= b[i + 1] for i in range(w))
            assert max(b[i + 1] - b[i] for i in range(w)) - min(b[i + 1] - b[i] for i in range(w)) <= 1


def test_pack_unpack_roundtrip():
    from alona_b200.pipeline import pack_rows, unpack_rows
    parts = [torch.arange(6.).reshape(3, 2), torch.arange(4.).reshape(2, 2) + 10]
    g = torch.cat([pack_rows(p, 3) for p in parts])
    assert torch.equal(unpack_rows(g, [3, 2]), torch.cat(parts))
    assert pack_rows(parts[0], 3).data_ptr() == parts[0].data_ptr()          # no copy when already full


def _reshard_worker(rank, world, port):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        import scipy.sparse as sp
        from alona_b200.pipeline import reshard_csr, shard_bounds
        rng = np.random.RandomState(0)
        full = sp.random(97, 40, density=0.2, random_state=rng, format='csr')
        full.data = rng.randint(1, 9, size=full.nnz).astype(np.int32)
        full.sort_indices()
        # ragged shards, as a filter leaves them: rank 0 holds 11 rows, rank 1 holds 80, rank 2 holds 6
        cuts = [0, 11, 91, 97]
        mine = full[cuts[rank]:cuts[rank + 1]]
        rowptr, col, val, row_begin, n_total = reshard_csr(torch.from_numpy(mine.indptr.astype(np.int64)),
                                                           torch.from_numpy(mine.indices.astype(np.int32)),
                                                           torch.from_numpy(mine.data.astype(np.int32)), rank, world)
        b = shard_bounds(97, world)
        want = full[b[rank]:b[rank + 1]]
        assert (row_begin, n_total) == (b[rank], 97)
        assert np.array_equal(rowptr.numpy(), want.indptr)
        assert np.array_equal(col.numpy(), want.indices)
        assert np.array_equal(val.numpy(), want.data)
    finally:
        dist.destroy_process_group()


def test_reshard_after_filter_over_gloo():
    """N3 on N > 1: ragged shards left by the filters are re-balanced to shard_bounds() by whole-row exchange."""
    mp.spawn(_reshard_worker, args=(3, _free_port()), nprocs=3, join=True)


def test_reshard_plan_covers_every_row_once():
    from alona_b200.pipeline import reshard_plan
    for counts in ([5, 0, 7], [0, 0, 9], [1, 1, 1, 1], [100, 3], [0, 0]):
        world = len(counts)
        plan, bounds = reshard_plan(counts, world)
        for r in range(world):
            sent = sorted((s, c) for (s, c) in plan[r] if c)
            assert sum(c for _, c in sent) == counts[r]
            pos = sent[0][0] if sent else 0
            for s, c in sent:
                assert s == pos
                pos += c
        for t in range(world):
            assert sum(plan[r][t][1] for r in range(world)) == bounds[t + 1] - bounds[t]
