"""CPU baseline study of BASELINE.md §4 (a reported baseline, not the optimisation target).

Times the reference-style CPU path (oracle/alona_port.py: the same pandas / numpy+OpenBLAS / scikit-learn / scipy calls
as alona, line by line) on the host cores of the box it runs on:
  * C1 in full (2,700 cells x 32,738 genes, hvg 1000, 40 PCs, k 10), stage by stage;
  * knn (BallTree, single thread) and snn (scipy csr_matmat + the reference's Python loop) on clustered 50-d
    embeddings of N in --knn-sizes cells, with the fitted scaling exponents and the extrapolation to 1.3M cells.
No CUDA library of this repository is loaded.  One JSON document on stdout.

    python benchmarks/cpu_baseline_study.py [--knn-sizes 5000,10000,20000] [--skip-c1]
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--knn-sizes', default='5000,10000,20000')
    ap.add_argument('--skip-c1', action='store_true')
    a = ap.parse_args()
    from threadpoolctl import threadpool_limits, threadpool_info
    ncores = os.cpu_count() or 1
    limiter = threadpool_limits(limits=ncores, user_api='blas')     # noqa: F841
    blas = max([p.get('num_threads', 1) for p in threadpool_info() if p.get('user_api') == 'blas'] or [1])
    from oracle import alona_port, synth_np
    out = {'host_cores': ncores, 'blas_threads': blas,
           'single_threaded_stages': 'normalize, hvg (pandas), knn (BallTree, n_jobs default), snn (Python loop)'}
    if not a.skip_c1:
        import bench
        from alona_b200.synth import WORKLOADS
        wl = WORKLOADS['c1']
        df = bench.cpu_sample_frame(bench.sample_counts_host(wl, wl['cells']))
        tm = {}
        t0 = time.perf_counter()
        alona_port.run_pipeline(df, wl['hvg_n'], wl['pca_n'], wl['k'], 0.067, 42, timings=tm)
        dt = time.perf_counter() - t0
        out['c1'] = {'cells': wl['cells'], 'expressed_genes': int(df.shape[0]), 'seconds': dt,
                     'cells_per_s': wl['cells'] / dt, 'stage_s': tm}
    pts = []
    for n in [int(x) for x in a.knn_sizes.split(',')]:
        emb = synth_np.gaussian_mixture_embedding(n, d=50, seed=2)
        t0 = time.perf_counter()
        nn = alona_port.knn(emb, 20)
        t1 = time.perf_counter()
        edges = alona_port.snn(nn, 20, 0.067)
        t2 = time.perf_counter()
        pts.append({'cells': n, 'knn_s': t1 - t0, 'snn_s': t2 - t1, 'edges': int(len(edges))})
    out['knn_snn_k20_d50'] = pts
    if len(pts) >= 2:
        ln = np.log([p['cells'] for p in pts])
        for key in ('knn_s', 'snn_s'):
            slope, icpt = np.polyfit(ln, np.log([p[key] for p in pts]), 1)
            out[key.replace('_s', '_exponent')] = float(slope)
            out[key.replace('_s', '_extrapolated_s_at_1306127')] = float(np.exp(icpt + slope * np.log(1306127.0)))
    print(json.dumps(out))


if __name__ == '__main__':
    main()
