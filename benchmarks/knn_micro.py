"""C4 microbench: kNN + SNN on an N x d FP64 mixture embedding, k sweep."""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from alona_b200.engine import get_engine
from alona_b200.pipeline import prune_threshold


def mixture(n, d, dev, seed=2, n_clusters=30):
    g = torch.Generator(device=dev); g.manual_seed(seed)
    scale_c = torch.linspace(3.0, 0.3, d, device=dev, dtype=torch.float64)
    centres = torch.randn(n_clusters, d, generator=g, device=dev, dtype=torch.float64) * 6.0 * scale_c
    scale = torch.linspace(2.0, 0.5, d, device=dev, dtype=torch.float64)
    cl = torch.randint(0, n_clusters, (n,), generator=g, device=dev)
    x = centres[cl] + torch.randn(n, d, generator=g, device=dev, dtype=torch.float64) * scale
    return x.contiguous()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--n', type=int, default=1306127)
    ap.add_argument('--d', type=int, default=50)
    ap.add_argument('--ks', default='20')
    ap.add_argument('--check', type=int, default=2000, help='queries re-checked with the exact FP64 scan')
    ap.add_argument('--reps', type=int, default=2)
    ap.add_argument('--snn', action='store_true')
    a = ap.parse_args()
    eng = get_engine()
    eng.prof_set_level(1)
    emb = mixture(a.n, a.d, eng.device)
    for k in [int(x) for x in a.ks.split(',')]:
        for rep in range(a.reps):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            idx, rd, st = eng.knn(emb, k)
            torch.cuda.synchronize(); dt = time.perf_counter() - t0
        s = st.cpu().numpy()
        prof = eng.prof_read()
        relerr = float(np.array([s[4]], dtype=np.uint32).view(np.float32)[0])
        rec = {'n': a.n, 'd': a.d, 'k': k, 'knn_s': dt, 'tflops_alg': 2.0 * a.n * a.n * a.d / dt / 1e12,
               'tie_rows': int(s[0]), 'fallback_rows': int(s[1]), 'dup_rows': int(s[2]), 'certified': int(s[3]),
               'max_rel_err': relerr, 'kappa': 2.0 ** -16, 'm': int(s[5]), 'mt': int(s[6]),
               'kernel_ms': {k2: round(v[0] / v[1], 3) for k2, v in prof.items()},
               'tflops_alg_main': 2.0 * a.n * a.n * a.d / (prof['knn_main'][0] / prof['knn_main'][1] / 1e3) / 1e12
               if 'knn_main' in prof else None}
        if a.check:
            lo = (a.n // 2 // 128) * 128
            ex, exd, _ = eng.knn(emb, k, q_begin=lo, q_end=min(a.n, lo + a.check), mode=1)
            rec['check_rows'] = int(ex.shape[0])
            rec['check_equal'] = bool(torch.equal(ex, idx[lo:lo + ex.shape[0]]) and torch.equal(exd, rd[lo:lo + ex.shape[0]]))
        if a.snn:
            torch.cuda.synchronize(); t0 = time.perf_counter()
            res = eng.snn(idx, prune_threshold(k, 0.067), want_overlap=False)
            torch.cuda.synchronize(); rec['snn_s'] = time.perf_counter() - t0
            rec['edges'] = int(res['src'].numel()); rec['snn_max_walk'] = res['max_walk']
        print(json.dumps(rec), flush=True)


if __name__ == '__main__':
    main()
