"""N2 microbenchmark: per-cluster moment tables and medians at a workload's full size (device-resident counts,
30 pseudo-clusters assigned round-robin over blocks of cells)."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from alona_b200.engine import get_engine
from alona_b200.synth import WORKLOADS, make_profiles


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--workload', default='c3')
    ap.add_argument('--cells', type=int, default=0)
    ap.add_argument('--clusters', type=int, default=30)
    a = ap.parse_args()
    wl = dict(WORKLOADS[a.workload])
    if a.cells:
        wl['cells'] = a.cells
    eng = get_engine()
    prof = torch.from_numpy(make_profiles(wl['genes'])).to(eng.device)
    A = eng.synth_counts(prof, 0, wl['cells'], wl['lbar'], n_total=wl['cells'])
    lib, _, _ = eng.norm_moments(A)
    n, G, C = A.n_local, A.n_genes, a.clusters
    labels = ((np.arange(n) // 16384) % C).astype(np.int32)        # the generator's tiles share a cluster profile
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    eng.cluster_moments(A, lib, labels, C)                          # warm-up
    torch.cuda.synchronize()
    ev[0].record()
    mom = eng.cluster_moments(A, lib, labels, C)
    ev[1].record()
    med, info = eng.cluster_median(A, lib, labels, C, nnz=mom['nnz'])
    ev[2].record()
    torch.cuda.synchronize()
    by_m = A.nnz * 8 + n * 20 + 3 * C * G * 8
    print(json.dumps({'cells': n, 'genes': G, 'clusters': C, 'nnz': A.nnz,
                      'moments_ms': round(ev[0].elapsed_time(ev[1]), 3), 'moments_bytes': by_m,
                      'moments_gbps': round(by_m / ev[0].elapsed_time(ev[1]) / 1e6, 1),
                      'median_ms': round(ev[1].elapsed_time(ev[2]), 3), 'median_gathered_values': info['gathered'],
                      'median_batch': info['batch'], 'nonzero_medians': int((med > 0).sum().item())}))


if __name__ == '__main__':
    main()
