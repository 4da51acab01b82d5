"""N3 microbenchmark: QC metrics pass and CSR compaction at a workload's full size, device-resident input.

Prints one JSON line: time per call (CUDA events, after a warm-up call), algorithmic bytes and the fraction of the
measured HBM peak (MEASURED_PEAKS.json if present, else the profiling guide's fallback)."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from alona_b200.engine import get_engine
from alona_b200.synth import WORKLOADS, make_profiles


def timed(fn, reps):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--workload', default='c3')
    ap.add_argument('--cells', type=int, default=0)
    ap.add_argument('--reps', type=int, default=3)
    a = ap.parse_args()
    wl = dict(WORKLOADS[a.workload])
    if a.cells:
        wl['cells'] = a.cells
    eng = get_engine()
    prof = torch.from_numpy(make_profiles(wl['genes'])).to(eng.device)
    A = eng.synth_counts(prof, 0, wl['cells'], wl['lbar'], n_total=wl['cells'])
    n, G, nnz = A.n_local, A.n_genes, A.nnz
    peak = 6538.9
    pk = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'MEASURED_PEAKS.json')
    src = 'fallback'
    if os.path.exists(pk):
        try:
            peak = float(json.load(open(pk))['hbm_gbs'])
            src = 'MEASURED_PEAKS.json'
        except Exception:
            pass
    flags = (np.arange(G) % 50 == 0).astype(np.uint8)
    ms_q, q = timed(lambda: eng.qc_metrics(A, flags), a.reps)
    reads = q['reads_per_cell']
    thr = int(torch.quantile(reads[:: max(1, n // 100000)].double(), 0.1).item())
    keep_row = (reads > thr).cpu().numpy()
    gc = q['gene_cells'].cpu().numpy()
    keep_gene = gc / n > 0.01                                         # the reference's default --minexpgenes
    gene_map = np.where(keep_gene, np.cumsum(keep_gene) - 1, -1).astype(np.int32)
    ms_f, (out, _) = timed(lambda: eng.filter_csr(A, keep_row, gene_map, n_genes_out=int(keep_gene.sum())), a.reps)
    by_q = nnz * 8 + n * 8 + n * (8 + 4 + 16) + G * 8
    by_f = nnz * 8 * 2 + out.nnz * 8 + n * 8 * 2 + out.n_local * 16
    print(json.dumps({'cells': n, 'genes': G, 'nnz': nnz, 'kept_cells': out.n_local, 'kept_genes': int(keep_gene.sum()),
                      'kept_nnz': out.nnz,
                      'qc_metrics': {'ms': round(ms_q, 3), 'bytes': by_q, 'gbps': round(by_q / ms_q / 1e6, 1),
                                     'frac_of_hbm_peak': round(by_q / ms_q / 1e6 / peak, 3)},
                      'filter_count_fill': {'ms': round(ms_f, 3), 'bytes': by_f, 'gbps': round(by_f / ms_f / 1e6, 1),
                                            'frac_of_hbm_peak': round(by_f / ms_f / 1e6 / peak, 3),
                                            'note': 'includes host->device copy of the masks and output allocation'},
                      'peak_gbps': peak, 'peak_source': src}))


if __name__ == '__main__':
    main()
