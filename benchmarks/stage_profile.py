"""Per-kernel device times (CUDA-event slots of the C-ABI, level 2) of one front-end pass."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from alona_b200.engine import get_engine
from alona_b200.pipeline import frontend, StageTimer, start_vector
from alona_b200.synth import WORKLOADS, make_profiles


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--workload', default='c3')
    ap.add_argument('--cells', type=int, default=0)
    ap.add_argument('--level', type=int, default=2)
    a = ap.parse_args()
    wl = dict(WORKLOADS[a.workload])
    if a.cells:
        wl['cells'] = a.cells
    eng = get_engine()
    prof = torch.from_numpy(make_profiles(wl['genes'])).to(eng.device)
    A = eng.synth_counts(prof, 0, wl['cells'], wl['lbar'], n_total=wl['cells'])
    v0 = start_vector(wl['cells'], 42)
    frontend(eng, A, wl['hvg_n'], wl['pca_n'], wl['k'], 0.067, 42, v0=v0)          # warm-up
    eng.prof_set_level(a.level)
    eng.prof_read()
    t = StageTimer(True)
    out = frontend(eng, A, wl['hvg_n'], wl['pca_n'], wl['k'], 0.067, 42, v0=v0, timer=t)
    st = t.result()
    pr = eng.prof_read()
    print(json.dumps({'cells': wl['cells'], 'stage_ms': st,
                      'kernel_ms': {k: {'total_ms': round(v[0], 3), 'launches': v[1]} for k, v in pr.items()},
                      'irlb': {k: out['irlb'][k] for k in ('steps', 'nmult', 'sum_panel')},
                      'nnz': A.nnz, 'nnz_hvg': out['AH'].nnz, 'edges': int(out['snn']['src'].numel()),
                      'snn': {k: out['snn'][k] for k in ('max_walk', 'sum_walk', 'nonzeros_unpruned')}}))


if __name__ == '__main__':
    main()
