"""N3 on N > 1 GPUs (run under torchrun, one rank per GPU; not collected by pytest):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tests/multi_gpu_qc_check.py

Every rank drives the AlonaCell mirror (remove_empty -> remove_mito -> simple_filters -> normalize) on its shard of
the reference-minted fixtures; the kept genes / cells must equal the reference's, the re-balanced shards must be
the rows shard_bounds() assigns, and the library sizes after normalize() must be those of the surviving frame."""
import os
import sys

import numpy as np
import pandas as pd
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))


def main():
    local_rank = int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local_rank)
    dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    from conftest import load_golden, golden_csr, QC_CASES
    from alona_b200.cell import AlonaCell
    from alona_b200.engine import get_engine
    from alona_b200.pipeline import shard_bounds
    eng = get_engine()
    checked = 0
    for case in QC_CASES:
        g = load_golden(case)
        m = golden_csr(g)
        genes = g['genes']
        cells = np.array(['c%d' % i for i in range(m.shape[0])])
        df = pd.DataFrame(m.T.toarray().astype(np.int64), index=genes, columns=cells)
        for remove_mito in ('no', 'yes'):
            for ti, thr in enumerate(g['thresholds']):
                o = AlonaCell()
                o.params = {'dataformat': 'raw', 'minreads': int(g['minreads']), 'minexpgenes': float(thr),
                            'remove_mito': remove_mito, 'qc_auto': False}
                o.data = df
                for step, fn in (('after_remove_empty', o.remove_empty), ('after_remove_mito', o.remove_mito),
                                 ('after_simple_filters', o.simple_filters)):
                    fn()
                    key = '%s_mito%s_t%d' % (step, remove_mito, ti)
                    assert list(o.data.index) == list(genes[g[key + '_genes']]), (key, eng.rank)
                    assert list(o.data.columns) == list(cells[g[key + '_cells']]), (key, eng.rank)
                    b = shard_bounds(len(g[key + '_cells']), eng.world)
                    assert (o.data.csr.row_begin, o.data.csr.n_local) == (b[eng.rank], b[eng.rank + 1] - b[eng.rank]), key
                ref = df.loc[genes[g[key + '_genes']], cells[g[key + '_cells']]]
                lo = o.data.csr.row_begin
                c = o.data.csr
                import scipy.sparse as sp
                mine = sp.csr_matrix((c.val.cpu().numpy(), c.col.cpu().numpy(), c.rowptr.cpu().numpy()),
                                     shape=(c.n_local, c.n_genes)).toarray()
                assert np.array_equal(mine, ref.values.T[lo:lo + c.n_local]), key
                dn = o.normalize(o.data, '', 'raw', False)
                assert np.array_equal(dn.libsize.cpu().numpy(), ref.sum(axis=0).values[lo:lo + c.n_local]), key
                checked += 1
    dist.barrier()
    if eng.rank == 0:
        print('multi_gpu_qc_check: %d filter chains identical to the reference on %d ranks' % (checked, eng.world))
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
