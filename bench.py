#!/usr/bin/env python
"""bench.py — cells/s of normalise -> HVG -> PCA -> kNN -> SNN on synthetic NB counts.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|c3|...] [--impl ours|reference]

One "step" = one pass of the whole front end over the workload (device-resident cell-major CSR
counts in, device-resident SNN edge list out).  N>1 is launched by torchrun, one rank per GPU;
cells are sharded by rows (strong scaling: total work fixed).  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = 'cells/sec normalise->PCA->kNN->SNN'
UNIT = 'cells/s'
CPU_SAMPLE_CELLS = 6000          # ~12 s of the reference-style CPU path per pass on the box's host cores


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=3)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--workload', default=os.environ.get('AB_WORKLOAD', 'c3'))
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--seed', type=int, default=42)
    ap.add_argument('--knn-mode', type=int, default=0)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--cells', type=int, default=0, help='override the workload cell count (debug)')
    ap.add_argument('--verbose', action='store_true', help='progress markers on stderr (debug)')
    return ap.parse_args()


# ------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index=0):
        self.index = index
        self.samples = []
        self._stop = threading.Event()
        self._t = None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                      '--format=csv,noheader,nounits'], capture_output=True, text=True, timeout=5).stdout
                f = [x.strip() for x in out.strip().split(',')]
                if len(f) >= 7:
                    self.samples.append(f)
            except Exception:
                pass
            self._stop.wait(0.2)

    def start(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()

    def stop(self):
        self._stop.set()
        if self._t:
            self._t.join(timeout=6)
        sm = [float(s[0]) for s in self.samples if s[0].replace('.', '').isdigit()]
        mx = [float(s[1]) for s in self.samples if s[1].replace('.', '').isdigit()]
        reasons = set()
        for s in self.samples:
            for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), s[3:7]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': sorted(reasons), 'samples': len(self.samples)}


# ------------------------------------------------------------------------------------------
# CPU baseline: the oracle port (same third-party calls as the reference) on a bounded sample
# ------------------------------------------------------------------------------------------
def cpu_sample_frame(host_csr_sample):
    """genes x cells int64 DataFrame of the sample, all-zero genes dropped (alona/cell.py:64-81)."""
    import pandas as pd
    keep = np.asarray((host_csr_sample != 0).sum(axis=0)).ravel() > 0
    sub = host_csr_sample[:, keep]
    return pd.DataFrame(sub.T.toarray().astype(np.int64), index=['g%d' % i for i in np.where(keep)[0]],
                        columns=['c%d' % i for i in range(sub.shape[0])])


def run_cpu_port(df, wl, seed):
    from oracle import alona_port
    tm = {}
    t0 = time.perf_counter()
    alona_port.run_pipeline(df, min(wl['hvg_n'], df.shape[0]), wl['pca_n'], wl['k'], 0.067, seed, timings=tm)
    return time.perf_counter() - t0, tm


def cpu_threads():
    try:
        from threadpoolctl import threadpool_info
        n = [p.get('num_threads', 1) for p in threadpool_info() if p.get('user_api') == 'blas']
        return max(n) if n else 1
    except Exception:
        return os.cpu_count() or 1


def sample_counts_host(wl, n_cells, seed_data=20251019):
    """First n_cells cells of the workload's synthetic matrix, built on the HOST by the numpy twin of the device
    generator (oracle/synth_np.device_twin_counts): the CPU arm loads no CUDA library of this repo."""
    from oracle import synth_np
    from alona_b200.synth import make_profiles
    return synth_np.device_twin_counts(make_profiles(wl['genes']), 0, n_cells, wl['lbar'], seed=seed_data)


def blas_threads(n):
    """Pins the BLAS pool to n threads (torchrun exports OMP_NUM_THREADS=1) and returns the limiter to keep alive."""
    try:
        from threadpoolctl import threadpool_limits
        return threadpool_limits(limits=int(n), user_api='blas')
    except Exception:
        return None


def table_hashes(eng, out, lo, world):
    """Order-sensitive 64-bit checksums of the kNN table and of the edge list, combinable across ranks (a wrapping
    sum of per-entry mixes keyed on the GLOBAL position), plus the global edge count: identical lines at
    N = 1/2/4/8 prove that the sharded runs produced the same graph.  Measurement only (torch integer ops)."""
    import torch
    import torch.distributed as dist
    dev = eng.device
    c1, c2, c3 = -7046029254386353131, -4417276706812531889, 1609587929392839161      # odd 64-bit constants

    def mix(pos, val):
        h = ((pos + 1) * c1) ^ ((val.to(torch.int64) + 1) * c2)
        h = (h ^ (h >> 29)) * c3
        return (h ^ (h >> 32)).sum()
    knn = out['knn_idx']
    k = knn.shape[1]
    pos = (torch.arange(knn.shape[0], device=dev, dtype=torch.int64) + lo).unsqueeze(1) * k + \
        torch.arange(k, device=dev, dtype=torch.int64).unsqueeze(0)
    hk = mix(pos, knn)
    ne = torch.tensor([out['snn']['src'].numel()], dtype=torch.int64, device=dev)
    off = 0
    if world > 1:
        allc = [torch.zeros_like(ne) for _ in range(world)]
        dist.all_gather(allc, ne)
        off = int(sum(int(c.item()) for c in allc[:dist.get_rank()]))
    epos = torch.arange(int(ne.item()), device=dev, dtype=torch.int64) + off
    he = mix(2 * epos, out['snn']['src']) + mix(2 * epos + 1, out['snn']['dst'])
    t = torch.stack([hk, he, ne[0]])
    if world > 1:
        dist.all_reduce(t)
    v = t.cpu().numpy().astype(np.uint64)
    return {'knn_hash': '%016x' % int(v[0]), 'edges_hash': '%016x' % int(v[1]), 'edges_total': int(t[2].item())}


# ------------------------------------------------------------------------------------------
def main():
    args = parse()
    from alona_b200.synth import WORKLOADS
    wl = dict(WORKLOADS[args.workload])
    if args.cells:
        wl['cells'] = args.cells
    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local_rank = int(os.environ.get('LOCAL_RANK', 0))
    config = {'workload': '%s: synthetic NB %d cells x %d genes, hvg_n %d, pca_n %d, k %d, prune 0.067'
              % (args.workload, wl['cells'], wl['genes'], wl['hvg_n'], wl['pca_n'], wl['k']),
              'parallelism': 'cells row-sharded over %d GPU(s)' % world,
              'cache': 'inputs (CSR counts) larger than L2; no flush needed'}

    if args.impl == 'reference':
        if rank != 0:
            return
        n_s = min(CPU_SAMPLE_CELLS, wl['cells'])
        ncores = os.cpu_count() or 1
        limiter = blas_threads(ncores)                      # noqa: F841  (kept alive for the whole run)
        df = cpu_sample_frame(sample_counts_host(wl, n_s))
        times = []
        for i in range(args.warmup + args.steps):
            dt, tm = run_cpu_port(df, wl, args.seed)
            if i >= args.warmup:
                times.append(dt)
        ms = 1e3 * float(np.mean(times))
        val = n_s / (ms / 1e3)
        if n_s == wl['cells']:
            sample = 'the whole %d-cell matrix (%d expressed genes), dense pandas path of the reference' % (n_s, df.shape[0])
        else:
            sample = ('first %d cells of the workload matrix (%d expressed genes), dense pandas path of the reference; '
                      'the full %d-cell matrix is infeasible on a CPU (dense float64 frames of %.0f GB, ~quadratic kNN)'
                      % (n_s, df.shape[0], wl['cells'], wl['cells'] * wl['genes'] * 8 / 1e9))
            config['workload'] += ' -- CPU ARM TIMED ON A SAMPLE: ' + sample
        print(json.dumps({
            'impl': 'reference', 'metric': METRIC, 'value': val, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'strong',
            'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic', 'config': config,
            'cpu_baseline': {'value': val, 'unit': UNIT, 'cores': ncores, 'blas_threads': cpu_threads(), 'kind': 'port',
                             'sample': sample, 'sample_cells': n_s, 'stage_s': tm,
                             'single_threaded_stages': 'normalize, hvg (pandas), knn (BallTree, n_jobs default), snn'},
            'e2e': {'value': val, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}))
        return

    def note(msg):
        if args.verbose:
            sys.stderr.write('[bench rank %d %.1fs] %s\n' % (rank, time.perf_counter() - t_start, msg))
            sys.stderr.flush()
    t_start = time.perf_counter()
    if args.verbose:
        import faulthandler
        import signal
        faulthandler.register(signal.SIGUSR1, all_threads=True)
        faulthandler.dump_traceback_later(int(os.environ.get('AB_BENCH_DUMP_AFTER', '90')), exit=True)
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a GPU (no CPU fallback); use --impl reference for the CPU arm')
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    from alona_b200.engine import get_engine
    from alona_b200.pipeline import frontend, StageTimer, shard_bounds, start_vector, upload_and_normalize
    from alona_b200.synth import make_profiles
    note('process group up')
    eng = get_engine()
    dev = eng.device
    note('engine + NCCL communicator up')

    # ---- synthetic input, generated on the device, shard by shard ---------------------------
    bounds = shard_bounds(wl['cells'], world)
    lo, hi = bounds[rank], bounds[rank + 1]
    prof = torch.from_numpy(make_profiles(wl['genes'])).to(dev)
    A = eng.synth_counts(prof, lo, hi - lo, wl['lbar'], n_total=wl['cells'])
    torch.cuda.synchronize()
    note('synthetic shard generated: %d nnz' % A.nnz)
    nnz_local = A.nnz
    v0 = start_vector(wl['cells'], args.seed)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def one_step(timer=None):
        return frontend(eng, A, wl['hvg_n'], wl['pca_n'], wl['k'], 0.067, args.seed, knn_mode=args.knn_mode, v0=v0,
                        timer=timer)

    for i in range(args.warmup):
        out = one_step(StageTimer(True, verbose=True) if args.verbose else None)
        note('warm-up step %d done' % i)
    barrier()
    note('barrier after warm-up passed')
    eng.prof_set_level(1)                       # one CUDA-event pair per stage-level kernel (no per-launch events)
    eng.prof_read()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = eng.launch_count()
    timer = StageTimer(True)
    ev0 = torch.cuda.Event(enable_timing=True)
    ev1 = torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        out = one_step(timer)
    ev1.record()
    note('timed steps enqueued')
    barrier()
    note('timed region closed')
    total_ms = ev0.elapsed_time(ev1)
    launches = eng.launch_count() - launches0
    kernel_ms = {k: v[0] / v[1] for k, v in eng.prof_read().items()}     # mean device ms per launch, timed region
    clocks = sampler.stop() if rank == 0 else None
    stage_ms = {k: v / args.steps for k, v in timer.result().items()}
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_per_step = float(t.item()) / args.steps
    value = wl['cells'] / (ms_per_step / 1e3)
    edges_local = int(out['snn']['src'].numel())
    nnz_h = out['AH'].nnz
    hashes = table_hashes(eng, out, lo, world)
    info = {'steps': out['irlb']['steps'], 'nmult': out['irlb']['nmult'], 'sum_panel': out['irlb']['sum_panel'],
            'fused_allreduce': out['irlb'].get('fused_allreduce', False), **hashes,
            'nnz': nnz_local, 'nnz_hvg': nnz_h, 'edges_local': edges_local,
            'knn_stats': out['knn_stats'].cpu().numpy().tolist(), 'snn_max_walk': out['snn']['max_walk'], 'snn_nonzeros_unpruned': out['snn']['nonzeros_unpruned'],
            'hvg_stats': out['hvg_stats'].cpu().numpy().tolist()}

    # ---- per-kernel table: one extra, untimed pass (ALL ranks: it contains the collectives) with an
    # event pair around every launch -----------------------------------------------------------
    eng.prof_set_level(2)
    eng.prof_read()
    o2 = one_step()
    torch.cuda.synchronize()
    fine = eng.prof_read()
    eng.prof_set_level(0)
    fine_snn = {'sum_walk': o2['snn'].get('sum_walk', 0), 'edges': int(o2['snn']['src'].numel())}
    del o2
    note('fine-grained profiling pass done')

    # ---- end to end through the public API with HOST buffers -------------------------------
    e2e = None
    if not args.no_e2e:
        h_rowptr = A.rowptr.cpu().pin_memory()
        h_col = A.col.cpu().pin_memory()
        h_val = A.val.cpu().pin_memory()
        h_v0 = torch.from_numpy(v0[lo:hi].copy()).pin_memory()
        from alona_b200.engine import DeviceCSR

        def e2e_step():
            # upload in row tiles, K1 on tile i while tile i+1 is still crossing PCIe
            d, lib, gs, gq = upload_and_normalize(eng, h_rowptr, h_col, h_val, wl['genes'], lo, wl['cells'])
            o = frontend(eng, d, wl['hvg_n'], wl['pca_n'], wl['k'], 0.067, args.seed, knn_mode=args.knn_mode, v0=v0,
                         normalized=(lib, gs, gq))
            src = o['snn']['src'].to('cpu', non_blocking=True)
            dst = o['snn']['dst'].to('cpu', non_blocking=True)
            torch.cuda.synchronize()
            return src, dst
        del out
        for _ in range(1):
            e2e_step()
        barrier()
        t0 = time.perf_counter()
        ne = max(1, min(args.steps, 3))
        for _ in range(ne):
            src, dst = e2e_step()
        barrier()
        dt = (time.perf_counter() - t0) / ne
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        h2d = (h_rowptr.numel() * 8 + h_col.numel() * 4 + h_val.numel() * 4 + h_v0.numel() * 8)
        d2h = src.numel() * 4 * 2
        tot = torch.tensor([h2d, d2h], dtype=torch.int64, device=dev)
        if world > 1:
            dist.all_reduce(tot)
        e2e = {'value': wl['cells'] / float(tt.item()), 'unit': UNIT, 'h2d_bytes_per_step': int(tot[0].item()),
               'd2h_bytes_per_step': int(tot[1].item()), 'ms_per_step': 1e3 * float(tt.item()),
               'h2d_gb_per_s': float(tot[0].item()) / world / 1e9 / max(1e-9, float(tt.item())),
               'path': 'pinned host CSR -> alona_b200.pipeline.upload_and_normalize (row tiles, K1 overlapped with the '
                       'upload) -> frontend (C-ABI stages) -> host edge list'}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel ------------------------------------------------------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except Exception:
        pass
    roof = roofline(wl, world, kernel_ms, peaks)
    stages_roof = stage_rooflines(wl, world, fine, info, fine_snn, peaks)

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        n_s = min(CPU_SAMPLE_CELLS, wl['cells'])
        rp = A.rowptr[:n_s + 1].cpu().numpy()
        import scipy.sparse as sp
        m = sp.csr_matrix((A.val[:rp[-1]].cpu().numpy(), A.col[:rp[-1]].cpu().numpy(), rp), shape=(n_s, wl['genes']))
        df = cpu_sample_frame(m)
        dt, tm = run_cpu_port(df, wl, args.seed)
        cpu = {'value': n_s / dt, 'unit': UNIT, 'cores': cpu_threads(), 'kind': 'port',
               'sample': 'first %d cells of the same matrix (%d expressed genes), dense pandas path, one run'
                         % (n_s, df.shape[0]), 'stage_s': tm}

    line = {'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': ms_per_step, 'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None,
            'dtype': 'f64 (normalise/HVG/IRLBA/re-rank), fp16 hi/lo operands with fp32 accumulate (distance GEMM candidates), int32 (SNN)',
            'data': 'synthetic', 'config': config, 'clocks': clocks, 'gpu_launches': launches // max(1, args.steps),
            'stage_ms': stage_ms, 'kernel_ms': kernel_ms,
            'kernel_ms_fine_pass': {k: {'ms_total': v[0], 'launches': v[1]} for k, v in fine.items()},
            'info': info, 'roofline': roof,
            'roofline_stages': stages_roof, 'cpu_baseline': cpu, 'e2e': e2e}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def roofline(wl, world, kernel_ms, peaks):
    """Dominant kernel = knn_tc_kernel (the FP16 hi/lo distance GEMM + top-m epilogue).
    Algorithmic work per launch: 2 * N_q * N * d flops (true d); duration = the CUDA-event
    pair the library records around the launch on its stream, averaged over the timed steps."""
    n, d = wl['cells'], wl['pca_n']
    nq = n / world
    ms = kernel_ms.get('knn_main')
    if not ms:
        return None
    flops = 2.0 * nq * n * d
    achieved = flops / (ms / 1e3) / 1e12
    peak = peaks.get('bf16_tflops_sustained') or 1400.0
    traffic = None
    try:
        t = json.load(open(os.path.join(ROOT, 'profiles', 'knn_main_traffic.json')))
        if t.get('n_total') == n and t.get('n_query') == int(nq):
            traffic = t['dram_bytes']
    except Exception:
        pass
    return {'bound': 'tensor', 'kernel': 'knn_tc_kernel', 'achieved': achieved, 'peak': peak, 'unit': 'TFLOP/s',
            'frac': achieved / peak, 'traffic': traffic, 'ms_per_launch': ms, 'flops_per_launch': flops,
            'peak_source': ('MEASURED_PEAKS.json bf16_tflops_sustained (kernel timed inside a long step; FP16 and '
                            'bf16 share the tensor rate); 2 FP16 products per useful flop + K padding 103->112 cap '
                            'this fraction at 0.446') if peaks else 'fallback 1.4 PFLOP/s sustained'}


def stage_rooflines(wl, world, fine, info, out, peaks):
    """HBM-bound kernels: achieved GB/s = algorithmic bytes / device ms of the event slot."""
    hbm = peaks.get('hbm_gbs') or 6650.0
    n_loc = wl['cells'] / world
    G, H, k = wl['genes'], wl['hvg_n'], wl['k']
    nnz, nnz_h = info['nnz'], info['nnz_hvg']
    nmult, sum_panel = info['nmult'], info['sum_panel']
    rows = []

    def add(name, slot, nbytes, note=None):
        if slot in fine and fine[slot][0] > 0:
            ms = fine[slot][0]
            row = {'kernel': name, 'bound': 'hbm', 'ms_total': ms, 'launches': fine[slot][1],
                   'bytes': nbytes, 'achieved': nbytes / (ms / 1e3) / 1e9, 'peak': hbm, 'unit': 'GB/s',
                   'frac': nbytes / (ms / 1e3) / 1e9 / hbm}
            if note:
                row['note'] = note
            rows.append(row)
    add('norm_moments_kernel', 'norm_moments', nnz * 8 + (n_loc + 1) * 8 + n_loc * 8 + 2 * G * 8)
    add('hvg_count+hvg_fill', 'hvg_extract', nnz * 8 + nnz * 4 + n_loc * 8 + nnz_h * 12)
    add('atw_compact_kernel (A^T w)', 'irlb_atw', (nmult / 2) * (nnz_h * 12 + n_loc * 24),
        note='algorithmic bytes per SURVEY 8(d) (12 B per nonzero); the kernel reads the compact operand, 4 B per nonzero '
             '+ per-row value tables (ncu: 1.14 GB per launch at C3 against 2.68 GB algorithmic), so this fraction can '
             'exceed what the DRAM actually delivers; the slot includes the one-off re-encoding pass')
    add('av_kernel (A v)', 'irlb_av', (nmult / 2) * (nnz_h * 12 + n_loc * 24))
    add('panel_dot+panel_update', 'irlb_panel', 8 * n_loc * (2 * sum_panel + 3 * (nmult / 2)))
    walk = out['sum_walk']
    e = out['edges']
    add('snn walk kernels (count + staged edges)', 'snn_count', 4 * wl['cells'] * k + 4 * walk + e * 5)
    add('snn compaction (fill)', 'snn_fill', e * 5 + e * 9)
    return rows


if __name__ == '__main__':
    main()
