"""Summarise an .ncu-rep: key raw metrics + top stall lines of the source page (SASS)."""
import csv
import subprocess
import sys

rep = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'smsp__inst_executed.sum',
        'sm__inst_executed_pipe_lsu.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'sm__cycles_elapsed.avg', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed']
for r in rows[2:]:
    print('== kernel:', r[hdr.index('Kernel Name')] if 'Kernel Name' in hdr else '')
    for k in KEYS:
        # some metrics carry a section prefix (TPC.TriageCompute.): exact name first, then a non-empty suffix match
        hit = [i for i, name in enumerate(hdr) if name == k and r[i] != ''] or \
              [i for i, name in enumerate(hdr) if name.endswith('.' + k) and r[i] != '']
        if hit:
            print('  %-75s %s %s' % (k, r[hit[0]], units[hit[0]]))
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
lines = src.splitlines()
# the source page repeats a 2-line header per kernel
i = 0
while i < len(lines):
    if lines[i].startswith('"Kernel Name"'):
        kname = lines[i]
        h = next(csv.reader([lines[i + 1]]))
        j = i + 2
        data = []
        while j < len(lines) and not lines[j].startswith('"Kernel Name"'):
            r = next(csv.reader([lines[j]]))
            data.append(r)
            j += 1
        si = h.index('# Samples')
        so = h.index('Source')
        stall_cols = [c for c in range(len(h)) if h[c].startswith('stall_') and 'Not Issued' not in h[c]]
        tot = sum(float(r[si]) for r in data if r[si].replace('.', '').isdigit()) or 1
        print(kname[:120], 'total samples', tot)
        agg = {}
        for r in data:
            for c in stall_cols:
                try:
                    agg[h[c]] = agg.get(h[c], 0) + float(r[c])
                except ValueError:
                    pass
        print('  stall mix:', ', '.join('%s %.1f%%' % (k, 100 * v / tot) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
        ranked = sorted(range(len(data)), key=lambda x: -float(data[x][si]) if data[x][si].replace('.', '').isdigit() else 0)
        for x in ranked[:topn]:
            r = data[x]
            top = sorted(((float(r[c]), h[c]) for c in stall_cols if r[c].replace('.', '').isdigit()), reverse=True)[:2]
            print('  %5.1f%%  #%4d %-70s %s' % (100 * float(r[si]) / tot, x, r[so].strip()[:70], ' '.join('%s=%d' % (n, v) for v, n in top)))
        i = j
    else:
        i += 1
