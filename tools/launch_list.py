"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv) to ONE front-end pass:
from the first libsize_kernel in the window to the launch before the next one.

    python tools/launch_list.py gpurun_out/launches.csv "<command line that produced it>" > profiles/rN_launches_c3_step.txt
"""
import csv
import re
import sys
from collections import OrderedDict

path, cmd = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else '')
rows = []
with open(path, newline='') as fh:
    lines = [l for l in fh if l.startswith('"')]
rd = csv.reader(lines)
hdr = next(rd)
ik, iv, iu = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
for r in rd:
    if len(r) <= iv:
        continue
    v = float(r[iv].replace(',', ''))
    u = r[iu]
    ms = v / 1e6 if u in ('ns', 'nsecond') else v / 1e3 if u in ('us', 'usecond') else v if u in ('ms', 'msecond') else v * 1e3
    name = re.sub(r'\(.*', '', r[ik]).strip()
    rows.append((name, ms))
starts = [i for i, (n, _) in enumerate(rows) if n.startswith('libsize_kernel')]
if len(starts) >= 2:
    rows = rows[starts[0]:starts[1]]
elif starts:
    rows = rows[starts[0]:]
agg = OrderedDict()
for n, ms in rows:
    a = agg.setdefault(n, [0, 0.0])
    a[0] += 1
    a[1] += ms
tot = sum(a[1] for a in agg.values())
print('# ncu launch list of ONE front-end pass (C3, 1,306,127 x 27,998, one B200)')
print('# command: %s' % cmd)
print('# (the same command exited 0 without ncu first).  One pass = from the first libsize_kernel in the window to the')
print('# launch before the next one.  Per-launch times are cold-cache and serialised: compare SHARES with bench.py, not absolutes.')
print('# launches: %d, summed device time %.1f ms' % (len(rows), tot))
print('%-52s %5s %10s %7s %10s' % ('kernel', 'n', 'total ms', 'share', 'avg us'))
for n, (c, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print('%-52s %5d %10.3f %6.1f%% %10.1f' % (n[:52], c, ms, 100 * ms / tot, 1e3 * ms / c))
