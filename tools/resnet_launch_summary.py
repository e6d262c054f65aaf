"""Per-kernel summary of an ncu launch list of tools/profile_resnet.py (last forward only)."""
import csv
import sys
from collections import OrderedDict, defaultdict

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if not l.startswith('=='))]
hdr = rows[0]
ki, mi, vi, ii = hdr.index('Kernel Name'), hdr.index('Metric Name'), hdr.index('Metric Value'), hdr.index('ID')
recs = OrderedDict()
for r in rows[1:]:
    if len(r) <= vi:
        continue
    d = recs.setdefault(r[ii], {'k': r[ki]})
    try:
        d[r[mi]] = float(r[vi].replace(',', ''))
    except ValueError:
        pass
fw = [d for d in recs.values() if 'gpu__time_duration.sum' in d and ('toad' in d['k'] or 'tc::' in d['k'] or 'resnet::' in d['k'] or 'stem' in d['k'] or 'halo' in d['k'])]
half = len(fw) // 2
fw = fw[half:]
tot = sum(d['gpu__time_duration.sum'] for d in fw)
print('# last forward: %d launches, %.1f us (ncu, serialised)' % (len(fw), tot / 1e3))
# one forward = from the stem (im2col or fused stem kernel) to the average pool: DRAM bytes per image of that span
starts = [i for i, d in enumerate(fw) if 'stem' in d['k']]
if starts:
    one = fw[starts[-1]:]
    by = sum(d.get('dram__bytes_read.sum', 0) + d.get('dram__bytes_write.sum', 0) for d in one)
    t1 = sum(d['gpu__time_duration.sum'] for d in one)
    batch = int(sys.argv[3]) if len(sys.argv) > 3 else 128
    print('# one forward (stem .. avgpool): %d launches, %.1f us, %.1f MB of DRAM traffic = %.1f MB per image at batch %d'
          % (len(one), t1 / 1e3, by / 1e6, by / 1e6 / batch, batch))
if len(sys.argv) > 2:   # per-launch listing
    for d in fw:
        t = d['gpu__time_duration.sum']
        by = d.get('dram__bytes_read.sum', 0) + d.get('dram__bytes_write.sum', 0)
        print('%-58s t=%7.1fus dram=%7.1fMB bw=%5.2fTB/s tensor=%4.1f%%' % (d['k'][:58], t / 1e3, by / 1e6, by / t / 1e3, d.get('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 0)))
agg = defaultdict(lambda: [0, 0, 0, 0])
for d in fw:
    a = agg[d['k'][:60]]
    a[0] += 1
    a[1] += d['gpu__time_duration.sum']
    a[2] += d.get('dram__bytes_read.sum', 0) + d.get('dram__bytes_write.sum', 0)
    a[3] += d.get('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 0) * d['gpu__time_duration.sum']
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print('%-62s n=%3d t=%8.1fus share=%5.1f%% dram=%7.1fMB bw=%5.2fTB/s tensor=%4.1f%%' % (k, a[0], a[1] / 1e3, 100 * a[1] / tot, a[2] / 1e6, a[2] / a[1] / 1e3, a[3] / a[1]))
