"""Prints a per-launch table from an `ncu --csv --metrics ...` log (tools/r2_run14.sh)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if 'Kernel Name' in r][0]
hdr = rows[hi]
ki, mi, vi, idi, gi, ui = (hdr.index(k) for k in ('Kernel Name', 'Metric Name', 'Metric Value', 'ID', 'Grid Size', 'Metric Unit'))
d = {}
for r in rows[hi + 1:]:
  if len(r) <= vi:
    continue
  v = float(r[vi].replace(',', ''))
  u = r[ui]
  scale = {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 'byte': 1e-6, 'Kbyte': 1e-3, 'Mbyte': 1.0, 'Gbyte': 1e3}.get(u, 1.0)
  d.setdefault(int(r[idi]), {'name': r[ki].split('(')[0].split('::')[-1][:30], 'grid': r[gi]})[r[mi]] = v * scale
tot = 0.0
for i in sorted(d):
  x = d[i]
  t = x['gpu__time_duration.sum']
  tot += t
  print('%3d %-30s %-16s %7.1f us  rd %7.1f MB  wr %7.1f MB  tensor %4.1f%%  l2hit %4.1f%%' % (
      i, x['name'], x['grid'], t, x.get('dram__bytes_read.sum', 0), x.get('dram__bytes_write.sum', 0),
      x.get('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 0), x.get('lts__t_sector_hit_rate.pct', 0)))
print('total %.1f us over %d launches' % (tot, len(d)))
