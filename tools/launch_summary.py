"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel."""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1], errors="ignore")))
hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
h = rows[hi]
kn, mv = h.index("Kernel Name"), h.index("Metric Value")
L = [(r[kn], float(r[mv].replace(",", "")) / 1000) for r in rows[hi + 1:] if len(r) > mv]
agg = collections.defaultdict(lambda: [0, 0.0, []])
for k, t in L:
  m = re.search(r"run_functor<oat::train::(\w+)>", k) or re.search(r"run_functor<oat::(\w+)", k) or \
      re.search(r"oat::(?:<unnamed>::)?(\w+)", k)
  a = agg[m.group(1) if m else k[:48]]
  a[0] += 1
  a[1] += t
  a[2].append(t)
print("%d launches, %.1f us in total" % (len(L), sum(t for _, t in L)))
for k, (c, t, l) in sorted(agg.items(), key=lambda x: -x[1][1]):
  l = sorted(l)
  print("%-28s n=%4d total %8.1f us  median %7.1f  max %7.1f" % (k, c, t, l[len(l) // 2], l[-1]))
