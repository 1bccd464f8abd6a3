"""Summarises an `ncu --csv` launch list of one RIP step (tools/r2_step_only.py) per kernel
family (the names bench.py's live profile uses) and writes profiles/traffic.json:
  python tools/ncu_families.py gpurun_out/r2_step_launches.csv [profiles/traffic.json]"""
import csv, json, sys

FAMILIES = [("transform_visual", "transform"), ("stem_kernel", "stem"), ("dw_project_kernel", "fused_dw_project"),
            ("expand_dw", "fused_expand_dw"), ("ExpandDw", "fused_expand_dw"), ("front_kernel", "fused_front"),
            ("tc_pw_gemm", "tc_pw_gemm"), ("dw2_kernel", "depthwise"), ("dw_kernel", "depthwise"),
            ("pool_kernel", "pool"), ("merger_kernel", "merger"), ("flow_tc2_kernel<0>", "flow_sample"),
            ("flow_tc2_kernel<1>", "flow_score"), ("aggregate_kernel", "aggregate")]

def family(name):
  for key, fam in FAMILIES:
    if key in name:
      return fam
  return None

def main():
  rows = list(csv.reader(open(sys.argv[1])))
  hi = [i for i, r in enumerate(rows) if 'Kernel Name' in r][0]
  hdr = rows[hi]
  ki, mi, vi, idi, ui = (hdr.index(k) for k in ('Kernel Name', 'Metric Name', 'Metric Value', 'ID', 'Metric Unit'))
  scale = {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
  launches = {}
  for r in rows[hi + 1:]:
    if len(r) <= vi:
      continue
    launches.setdefault(int(r[idi]), {"name": r[ki]})[r[mi]] = float(r[vi].replace(',', '')) * scale.get(r[ui], 1.0)
  # keep the LAST step only: everything from the last `transform` launch on
  ids = sorted(launches)
  starts = [i for i in ids if family(launches[i]["name"]) == "transform"]
  if starts:
    ids = [i for i in ids if i >= starts[-1]]
  fam = {}
  for i in ids:
    l = launches[i]
    f = family(l["name"])
    if f is None:
      continue
    a = fam.setdefault(f, {"launches": 0, "us": 0.0, "dram_read": 0.0, "dram_write": 0.0})
    a["launches"] += 1
    a["us"] += l.get("gpu__time_duration.sum", 0.0)
    a["dram_read"] += l.get("dram__bytes_read.sum", 0.0)
    a["dram_write"] += l.get("dram__bytes_write.sum", 0.0)
  total = sum(a["us"] for a in fam.values())
  print("%-20s %8s %10s %8s %12s %12s" % ("family", "launches", "us (ncu)", "share", "dram rd MB", "dram wr MB"))
  for f, a in sorted(fam.items(), key=lambda kv: -kv[1]["us"]):
    print("%-20s %8d %10.1f %7.1f%% %12.1f %12.1f" % (f, a["launches"], a["us"], 100 * a["us"] / total,
                                                      a["dram_read"] / 1e6, a["dram_write"] / 1e6))
  print("%-20s %8d %10.1f" % ("total", sum(a["launches"] for a in fam.values()), total))
  if len(sys.argv) > 2:
    out = {f: {"dram_bytes_per_step": a["dram_read"] + a["dram_write"], "launches_per_step": a["launches"],
               "ncu_us_per_step": a["us"], "ncu_share_of_step": a["us"] / total,
               "source": "ncu dram__bytes_read.sum + dram__bytes_write.sum over the launches of one step "
                         "(%s; cold-cache, serialised replay)" % sys.argv[1]} for f, a in fam.items()}
    json.dump(out, open(sys.argv[2], "w"), indent=1, sort_keys=True)

main()
