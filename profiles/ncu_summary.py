"""Summarise an ncu `--page raw --csv` dump: python profiles/ncu_summary.py raw.csv"""
import csv
import sys

WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "launch__waves_per_multiprocessor",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__inst_executed.sum",
    "sm__cycles_elapsed.avg", "lts__t_sector_hit_rate.pct",
]


def main(path, limit=None):
  rows = list(csv.reader(open(path)))
  hdr, units = rows[0], rows[1]
  idx = {h: i for i, h in enumerate(hdr)}
  for r in rows[2:][:limit]:
    print("----", r[idx["Kernel Name"]][:90], "grid", r[idx["Grid Size"]], "block",
          r[idx["Block Size"]])
    for w in WANT:
      if w in idx:
        print("  %-64s %14s %s" % (w, r[idx[w]], units[idx[w]]))
    st = []
    for h in hdr:
      if h.startswith("smsp__average_warps_issue_stalled") and h.endswith(
          "_per_issue_active.ratio"):
        try:
          st.append((h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")],
                     float(r[idx[h]].replace(",", ""))))
        except ValueError:
          pass
    st.sort(key=lambda x: -x[1])
    print("  stalls/issue:", ", ".join("%s=%.2f" % s for s in st[:7]))


if __name__ == "__main__":
  main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else None)
