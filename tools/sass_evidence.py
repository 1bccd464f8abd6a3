"""Counts the SASS mnemonics that show which hardware paths each kernel of the built library uses
(UTCMMA/UTCBAR/UTCCP/LDTM/STTM = tcgen05 + tensor memory, UTMALDG/UTMASTG = TMA, SYNCS = mbarrier,
ELECT), per kernel, with registers / shared memory from the resource usage section.
Usage: python tools/sass_evidence.py [path/to/liboat_b200.so] > profiles/rN_sass_evidence.txt"""
import collections
import glob
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WATCH = ("UTCMMA", "UTCHMMA", "UTCQMMA", "UTCBAR", "UTCCP", "LDTM", "STTM", "UTCATOM", "UTMALDG", "UTMASTG",
         "UTMAPF", "UBLKCP", "SYNCS", "ELECT", "HMMA", "IMMA", "FFMA", "LDG", "STG", "LDS", "STS", "RED", "ATOMG")


def demangle(names):
  out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
  return dict(zip(names, out))


def short(name):
  name = name.replace("(anonymous namespace)::", "")
  name = re.sub(r"^void ", "", name)
  depth, out = 0, []
  for ch in name:  # keep template arguments, drop the parameter list
    if ch == "<":
      depth += 1
    elif ch == ">":
      depth -= 1
    elif ch == "(" and depth == 0:
      break
    out.append(ch)
  return "".join(out).replace("oat::", "")


def main():
  so = sys.argv[1] if len(sys.argv) > 1 else sorted(glob.glob(os.path.join(ROOT, "oatomobile_b200", "*.so")))[0]
  sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
  res = subprocess.run(["cuobjdump", "-res-usage", so], capture_output=True, text=True, check=True).stdout
  usage = {}
  cur = None
  for line in res.splitlines():
    m = re.search(r"Function (\S+):", line)
    if m:
      cur = m.group(1)
      continue
    m = re.search(r"REG:(\d+).*?SHARED:(\d+)", line)
    if m and cur:
      usage[cur] = (int(m.group(1)), int(m.group(2)))
  counts = collections.OrderedDict()
  cur = None
  for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
      cur = m.group(1)
      counts[cur] = collections.Counter()
      continue
    if cur is None:
      continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m:
      op = m.group(1)
      counts[cur]["_total"] += 1
      for w in WATCH:
        if op == w or op.startswith(w):
          counts[cur][w] += 1
          break
  names = demangle(list(counts))
  print("# %s — %d kernels (cuobjdump -sass / -res-usage, sm_100a)" % (os.path.basename(so), len(counts)))
  print("# columns: kernel | regs | static smem B | SASS instr | mnemonic counts (non-zero of: %s)" % " ".join(WATCH[:14]))
  rows = []
  for k, c in counts.items():
    regs, smem = usage.get(k, (0, 0))
    hot = " ".join("%s=%d" % (w, c[w]) for w in WATCH[:14] if c[w])
    rows.append((0 if hot else 1, short(names.get(k, k)), regs, smem, c["_total"], hot or "-"))
  for _, n, regs, smem, total, hot in sorted(rows):
    print("%-64s %4d %7d %6d  %s" % (n[:64], regs, smem, total, hot))
  tc = [r for r in rows if "UTCMMA" in r[5] or "UTCHMMA" in r[5]]
  tma = [r for r in rows if "UTMALDG" in r[5] or "UTMASTG" in r[5]]
  print("# kernels issuing tcgen05.mma: %d; kernels using TMA tensor copies: %d" % (len(tc), len(tma)))


if __name__ == "__main__":
  main()
