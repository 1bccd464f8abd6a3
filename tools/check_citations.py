"""Checks that every reference citation `path/to/file.py:LINE[-LINE]` in the repo's docs, headers,
sources and tests names a file that exists under the reference checkout and has that many lines.
Only runs where /root/reference exists (the build container).  Usage: python tools/check_citations.py"""
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("OAT_REFERENCE_DIR", "/root/reference")
CITE = re.compile(r"(?<![\w/.])((?:oatomobile|tests|examples)/[\w/.-]+?\.(?:py|md|txt|cfg)):(\d+)(?:[-–](\d+))?")
SHORT = re.compile(r"(?<![\w/.])((?:dim|cil|rip)/(?:model|train|agent)\.py|torch/(?:networks/\w+|types|savers|loggers|transforms)\.py|"
                   r"networks/(?:perception|sequence|mlp)\.py|utils/carla\.py|datasets/carla\.py|perception\.py|sequence\.py|mlp\.py|transforms\.py|savers\.py):(\d+)(?:[-–](\d+))?")
SHORT_ROOTS = ("oatomobile/baselines/torch/", "oatomobile/torch/", "oatomobile/", "oatomobile/torch/networks/",
               "oatomobile/baselines/rulebased/", "oatomobile/core/")
SKIP_DIRS = {".git", "gpurun_out", "baseline", "__pycache__", "build", "_ref", "profiles"}
EXTS = (".py", ".md", ".h", ".cu", ".cuh", ".c", ".sh")


def _lines(path, cache={}):
  if path not in cache:
    try:
      with open(path, "rb") as f:
        cache[path] = f.read().count(b"\n") + 1
    except OSError:
      cache[path] = None
  return cache[path]


def _resolve_short(rel):
  for root in SHORT_ROOTS:
    p = os.path.join(REF, root, rel)
    if os.path.exists(p):
      return p
  return None


def scan():
  bad, total = [], 0
  for dirpath, dirnames, filenames in os.walk(ROOT):
    dirnames[:] = [d for d in dirnames if d not in SKIP_DIRS]
    for fn in filenames:
      if not fn.endswith(EXTS) or fn in ("VERDICT.md", "ADVICE.md", "SURVEY.md", "PAPERS.md", "SNIPPETS.md"):
        continue  # driver-written files are not ours to fix
      src = os.path.join(dirpath, fn)
      try:
        text = open(src, encoding="utf-8", errors="replace").read()
      except OSError:
        continue
      for lineno, line in enumerate(text.splitlines(), 1):
        for rx, short in ((CITE, False), (SHORT, True)):
          for m in rx.finditer(line):
            rel, lo, hi = m.group(1), int(m.group(2)), int(m.group(3) or m.group(2))
            if not short and rel.startswith("tests/") and os.path.exists(os.path.join(ROOT, rel)):
              continue  # our own tests
            path = _resolve_short(rel) if short else os.path.join(REF, rel)
            total += 1
            n = _lines(path) if path else None
            where = "%s:%d" % (os.path.relpath(src, ROOT), lineno)
            if n is None:
              bad.append("%s cites %s — no such reference file" % (where, rel))
            elif hi > n or lo > hi or lo < 1:
              bad.append("%s cites %s:%d-%d — file has %d lines" % (where, rel, lo, hi, n))
  return total, bad


def main():
  if not os.path.isdir(REF):
    print("reference checkout not present; nothing checked")
    return 0
  total, bad = scan()
  for b in bad:
    print(b)
  print("%d citations checked, %d dangling" % (total, len(bad)))
  return 1 if bad else 0


if __name__ == "__main__":
  sys.exit(main())
