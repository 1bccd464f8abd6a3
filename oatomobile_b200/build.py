"""In-tree build of the C-ABI library (`liboat_b200.so`) for sm_100a.

Plain `nvcc -shared` — no torch headers, no JIT cache: the `.so` lives next to
the package so it travels to the GPU box with the repo snapshot.  Run as
`python -m oatomobile_b200.build` or through `__graft_entry__.build()`.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "liboat_b200.so")
SOURCES = ["api.cu", "flow.cu", "aggregate.cu", "encoder.cu", "tc_gemm.cu", "flow_tc.cu", "flow_tc2.cu", "planner.cu", "lidar.cu", "train.cu", "fused.cu", "mlp.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "--expt-relaxed-constexpr",
]


def _nvcc():
  for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
    if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
      return cand
  return "nvcc"


def _stale():
  if not os.path.exists(LIB):
    return True
  t = os.path.getmtime(LIB)
  deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
  deps.append(os.path.join(HERE, "..", "include", "oat_b200.h"))
  return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force: bool = False, verbose: bool = False, defines=(), out: str = None) -> str:
  """`defines` / `out` build an experimental variant next to the product library
  (select it at run time with OAT_B200_LIB=<path>)."""
  if out is None and not defines and not force and not _stale():
    return LIB
  objs = []
  procs = []
  os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
  for src in SOURCES:
    obj = os.path.join(HERE, "build", src.replace(".cu", ".o"))
    if out is not None:
      obj = obj.replace(".o", "." + os.path.basename(out) + ".o")
    cmd = [_nvcc()] + NVCC_FLAGS + ["-D" + d for d in defines] + (
        ["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, src), "-o", obj]
    procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
    objs.append(obj)
  for src, p in procs:
    log, _ = p.communicate()
    if verbose or p.returncode != 0:
      sys.stderr.write(log.decode())
    if p.returncode != 0:
      raise RuntimeError("nvcc failed on %s" % src)
  target = out or LIB
  link = [_nvcc(), "-shared", "-cudart", "static", "-o", target] + objs
  subprocess.check_call(link)
  return target


if __name__ == "__main__":
  print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
