#!/bin/bash
# GPU call: quick parity + timing of the fused kernels (fusion bench, per-kernel ncu durations).
mkdir -p gpurun_out
{
echo "=== pytest fused"; timeout 300 python -m pytest tests/test_gpu_fused.py -x -q 2>&1 | tail -4
echo "=== transform"; timeout 100 python - <<'PY'
import torch, sys
sys.path.insert(0, ".")
from oatomobile_b200 import ops
x = torch.rand(256, 4, 200, 200, device="cuda")
xh = x.permute(0, 2, 3, 1).contiguous()
for name, fn in (("tiled nchw", lambda: ops.transform_visual(x)), ("gather hwc", lambda: ops.transform_visual_hwc(xh))):
  for _ in range(3): fn()
  torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(20): fn()
  e1.record(); torch.cuda.synchronize()
  print("TRANSFORM %s: %.1f us" % (name, 1e3 * e0.elapsed_time(e1) / 20))
PY
echo "=== fusion bench"; timeout 200 python tools/fusion_bench.py 2>&1 | tail -8
echo "=== per-kernel durations (ncu, cold)"
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"expand_dw|front_kernel|dw_project|stem_kernel" -c 5 --csv python tools/fusion_bench.py --once 30 2>&1 | grep -o 'ExpandDw[A-Za-z]*<[^>]*>\|front_kernel\|dw_project_kernel\|stem_kernel\|"gpu__time_duration.sum","[a-z]*","[0-9.,]*"' | paste - - | cut -c1-160
} > gpurun_out/fused_quick.log 2>&1
tail -40 gpurun_out/fused_quick.log
