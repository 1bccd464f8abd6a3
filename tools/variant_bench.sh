#!/bin/bash
# A/B harness for the GPU box: times encoder + flow (tools/variant_bench.py: encode ms, flow ms,
# parity of z / q against the FP32 SIMT path) with the product library and with every variant
# build under oatomobile_b200/variants/ (made with oatomobile_b200.build.build(defines=..., out=...)).
mkdir -p gpurun_out
{
timeout 200 python tools/variant_bench.py 2>&1 | tail -1
for v in oatomobile_b200/variants/*.so; do
  OAT_B200_LIB=$PWD/$v timeout 200 python tools/variant_bench.py 2>&1 | tail -1
done
} 2>&1 | tee gpurun_out/variant_bench.log
