#!/bin/bash
# GPU call: A/B of fused-kernel build variants (oatomobile_b200/variants/*.so): per-kernel
# durations of one encode with every fused kernel on (ncu launch list, gpu__time_duration).
mkdir -p gpurun_out
{
for so in default oatomobile_b200/variants/*.so; do
  echo "=== $so"
  if [ "$so" != default ]; then export OAT_B200_LIB=$PWD/$so; fi
  timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"expand_dw|front_kernel" -c 4 --csv python tools/fusion_bench.py --once 15 2>&1 | grep -o '"[^"]*ExpandDw[^"]*\|"gpu__time_duration.sum","[a-z]*","[0-9.,]*"' | paste - - | sed 's/oat::fused:://g' | cut -c1-160
done
} > gpurun_out/variants.log 2>&1
tail -40 gpurun_out/variants.log
