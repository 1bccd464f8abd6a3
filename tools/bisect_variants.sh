#!/bin/bash
# Runs the small end-to-end step with each library variant (bisect helper for the GPU box).
for v in tools/variants_*.so; do
  echo "=== $v"
  OAT_B200_LIB=$PWD/$v timeout 120 python tools/sanitize_small.py tcgen05 2>&1 | tail -2
done
