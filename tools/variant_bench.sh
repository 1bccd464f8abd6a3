#!/bin/bash
# A/B harness for the GPU box: times encoder + flow with every tools/variants_*.so and the product lib.
timeout 200 python tools/variant_bench.py 2>&1 | tail -1
for v in tools/variants_*.so; do
  OAT_B200_LIB=$PWD/$v timeout 200 python tools/variant_bench.py 2>&1 | tail -1
done
