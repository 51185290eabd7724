#!/bin/bash
# A/B timing of library variants with the torch-free harness: B=64 overlapped (16 output buffers) and B=512 back-to-back
for round in 1 2; do
for d in tools/variants/*/; do
  n=$(basename $d)
  o=$(LD_LIBRARY_PATH=$d tools/harness 64 128 3 0 2 1 1 16 | grep overlapped | tail -1 | awk '{print $3}')
  b=$(LD_LIBRARY_PATH=$d tools/harness 512 128 3 0 1 1 0 1 | grep back-to-back | awk '{print $3}')
  echo "$n: B=64 overlapped $o us  B=512 $b us"
done
done
