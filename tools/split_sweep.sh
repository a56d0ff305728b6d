#!/bin/bash
# nsplit cost-model sweep at small shapes (development aid)
for c in 0.5 2 4 8; do
  for hw in 64 100; do
    echo "DAGL_SPLIT_COST=$c HW=$hw"; DAGL_SPLIT_COST=$c HW=$hw python tools/ab_variants.py one
  done
  DAGL_SPLIT_COST=$c B=4 HW=64 python tools/ab_variants.py one
done
