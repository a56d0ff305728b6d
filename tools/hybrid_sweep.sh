#!/bin/bash
# share of the keys given to the concurrent 2-CTA tail launch (DAGL_HYBRID permille; 0 = off, unset = model): forward timing
O=gpurun_out; TAG=${1:-hyb}; shift
for h in 0 "" "$@"; do
  if [ -z "$h" ]; then echo "== model" >> $O/${TAG}.log; python tools/ab_variants.py one >> $O/${TAG}.log 2>&1
  else echo "== DAGL_HYBRID=$h" >> $O/${TAG}.log; DAGL_HYBRID=$h python tools/ab_variants.py one >> $O/${TAG}.log 2>&1; fi
done
cat $O/${TAG}.log
