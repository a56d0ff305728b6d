#!/bin/bash
# ResBlock chain: parity + timing for both kernels, strip (row reuse) vs flat decomposition (development aid)
O=gpurun_out; TAG=${1:-rbab}
for m in single pair; do
  timeout 200 python tools/resblock_check.py $m time > $O/${TAG}_$m.log 2>&1; echo "$m rc=$?"; tail -16 $O/${TAG}_$m.log
  DAGL_CONV_FLAT=1 timeout 200 python tools/resblock_check.py $m time 2>&1 | grep "chain" | sed 's/^/FLAT /' | tee -a $O/${TAG}_$m.log
  timeout 120 python tools/resblock_check.py $m timeline 2>&1 | grep "per-launch" | tee -a $O/${TAG}_$m.log
done
