#!/bin/bash
# per-launch timeline of the ResBlock chain for every library variant (development aid)
O=gpurun_out; TAG=${1:-rbab}
for lib in dagl_b200/libdagl_b200.so dagl_b200/variants/libcv*.so; do
  for m in single pair; do DAGL_B200_LIB=$PWD/$lib timeout 120 python tools/resblock_check.py $m timeline 2>&1 | grep "per-launch" >> $O/${TAG}.log; done
done
cat $O/${TAG}.log
