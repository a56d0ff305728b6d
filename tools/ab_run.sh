#!/bin/bash
# A/B of library variants: forward timing + per-launch timeline (development aid)
O=gpurun_out; TAG=${1:-ab}
python tools/ab_variants.py run > $O/${TAG}_ab.log 2>&1; B=64 HW=72 python tools/ab_variants.py run >> $O/${TAG}_ab.log 2>&1
for lib in dagl_b200/libdagl_b200.so dagl_b200/variants/lib*.so; do echo "== $lib" >> $O/${TAG}_ab.log; DAGL_B200_LIB=$PWD/$lib python tools/launch_timeline.py 2>&1 | grep -E "embed|attend|TOTAL" >> $O/${TAG}_ab.log; done
cat $O/${TAG}_ab.log
