#!/bin/bash
# One GPU session (round 2): parity tests, smoke, bench, warm timelines.  Usage: tools/gpu_round2.sh <tag> [quick]
TAG=${1:-r2x}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/${TAG}_gpu.txt 2>&1
nproc >> $O/${TAG}_gpu.txt
timeout 1500 python -m pytest tests -m gpu -q -s --durations=15 > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/${TAG}_smoke.log 2>&1; echo "smoke rc=$?" >> $O/${TAG}_smoke.log
timeout 900 python bench.py > $O/${TAG}_bench.log 2>&1
timeout 300 python tools/launch_timeline.py > $O/${TAG}_timeline.log 2>&1
HW=64 timeout 300 python tools/launch_timeline.py > $O/${TAG}_timeline_64.log 2>&1
B=64 HW=72 timeout 300 python tools/launch_timeline.py > $O/${TAG}_timeline_chop64x72.log 2>&1
HEADS=4 HW=64 timeout 300 python tools/launch_timeline.py > $O/${TAG}_timeline_stage_64.log 2>&1
HEADS=4 B=64 HW=72 timeout 300 python tools/launch_timeline.py > $O/${TAG}_timeline_stage_chop64x72.log 2>&1
HEADS=4 timeout 300 python tools/launch_timeline.py > $O/${TAG}_timeline_stage_256.log 2>&1
grep -E "passed|failed|error" $O/${TAG}_pytest.log | tail -5; cat $O/${TAG}_smoke.log; cut -c1-1500 $O/${TAG}_bench.log; cat $O/${TAG}_timeline.log $O/${TAG}_timeline_64.log $O/${TAG}_timeline_stage_64.log
python tools/ces_breakdown.py > $O/${TAG}_ces256.log 2>&1; B=64 HW=72 python tools/ces_breakdown.py > $O/${TAG}_ces_chop.log 2>&1; HW=64 python tools/ces_breakdown.py > $O/${TAG}_ces64.log 2>&1
cat $O/${TAG}_ces256.log $O/${TAG}_ces_chop.log $O/${TAG}_ces64.log
# ResBlock chain kernel: parity vs fp64 + timing vs torch/cuDNN, per-launch timeline, one ncu capture of the pair kernel
timeout 300 python tools/resblock_check.py pair time > $O/${TAG}_resblock_pair.log 2>&1; timeout 300 python tools/resblock_check.py single time > $O/${TAG}_resblock_single.log 2>&1
for m in single pair; do timeout 120 python tools/resblock_check.py $m timeline 2>&1 | grep "per-launch" >> $O/${TAG}_resblock_timeline.log; done
timeout 400 ncu --set full --clock-control none --import-source on -k regex:conv64 -s 4 -c 2 -f -o $O/${TAG}_conv64_pair_chop python tools/resblock_ncu.py pair 64 72 72 > $O/${TAG}_ncu_conv64.log 2>&1
tail -12 $O/${TAG}_resblock_pair.log; tail -9 $O/${TAG}_resblock_single.log; cat $O/${TAG}_resblock_timeline.log
