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
