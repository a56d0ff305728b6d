#!/bin/bash
# One GPU session: parity tests, smoke, bench, traces, ncu.  Usage: tools/gpu_round.sh <tag>
TAG=${1:-rX}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/${TAG}_gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/${TAG}_smoke.log 2>&1; echo "smoke rc=$?" >> $O/${TAG}_smoke.log
timeout 600 python bench.py > $O/${TAG}_bench.log 2>&1
timeout 300 python bench.py --impl tc --no-cpu-baseline > $O/${TAG}_bench_tc.log 2>&1
timeout 300 python tools/launch_timeline.py > $O/${TAG}_timeline.log 2>&1
B=64 HW=72 timeout 300 python tools/launch_timeline.py > $O/${TAG}_timeline_chop64x72.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/${TAG}_ncu_launch_run.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attend_tc4 -s 3 -c 1 -o $O/${TAG}_prof_tc4 -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/${TAG}_ncu_full_run.log 2>&1
tail -3 $O/${TAG}_pytest.log; cat $O/${TAG}_smoke.log; cat $O/${TAG}_bench.log | cut -c1-600; cat $O/${TAG}_timeline.log
