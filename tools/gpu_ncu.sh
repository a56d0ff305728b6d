#!/bin/bash
# ncu session: launch list of one bench run + full capture of the graph-stage kernels.  Usage: tools/gpu_ncu.sh <tag>
TAG=${1:-r2n}; O=gpurun_out; mkdir -p $O
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra > $O/${TAG}_ncu_launch_run.log 2>&1
# warm-up = 3 forwards, e2e warm-up etc. come later: skip the graph-stage kernels of the first 3 forwards (4-CTA kernel + its 2-CTA tail each)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attend_tc -s 6 -c 2 -o $O/${TAG}_prof_attend -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra > $O/${TAG}_ncu_full_run.log 2>&1
ls -la $O | tail -4
