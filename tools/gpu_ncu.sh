#!/bin/bash
# ncu session: launch list of one bench run + full capture of the dominant kernel.  Usage: tools/gpu_ncu.sh <tag>
TAG=${1:-r2n}; O=gpurun_out; mkdir -p $O
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra > $O/${TAG}_ncu_launch_run.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attend_tc4 -s 3 -c 1 -o $O/${TAG}_prof_tc4 -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra > $O/${TAG}_ncu_full_run.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:embed_tc -s 6 -c 2 -o $O/${TAG}_prof_embed -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra > $O/${TAG}_ncu_embed_run.log 2>&1
ls -la $O | tail -5
