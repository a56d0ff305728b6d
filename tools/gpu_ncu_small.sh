#!/bin/bash
# ncu --set full on the small prologue / epilogue kernels (development aid).  Usage: tools/gpu_ncu_small.sh <tag> [HW]
TAG=${1:-r2s}; HW=${2:-256}; O=gpurun_out; mkdir -p $O
HW=$HW timeout 900 ncu --set full --clock-control none -k regex:"pack_b_gamma|featmap_tc|gather_qpatch|kbar_kernel|embed_tc_kernel|fold_partials|rowmax_tc" -s 22 -c 11 -o $O/${TAG}_small_$HW -f python tools/launch_timeline.py > $O/${TAG}_ncu_small_run.log 2>&1
ls -la $O/${TAG}_small_$HW.ncu-rep
