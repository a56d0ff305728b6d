#!/bin/bash
# A/B kernel timing of several builds of the library in one GPU session (development aid).
# usage: tools/ab_bench.sh tag libA.so libB.so ...   (paths relative to the repo root)
TAG=$1; shift
for rep in 1 2; do
  for L in "$@"; do
    DAGL_B200_LIB=$PWD/$L python bench.py --steps 30 --no-cpu-baseline 2>&1 | python -c "
import sys,json
for line in sys.stdin:
    if line.startswith('{'):
        d=json.loads(line); print('$L', 'step_ms=%.4f kernel_ms=%.4f e2e_ms=%.4f' % (d['ms_per_step'], d['roofline']['kernel_ms'], d['e2e']['ms_per_step']))
" | tee -a gpurun_out/${TAG}_ab.log
  done
done
