#!/bin/bash
# Cycle-average device time per step (steps 1000-2000 of tools/exp_series.py) under the tuning knobs.
run() { echo "== $*"; env "$@" timeout 200 python tools/exp_series.py 2000 2>&1 | python -c "
import sys
for l in sys.stdin:
    if l.startswith('us/step'):
        v=[float(x) for x in l.split(':')[1].split()]
        print('cycle avg (steps 1000-2000) %.1f us | light (last 5 buckets) %.1f us | heavy (buckets 20-24) %.1f us' % (sum(v[20:40])/20, sum(v[35:40])/5, sum(v[20:25])/5))
"; }
run RG_NOP=1
run RG_PF_WPB=8
run RG_PF_WPB=4
run RG_PREFETCH_EVERY=1
run RG_PREFETCH_EVERY=3
run RG_SPEC_WARPS=64
run RG_BRANCHES=0
run RG_FAST=0
run RG_SPEC=0
