#!/bin/bash
# Like exp_knobs.sh, for the knobs given on the command line (each argument one environment setting).
run() { echo "== $*"; env "$@" timeout 200 python tools/exp_series.py 2000 2>&1 | python -c "
import sys
for l in sys.stdin:
    if l.startswith('us/step'):
        v=[float(x) for x in l.split(':')[1].split()]
        print('cycle avg (steps 1000-2000) %.1f us | light (last 5 buckets) %.1f us | heavy (buckets 20-24) %.1f us | buckets 12-13 (driver window) %.1f us' % (sum(v[20:40])/20, sum(v[35:40])/5, sum(v[20:25])/5, sum(v[12:14])/2))
"; }
for k in "$@"; do run $k; done
