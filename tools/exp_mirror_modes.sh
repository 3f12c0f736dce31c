#!/bin/bash
# Host-facing step (rg_step_mirror, one call per step) under the host mirror's knobs given on the command line (each
# argument one environment setting, e.g. RG_MIRROR_SMS=24 RG_MIRROR_MODE=direct RG_MIRROR_FAST_FIRST=0): wall clock per
# step (tools/exp_e2e.py) and one timeline line (tools/timeline.py, TL_MIRROR=1). Results: profiles/r2_mirror_modes.log
run() {
  echo "== $*"
  env "$@" timeout 100 python tools/exp_e2e.py 2>&1 | grep "rg_step_mirror"
  env "$@" TL_MIRROR=1 timeout 90 python tools/timeline.py 700 2>&1 | grep "^slot" | tail -2 | head -1 | cut -c1-900
}
for k in "$@"; do run $k; done
