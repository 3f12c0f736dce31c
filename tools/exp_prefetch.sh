#!/bin/bash
# Knob sweep on the bench workload (run on the GPU box): one line per setting with ms per step and the
# event counters. Usage: bash tools/exp_prefetch.sh "RG_PREFETCH_EVERY=1" "RG_PF_WPB=8 RG_PREFETCH_EVERY=2" ...
# With no arguments it sweeps the settings recorded in DESIGN.md §3.3.
run() {
  echo "== $*"
  env "$@" timeout 300 python bench.py --steps 2000 --warmup 200 --no-cpu-baseline --no-e2e 2>&1 |
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['config']['events_rank0_since_create'], d['config']['state_digest_rank0'])"
}
if [ $# -eq 0 ]; then
  set -- "RG_PREFETCH_EVERY=2" "RG_PREFETCH_EVERY=1" "RG_PREFETCH_EVERY=4" "RG_PF_WPB=8" "RG_PF_WPB=4" "RG_PF_EXCLUSIVE=1" \
         "RG_BG_PRIO=l" "RG_CHUNKS=2" "RG_PLAYER_BLOCKS=4736" "RG_PREFETCH=0" "RG_GRAPH=0"
fi
for setting in "$@"; do run $setting; done
