#!/bin/bash
run() { echo "== $*"; env "$@" timeout 100 python tools/exp_e2e.py 2>&1 | grep "rg_step_mirror"; }
run RG_NOP=1
run RG_MIRROR_PRIO=l
run RG_MIRROR_BLOCKS=148
run RG_MIRROR_BLOCKS=296
run RG_MIRROR_BLOCKS=592
run RG_MIRROR_BLOCKS=148 RG_MIRROR_PRIO=l
run RG_MIRROR_BLOCKS=2368
