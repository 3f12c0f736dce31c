#!/bin/bash
# Runs tools/exp_host_rate.py over prebuilt library variants (rogue-gym_b200/build/variants/*.so) and knobs.
V=rogue-gym_b200/build/variants
run() { echo "== $*"; env "$@" timeout 120 python tools/exp_host_rate.py 1200 2>&1 | head -1; }
run ROGUE_B200_LIB=$PWD/$V/v_mon16.so
run ROGUE_B200_LIB=$PWD/$V/v_mon16.so RG_BRANCHES=0
run ROGUE_B200_LIB=$PWD/$V/v_mon20.so RG_MON_WARPS=2960
run ROGUE_B200_LIB=$PWD/$V/v_mon24.so RG_MON_WARPS=3552
run ROGUE_B200_LIB=$PWD/$V/v_mon32.so RG_MON_WARPS=4736
run ROGUE_B200_LIB=$PWD/$V/v_mon32.so RG_MON_WARPS=4736 RG_BRANCHES=0
run ROGUE_B200_LIB=$PWD/$V/v_mon24.so RG_MON_WARPS=3552 RG_BRANCHES=0
run ROGUE_B200_LIB=$PWD/$V/v_mon16.so RG_MON_WARPS=1184
run ROGUE_B200_LIB=$PWD/$V/v_mon16.so RG_PREFETCH_EVERY=1
run ROGUE_B200_LIB=$PWD/$V/v_mon16.so RG_PREFETCH_EVERY=4
