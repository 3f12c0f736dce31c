#!/bin/bash
# Runs on the GPU box (gpurun): the two ncu passes whose summaries go to profiles/.
#   1. launch list of the bench command (gpu__time_duration.sum per launch, cold cache, serialised)
#   2. --set full capture of one steady-state step (step ~1600) of every kernel on the step path
# Numbers printed by bench.py under ncu are not bench values.
set -u
TAG=${1:-r1}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/${TAG}_launches_raw.csv \
    python bench.py --steps 600 --warmup 100 --no-e2e --no-cpu-baseline > gpurun_out/${TAG}_launches_bench.log 2>&1
ncu --set full --clock-control none --import-source on \
    -k regex:"k_step_scan|k_step_player|k_step_monsters|k_step_gen|k_prefetch|k_step_end" --launch-skip 9000 -c 13 -f \
    -o gpurun_out/${TAG}_step_full python bench.py --steps 1700 --warmup 30 --no-e2e --no-cpu-baseline \
    > gpurun_out/${TAG}_full_bench.log 2>&1
ncu -i gpurun_out/${TAG}_step_full.ncu-rep --page raw --csv > gpurun_out/${TAG}_step_full_raw.csv 2>/dev/null
cp rogue-gym_b200/librogue_b200.so gpurun_out/librogue_b200.profiled.so
ls -la gpurun_out | tail -8
