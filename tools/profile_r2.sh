#!/bin/bash
# Round-2 ncu captures (run on the GPU box through gpurun). Numbers printed by bench.py under ncu are not bench values.
#   $1 = tag. Only CSV pages travel back (gpurun_out is capped at 64 MiB).
#   1. launch list of a short bench run (gpu__time_duration.sum per launch; cold cache, serialised)
#   2. --set full of the step kernels in the light phase (step ~1900 of the 1000-step cycle) and in the
#      monster-heavy phase just after the step-1000 mass reset (step ~1045)
#   3. dram bytes of one light-phase step with --replay-mode application (kernel replay restores 4 GB of arena
#      between passes and charges the evictions of the restored lines to the kernel: its dram write numbers are not usable)
set -u
TAG=${1:-r2}
B="python bench.py --burn-in 0 --no-e2e --no-extras --no-cpu-baseline --no-oracle-check"
K="k_step_scan|k_step_fast|k_step_player|k_step_monsters|k_step_gen"
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 16000 --csv --log-file gpurun_out/${TAG}_launches_raw.csv \
    $B --steps 300 --warmup 900 > gpurun_out/${TAG}_launches_bench.log 2>&1
# launches of the five step kernels per step: scan 1, fast 1, player 2, monsters 2, gen 2 = 8
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$K" --launch-skip $((8 * 1045)) -c 8 -f \
    -o gpurun_out/${TAG}_heavy $B --steps 20 --warmup 1040 > gpurun_out/${TAG}_heavy.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$K" --launch-skip $((8 * 1900)) -c 8 -f \
    -o gpurun_out/${TAG}_light $B --steps 20 --warmup 1895 > gpurun_out/${TAG}_light.log 2>&1
timeout 900 ncu --replay-mode application --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
    -k regex:"$K" --launch-skip $((8 * 1900)) -c 8 --csv --log-file gpurun_out/${TAG}_light_dram_appreplay.csv \
    $B --steps 20 --warmup 1895 > gpurun_out/${TAG}_light_dram.log 2>&1
# encoders and the host mirror kernel, one launch each (tools/bench_encode.py runs both encoders; exp_e2e the mirror)
timeout 300 ncu --set full --clock-control none -k regex:"k_encode|k_mirror|k_reset|k_prefetch|k_spec_build" -c 12 -f \
    -o gpurun_out/${TAG}_other python tools/profile_other.py > gpurun_out/${TAG}_other.log 2>&1
for f in heavy light other; do
  ncu -i gpurun_out/${TAG}_$f.ncu-rep --page raw --csv > gpurun_out/${TAG}_${f}_raw.csv 2>/dev/null
  if [ "$f" != other ]; then
    ncu -i gpurun_out/${TAG}_$f.ncu-rep --page source --csv --print-source sass > gpurun_out/${TAG}_${f}_source.csv 2>/dev/null
  fi
  rm -f gpurun_out/${TAG}_$f.ncu-rep
done
gzip -f gpurun_out/${TAG}_*_source.csv
ls -la gpurun_out | grep ${TAG}
