#!/bin/bash
# Round-2 ncu captures (run on the GPU box through gpurun). Numbers printed by bench.py under ncu are not bench values.
#   $1 = tag. Captures --set full of the step kernels in the monster-heavy phase (just after the step-1000 mass
#   reset) and in the light phase (step ~1900), plus the launch list of a short bench run.
set -u
TAG=${1:-r2}
B="python bench.py --burn-in 0 --no-e2e --no-extras --no-cpu-baseline --no-oracle-check"
mkdir -p gpurun_out
# heavy phase: step ~1045
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_step_monsters|k_step_player" --launch-skip 4180 -c 4 -f \
    -o gpurun_out/${TAG}_heavy_monpl $B --steps 20 --warmup 1040 > gpurun_out/${TAG}_heavy_monpl.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_step_fast|k_step_scan|k_step_gen" --launch-skip 4180 -c 4 -f \
    -o gpurun_out/${TAG}_heavy_fast $B --steps 20 --warmup 1040 > gpurun_out/${TAG}_heavy_fast.log 2>&1
# light phase: step ~1900
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_step_monsters|k_step_player" --launch-skip 7600 -c 4 -f \
    -o gpurun_out/${TAG}_light_monpl $B --steps 20 --warmup 1895 > gpurun_out/${TAG}_light_monpl.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_step_fast|k_step_scan|k_step_gen" --launch-skip 7600 -c 4 -f \
    -o gpurun_out/${TAG}_light_fast $B --steps 20 --warmup 1895 > gpurun_out/${TAG}_light_fast.log 2>&1
# only the CSV pages travel back (gpurun_out is capped at 64 MiB): raw metrics and the per-instruction source page
for f in heavy_monpl heavy_fast light_monpl light_fast; do
  ncu -i gpurun_out/${TAG}_$f.ncu-rep --page raw --csv > gpurun_out/${TAG}_${f}_raw.csv 2>/dev/null
  ncu -i gpurun_out/${TAG}_$f.ncu-rep --page source --csv --print-source sass > gpurun_out/${TAG}_${f}_source.csv 2>/dev/null
  rm -f gpurun_out/${TAG}_$f.ncu-rep
done
gzip -f gpurun_out/${TAG}_*_source.csv
ls -la gpurun_out | grep ${TAG}
