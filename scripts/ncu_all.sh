#!/bin/bash
# Capture every hot kernel ONE AT A TIME (ncu --set full, batch 4 per launch) plus the launch list of a full-size run.
# Usage (GPU box): bash scripts/ncu_all.sh [outdir]         then here: python scripts/ncu_summary.py r02
OUT=${1:-gpurun_out/ncu}
mkdir -p $OUT
KERNELS="raster_bins_kernel interpolate_fwd_fast texture_fwd_fast aa_discontinuity_kernel aa_copy_kernel texture_chain_grad_fast"
UNFUSED="texture_grad_fast interpolate_grad_fast raster_grad_kernel"
for k in $KERNELS; do
  timeout -k 5 150 ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -o $OUT/$k \
      python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-extras --min-seconds 0 --batch 4 > $OUT/$k.log 2>&1
done
for k in $UNFUSED; do
  B200DR_FUSION=0 timeout -k 5 150 ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -o $OUT/$k \
      python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-extras --min-seconds 0 --batch 4 > $OUT/$k.log 2>&1
done
timeout -k 5 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_c3.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-extras --min-seconds 0 > $OUT/launches_c3.log 2>&1
ls $OUT | head -40
