#!/bin/bash
# Capture every hot kernel ONE AT A TIME (ncu --set full, batch 4 per launch) plus the launch list of a full-size run.
# Usage (GPU box): bash scripts/ncu_all.sh
mkdir -p gpurun_out/ncu
KERNELS="raster_bins_kernel interpolate_fwd_fast texture_fwd_fast aa_discontinuity_kernel aa_copy_kernel texture_grad_fast interpolate_grad_fast raster_grad_kernel"
for k in $KERNELS; do
  BENCH_DEBUG=1 timeout 120 ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -o gpurun_out/ncu/$k \
      python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --batch 4 > gpurun_out/ncu/$k.log 2>&1
  grep -m1 coverage gpurun_out/ncu/$k.log
done
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/ncu/launches_c3.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu/launches_c3.log 2>&1
ls gpurun_out/ncu | head -30
