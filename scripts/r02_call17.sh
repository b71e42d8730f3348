#!/bin/bash
# round 2, GPU call 17: round-end evidence of the final build + ncu of the c2 / c4 kernels.  Only summaries and two reports travel
# back (gpurun merges at most 64 MiB into gpurun_out/).
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
bash scripts/final_run.sh
mkdir -p gpurun_out/ncu_c4
for k in raster_scatter_kernel raster_resolve_kernel interpolate_grad_fast interpolate_fwd_fast; do
  timeout -k 5 200 ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -f -o gpurun_out/ncu_c4/$k \
      python bench.py --workload c4 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-extras --min-seconds 0 --batch 2 > gpurun_out/ncu_c4/$k.log 2>&1
done
timeout -k 5 200 ncu --set full --clock-control none --import-source on -k regex:interpolate_grad_fast -s 3 -c 1 -f -o gpurun_out/ncu_c4/interpolate_chain_c2 \
      python bench.py --workload c2 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-extras --min-seconds 0 --batch 16 > gpurun_out/ncu_c4/ipc_c2.log 2>&1
python scripts/ncu_c2c4_summary.py r02 > /dev/null 2>&1 && cp profiles/r02_ncu_c2c4.md gpurun_out/final/profiles/
mkdir -p gpurun_out/keep; mv gpurun_out/ncu/texture_chain_grad_fast.ncu-rep gpurun_out/ncu/raster_bins_kernel.ncu-rep gpurun_out/keep/ 2>/dev/null
rm -f gpurun_out/ncu/*.ncu-rep gpurun_out/ncu_c4/*.ncu-rep
du -sh gpurun_out
