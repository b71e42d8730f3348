#!/bin/bash
# round 2, GPU call 15: bin occupancy mask (parity + A/B), fused interpolate->rasterize occupancy, ncu of the c4 kernels
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
timeout -k 5 120 python -c "import __graft_entry__ as g; g.smoke()" || { echo "SMOKE FAILED"; exit 1; }
echo "== whole GPU suite (bin mask in the default build)"
timeout -k 10 900 python -m pytest tests -m gpu -x -q --timeout 600 --timeout-method thread 2>&1 | tail -5
run() {
  B200DR_LIB=$2 B200DR_FUSION=$3 timeout -k 5 150 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-extras $4 > gpurun_out/b_$1.json 2>gpurun_out/b_$1.err
  python -c "
import json; d=json.load(open('gpurun_out/b_$1.json')); o=d['ops']; print('$1', round(d['ms_per_step'],3), 'rounds', d['timed_rounds'], 'clk', d['clocks']['sm_mhz'], d['clocks']['samples'], ' '.join('%s=%.3f'%(k.replace('rasterize','ras').replace('interpolate','ip').replace('texture','tex').replace('antialias','aa'),v['ms']) for k,v in o.items() if v['ms']>0.05))" || tail -5 gpurun_out/b_$1.err
}
V=$PWD/nvdiffrast_b200/variants
echo "== bin mask A/B"
run c3_mask "" 1
run c3_nomask $V/libb200dr_nomask.so 1
run c3_mask2 "" 1
run c2_mask "" 1 "--workload c2"
run c2_nomask $V/libb200dr_nomask.so 1 "--workload c2"
echo "== fused interpolate->rasterize: 5 CTAs/SM, 2 groups"
for v in ipc5 ipcu2; do
  run c2_$v $V/libb200dr_$v.so 1 "--workload c2"
  run c4_$v $V/libb200dr_$v.so 1 "--workload c4"
done
run c4_base "" 1 "--workload c4"
echo "== ncu of the c4 kernels (2 images per launch)"
mkdir -p gpurun_out/ncu_c4
for k in raster_scatter_kernel raster_resolve_kernel interpolate_grad_fast interpolate_fwd_fast; do
  timeout -k 5 200 ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -o gpurun_out/ncu_c4/$k \
      python bench.py --workload c4 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-extras --min-seconds 0 --batch 2 > gpurun_out/ncu_c4/$k.log 2>&1
done
timeout -k 5 200 ncu --set full --clock-control none --import-source on -k regex:raster_bins_kernel -s 3 -c 1 -o gpurun_out/ncu_c4/raster_bins_kernel_c3 \
      python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-extras --min-seconds 0 --batch 4 > gpurun_out/ncu_c4/raster_bins_c3.log 2>&1
timeout -k 5 200 ncu --set full --clock-control none --import-source on -k regex:interpolate_grad_fast -s 3 -c 1 -o gpurun_out/ncu_c4/interpolate_chain_c2 \
      python bench.py --workload c2 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-extras --min-seconds 0 --batch 16 > gpurun_out/ncu_c4/ipc_c2.log 2>&1
ls gpurun_out/ncu_c4
