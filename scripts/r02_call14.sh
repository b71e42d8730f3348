#!/bin/bash
# round 2, GPU call 14: fused interpolate -> rasterize backward (parity, c2 / c4 timings, occupancy variants); shrink-branch variant
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
timeout -k 5 120 python -c "import __graft_entry__ as g; g.smoke()" || { echo "SMOKE FAILED"; exit 1; }
echo "== fusion tests"
timeout -k 10 600 python -m pytest tests/test_fusion_gpu.py -x -q --timeout 200 --timeout-method thread 2>&1 | tail -8
echo "== full-size gradients vs the race-free reference (c2, c4 chains now run the fused node)"
timeout -k 10 900 python -m pytest tests/test_grad_reference_gpu.py -x -q -k "c2 or c4" --timeout 600 --timeout-method thread 2>&1 | tail -5
run() {
  B200DR_LIB=$2 B200DR_FUSION=$3 timeout -k 5 150 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-extras $4 > gpurun_out/b_$1.json 2>gpurun_out/b_$1.err
  python -c "
import json; d=json.load(open('gpurun_out/b_$1.json')); o=d['ops']; print('$1', round(d['ms_per_step'],3), 'rounds', d['timed_rounds'], 'clk', d['clocks']['sm_mhz'], d['clocks']['samples'], ' '.join('%s=%.3f'%(k.replace('rasterize','ras').replace('interpolate','ip').replace('texture','tex').replace('antialias','aa'),v['ms']) for k,v in o.items() if v['ms']>0.05))" || tail -5 gpurun_out/b_$1.err
}
echo "== c2 / c4: fused vs op-by-op"
run c2_fused "" 1 "--workload c2"
run c2_unfused "" 0 "--workload c2"
run c4_fused "" 1 "--workload c4"
run c4_unfused "" 0 "--workload c4"
for v in ipc2 ipc4; do
  run c2_$v $PWD/nvdiffrast_b200/variants/libb200dr_$v.so 1 "--workload c2"
  run c4_$v $PWD/nvdiffrast_b200/variants/libb200dr_$v.so 1 "--workload c4"
done
echo "== c3: base / shrink-branch"
run c3_base "" 1
run c3_shr $PWD/nvdiffrast_b200/variants/libb200dr_shr.so 1
run c3_base2 "" 1
echo "== whole GPU suite"
timeout -k 10 900 python -m pytest tests -m gpu -x -q --timeout 600 --timeout-method thread 2>&1 | tail -5
