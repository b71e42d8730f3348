#!/bin/bash
# round 2, GPU call 16: SCATTER path (resolve with staged gathers, scatter occupancy), fused interpolate->rasterize with 1 / 2 groups in flight
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
timeout -k 5 120 python -c "import __graft_entry__ as g; g.smoke()" || { echo "SMOKE FAILED"; exit 1; }
echo "== whole GPU suite"
timeout -k 10 900 python -m pytest tests -m gpu -x -q --timeout 600 --timeout-method thread 2>&1 | tail -5
run() {
  B200DR_LIB=$2 B200DR_FUSION=$3 timeout -k 5 150 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-extras $4 > gpurun_out/b_$1.json 2>gpurun_out/b_$1.err
  python -c "
import json; d=json.load(open('gpurun_out/b_$1.json')); o=d['ops']; print('$1', round(d['ms_per_step'],3), 'rounds', d['timed_rounds'], 'clk', d['clocks']['sm_mhz'], d['clocks']['samples'], ' '.join('%s=%.3f'%(k.replace('rasterize','ras').replace('interpolate','ip').replace('texture','tex').replace('antialias','aa'),v['ms']) for k,v in o.items() if v['ms']>0.05))" || tail -5 gpurun_out/b_$1.err
}
V=$PWD/nvdiffrast_b200/variants
echo "== c4"
run c4_base "" 1 "--workload c4"
for v in rilp1 rilp4 scm8 scm10 ipcu1; do run c4_$v $V/libb200dr_$v.so 1 "--workload c4"; done
run c4_base2 "" 1 "--workload c4"
echo "== c2"
run c2_base "" 1 "--workload c2"
run c2_ipcu1 $V/libb200dr_ipcu1.so 1 "--workload c2"
