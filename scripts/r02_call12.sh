#!/bin/bash
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
run() {
  B200DR_LIB=$2 B200DR_FUSION=$3 timeout -k 5 150 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-extras $4 > gpurun_out/b_$1.json 2>gpurun_out/b_$1.err
  python -c "
import json; d=json.load(open('gpurun_out/b_$1.json')); o=d['ops']; print('$1', round(d['ms_per_step'],3), 'rounds', d['timed_rounds'], 'clk', d['clocks']['sm_mhz'], d['clocks']['samples'], ' '.join('%s=%.3f'%(k.replace('rasterize','ras').replace('interpolate','ip').replace('texture','tex').replace('antialias','aa'),v['ms']) for k,v in o.items() if v['ms']>0.05))" || tail -5 gpurun_out/b_$1.err
}
run base "" 1
run base_unfused "" 0
run c5 "" 1 "--workload c5"
timeout -k 5 300 python -m pytest tests/test_fusion_gpu.py tests/test_ops_gpu.py -q -k "fus or texture" --timeout 200 --timeout-method thread 2>&1 | tail -2
