#!/bin/bash
# round 2, GPU call 9: full GPU tests on the current build, bench A/B (raster bins occupancy, texture forward unroll / occupancy)
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
echo "== smoke"; timeout -k 5 240 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 || { echo "SMOKE FAILED/HUNG - stopping"; exit 1; }
echo "== pytest -m gpu"; timeout -k 5 1500 python -m pytest tests -m gpu -q --timeout 200 --timeout-method thread 2>&1 | grep -E "^(FAILED|ERROR|E  .*(Error|elements outside|must stay))|passed|failed" | cut -c1-300 | head -40
run() {
  B200DR_LIB=$2 B200DR_FUSION=$3 timeout -k 5 150 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > gpurun_out/b_$1.json 2>gpurun_out/b_$1.err
  python -c "
import json; d=json.load(open('gpurun_out/b_$1.json')); o=d['ops']; print('$1', round(d['ms_per_step'],3), 'rounds', d['timed_rounds'], 'clk', d['clocks']['sm_mhz'], d['clocks']['samples'], ' '.join('%s=%.3f'%(k.replace('rasterize','ras').replace('interpolate','ip').replace('texture','tex').replace('antialias','aa'),v['ms']) for k,v in o.items() if v['ms']>0.05))" || tail -5 gpurun_out/b_$1.err
}
echo "== bench"
run base "" 1
for f in $(ls nvdiffrast_b200/variants/*.so 2>/dev/null); do
  n=$(basename $f .so); n=${n#libb200dr_}
  run ${n} $PWD/$f 1
done
run base2 "" 1
