#!/bin/bash
# round 2, GPU call 3: smoke first (bounded), then the GPU test suite, bench A/B of the variants, fusion on/off, triangle sweep.
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
echo "== smoke"; timeout -k 5 240 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 || { echo "SMOKE FAILED/HUNG - stopping"; exit 1; }
echo "== fusion quick"; timeout -k 5 400 python -m pytest tests/test_fusion_gpu.py -x -q --timeout 150 --timeout-method thread 2>&1 | tail -15
echo "== pytest -m gpu"; timeout -k 5 1200 python -m pytest tests -m gpu -q -x --timeout 200 --timeout-method thread --deselect tests/test_fusion_gpu.py 2>&1 | tail -25
run() {
  B200DR_LIB=$2 B200DR_FUSION=$3 timeout -k 5 150 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > gpurun_out/b_$1.json 2>gpurun_out/b_$1.err
  python -c "
import json; d=json.load(open('gpurun_out/b_$1.json')); o=d['ops']; print('$1', round(d['ms_per_step'],3), 'rounds', d['timed_rounds'], 'clk', d['clocks']['sm_mhz'], d['clocks']['samples'], ' '.join('%s=%.3f'%(k.replace('rasterize','ras').replace('interpolate','ip').replace('texture','tex').replace('antialias','aa'),v['ms']) for k,v in o.items() if v['ms']>0.05))" || tail -5 gpurun_out/b_$1.err
}
echo "== bench"
run fused "" 1
run unfused "" 0
for f in $(ls nvdiffrast_b200/variants/*.so 2>/dev/null); do
  n=$(basename $f .so); n=${n#libb200dr_}
  run ${n}_unfused $PWD/$f 0
done
run fused2 "" 1
echo "== triangle sweep"; timeout -k 5 400 python scripts/tri_sweep.py 4 2>&1 | tail -12
