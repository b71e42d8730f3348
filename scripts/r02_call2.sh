#!/bin/bash
# round 2, GPU call 2: full GPU test suite on the new default build, then bench A/B of the variants.
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -25
run() {
  B200DR_LIB=$2 timeout -k 5 100 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/b_$1.json 2>gpurun_out/b_$1.err
  python -c "
import json; d=json.load(open('gpurun_out/b_$1.json')); o=d['ops']; print('$1', round(d['ms_per_step'],3), ' '.join('%s=%.3f'%(k.replace('rasterize','ras').replace('interpolate','ip').replace('texture','tex').replace('antialias','aa'),v['ms']) for k,v in o.items() if v['ms']>0.1))" || tail -5 gpurun_out/b_$1.err
}
echo "== bench"
run base ""
for f in $(ls nvdiffrast_b200/variants/*.so 2>/dev/null); do
  n=$(basename $f .so); n=${n#libb200dr_}
  run $n $PWD/$f
done
run base2 ""
