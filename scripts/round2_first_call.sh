#!/bin/bash
# First GPU call of round 2: validate and time the two experiments that were written but not measured in round 1.
#   here (build container):  python scripts/build_variants.py tile=-DTEXG_TILE2D=1 fixed=-DIPG_FIXED=1 rgc=-DRG_COND_LOADS=1 && rm -rf nvdiffrast_b200/csrc/_build/var_*
#   then:                    gpurun --timeout 900 -- 'bash scripts/round2_first_call.sh'
# Each variant must pass the GPU parity tests of the ops it touches before its timing means anything.
run() {
  B200DR_LIB=$2 timeout -k 5 100 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/b_var.json 2>gpurun_out/b_var.err
  python -c "
import json; d=json.load(open('gpurun_out/b_var.json')); o=d['ops']; print('$1', round(d['ms_per_step'],3), ' '.join('%s=%.3f'%(k.replace('rasterize','ras').replace('interpolate','ip').replace('texture','tex').replace('antialias','aa'),v['ms']) for k,v in o.items() if v['ms']>0.1))" || tail -5 gpurun_out/b_var.err
}
run base ""
for f in $(ls nvdiffrast_b200/variants/*.so 2>/dev/null); do
  n=$(basename $f .so); n=${n#libb200dr_}
  echo "== $n"
  B200DR_LIB=$PWD/$f timeout -k 5 300 python -m pytest tests -m gpu -x -q -k "texture or interpolate or rasterize_grad or chain or fullsize" 2>&1 | tail -2
  run $n $PWD/$f
done
