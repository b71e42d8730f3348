#!/bin/bash
# round 2, GPU call 7: default bench (e2e with the cheaper loss), interpolate-accumulator variant on the target kernels, ncu of raster_bins / texture_fwd / chain
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
show() { python - "$1" <<'PY'
import json,sys
d=json.load(open(sys.argv[1]))
print(sys.argv[1], 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3) if 'e2e' in d else None, 'ops', {k:round(v['ms'],3) for k,v in d['ops'].items() if v['ms']>0.05})
print('  target', {k:(round(v['ms'],3), round(v['frac'],3)) for k,v in d.get('target_kernels',{}).items() if isinstance(v,dict) and 'frac' in v})
for k,v in d.get('other_configs',{}).items(): print('  ',k, round(v['ms_per_step'],3), {a:(round(b['ms'],3), round(b['frac'] or 0,3)) for a,b in v['ops'].items()}, round(v['chain_roofline']['frac'],3))
PY
}
timeout -k 5 300 python bench.py --no-cpu-baseline > gpurun_out/b7_base.json 2> gpurun_out/b7_base.err; show gpurun_out/b7_base.json
B200DR_LIB=$PWD/nvdiffrast_b200/variants/libb200dr_ipg.so timeout -k 5 300 python bench.py --no-cpu-baseline --no-e2e > gpurun_out/b7_ipg.json 2> gpurun_out/b7_ipg.err; show gpurun_out/b7_ipg.json
mkdir -p gpurun_out/ncu
for k in raster_bins_kernel texture_fwd_fast texture_chain_grad_fast; do
  timeout -k 5 200 ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -o gpurun_out/ncu/$k \
      python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-extras --min-seconds 0 --batch 4 > gpurun_out/ncu/$k.log 2>&1
done
ls -la gpurun_out/ncu | tail -5
