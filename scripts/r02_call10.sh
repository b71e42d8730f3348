#!/bin/bash
# round 2, GPU call 10: GPU tests + default bench after the reduction-step / occupancy changes
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
echo "== smoke"; timeout -k 5 240 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 || { echo "SMOKE FAILED/HUNG - stopping"; exit 1; }
echo "== pytest -m gpu"; timeout -k 5 1500 python -m pytest tests -m gpu -q --timeout 200 --timeout-method thread 2>&1 | grep -E "^(FAILED|ERROR|E  .*(Error|elements outside|must stay))|passed|failed" | cut -c1-300 | head -40
echo "== bench"; timeout -k 5 400 python bench.py --no-cpu-baseline > gpurun_out/b10.json 2> gpurun_out/b10.err; tail -2 gpurun_out/b10.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/b10.json'))
print('ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3), 'ops', {k:round(v['ms'],3) for k,v in d['ops'].items() if v['ms']>0.05})
print('unfused', round(d['ops_unfused']['ms_per_step'],3), {k:(round(v['ms'],3), round(v['frac'] or 0,3)) for k,v in d['ops_unfused']['ops'].items() if v['ms']>0.05})
print('target', {k:(round(v['ms'],3), round(v['frac'],3)) for k,v in d['target_kernels'].items() if isinstance(v,dict) and 'frac' in v})
for k,v in d['other_configs'].items(): print('  ',k, round(v.get('ms_per_step',0),3), {a:(round(b['ms'],3), round(b['frac'] or 0,3)) for a,b in v['ops'].items()})
PY
