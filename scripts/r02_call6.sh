#!/bin/bash
# round 2, GPU call 6: cube goldens for the new channel counts (+ race-free pass), full default bench line (ours + reference arm)
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
mkdir -p gpurun_out/golden
echo "== cube goldens"; timeout 300 python tests/golden/make_golden.py gpurun_out/golden cube 2>&1 | tail -2
cp gpurun_out/golden/texcube_7.npz gpurun_out/golden/texcube_8.npz tests/golden/ 
timeout 300 python tests/golden/make_golden.py gpurun_out/golden noca 2>&1 | tail -3
echo "== cube tests"; timeout 300 python -m pytest tests/test_ops_gpu.py -q -k cube 2>&1 | tail -3
echo "== bench ours (default)"; ( time timeout -k 5 600 python bench.py > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err ) 2>&1 | grep real; tail -3 gpurun_out/bench_ours.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_ours.json'))
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d.get('e2e'), 'clocks', d['clocks'])
print('roofline', d['roofline'])
print('chain', d['chain_roofline'])
print('ops', {k:(round(v['ms'],3), round(v.get('frac',0),3)) for k,v in d['ops'].items()})
print('unfused', d.get('ops_unfused',{}).get('ms_per_step'), {k:(round(v['ms'],3), round(v['frac'] or 0,3)) for k,v in d.get('ops_unfused',{}).get('ops',{}).items()})
print('target', {k:v for k,v in d.get('target_kernels',{}).items()})
for k,v in d.get('other_configs',{}).items(): print(k, v['ms_per_step'], v['value'], v['clocks'], {a:(round(b['ms'],3), round(b['frac'] or 0,3)) for a,b in v['ops'].items()}, v['chain_roofline'])
print('cpu', d.get('cpu_baseline'), 'launches', d.get('gpu_launches'))
PY
echo "== bench reference"; ( time timeout -k 5 600 python bench.py --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err ) 2>&1 | grep real; tail -3 gpurun_out/bench_ref.err
python -c "
import json; d=json.load(open('gpurun_out/bench_ref.json')); print('ref value', d['value'], 'ms', d['ms_per_step'], 'e2e', d.get('e2e'), d['clocks'])"
