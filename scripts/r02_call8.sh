#!/bin/bash
# round 2, GPU call 8: BASELINE.md table (reference vs ours), c4 launch list, interpolate unroll variants on the target kernels
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
echo "== baseline table"; timeout -k 5 400 python scripts/baseline_table.py 2>&1 | tail -3 | cut -c1-1500
echo "== c4 launch list"
timeout -k 5 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_c4.csv \
    python bench.py --workload c4 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-extras --min-seconds 0 > gpurun_out/launches_c4.log 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/launches_c4.csv')) if len(r)>10]; hdr=rows[0]
ki,vi=hdr.index('Kernel Name'),hdr.index('Metric Value')
agg=collections.OrderedDict()
for r in rows[1:]:
    try: v=float(r[vi].replace(',',''))
    except ValueError: continue
    k=r[ki][:60]; agg.setdefault(k,[0.0,0]); agg[k][0]+=v; agg[k][1]+=1
for k,(t,c) in sorted(agg.items(), key=lambda x:-x[1][0])[:12]: print('%-62s %3d launches %9.1f us each'%(k,c,t/c/1e3))
PY
echo "== interpolate unroll variants (target kernels)"
for v in ipf4 ipf1; do
B200DR_LIB=$PWD/nvdiffrast_b200/variants/libb200dr_$v.so timeout -k 5 300 python - <<PY
import json, torch, bench
mod,_=bench.load_module("ours")
t=bench.target_kernels(mod, torch.device("cuda",0), 6570.0)
print("$v", {k:(round(x["ms"],3), round(x["frac"],3)) for k,x in t.items() if isinstance(x,dict) and "frac" in x})
PY
done
echo "== cube tests"; timeout -k 5 300 python -m pytest tests/test_ops_gpu.py -q -k cube --timeout 120 --timeout-method thread 2>&1 | tail -3
echo "== envphong-shaped"; timeout -k 5 200 python - <<'PY'
import torch, bench, json
mod,_=bench.load_module("ours")
print(json.dumps(bench.envphong_shaped(mod, torch.device("cuda",0), 6570.0)))
PY
