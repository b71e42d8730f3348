#!/bin/bash
# round 2, GPU call 13: interpolate-backward occupancy variants on the target kernels, then the round-end evidence run
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
run() {
  B200DR_LIB=$2 B200DR_FUSION=$3 timeout -k 5 150 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-extras $4 > gpurun_out/b_$1.json 2>gpurun_out/b_$1.err
  python -c "
import json; d=json.load(open('gpurun_out/b_$1.json')); o=d['ops']; print('$1', round(d['ms_per_step'],3), 'rounds', d['timed_rounds'], 'clk', d['clocks']['sm_mhz'], d['clocks']['samples'], ' '.join('%s=%.3f'%(k.replace('rasterize','ras').replace('interpolate','ip').replace('texture','tex').replace('antialias','aa'),v['ms']) for k,v in o.items() if v['ms']>0.05))" || tail -5 gpurun_out/b_$1.err
}
echo "== interpolate grad variants (target kernels + c3 op-by-op + c2)"
for v in ipgb4 ipgu4b4 ipgu3b4; do
[ -f nvdiffrast_b200/variants/libb200dr_$v.so ] || continue
B200DR_LIB=$PWD/nvdiffrast_b200/variants/libb200dr_$v.so timeout -k 5 300 python - <<PY
import json, torch, bench
mod,_=bench.load_module("ours")
t=bench.target_kernels(mod, torch.device("cuda",0), 6570.0)
print("$v", {k:(round(x["ms"],3), round(x["frac"],3)) for k,x in t.items() if isinstance(x,dict) and "frac" in x})
PY
run ${v}_unfused $PWD/nvdiffrast_b200/variants/libb200dr_$v.so 0
run ${v}_c2 $PWD/nvdiffrast_b200/variants/libb200dr_$v.so 1 "--workload c2"
done
run base_c2 "" 1 "--workload c2"
echo "== final run"
bash scripts/final_run.sh
