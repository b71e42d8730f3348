#!/bin/bash
# round 2, GPU call 18: interpolate backward (A = 3 target kernel): groups in flight / occupancy / early dy requests
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
timeout -k 5 120 python -c "import __graft_entry__ as g; g.smoke()" || { echo "SMOKE FAILED"; exit 1; }
V=$PWD/nvdiffrast_b200/variants
tk() {
B200DR_LIB=$2 timeout -k 5 200 python - <<PY
import json, torch, bench
mod,_=bench.load_module("ours")
t=bench.target_kernels(mod, torch.device("cuda",0), 6570.0)
print("$1", {k:(round(x["ms"],3), round(x["frac"],3)) for k,x in t.items() if isinstance(x,dict) and "frac" in x and k.startswith("interp")})
PY
}
run() {
  B200DR_LIB=$2 B200DR_FUSION=$3 timeout -k 5 150 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-extras $4 > gpurun_out/b_$1.json 2>gpurun_out/b_$1.err
  python -c "
import json; d=json.load(open('gpurun_out/b_$1.json')); o=d['ops']; print('$1', round(d['ms_per_step'],3), 'rounds', d['timed_rounds'], 'clk', d['clocks']['sm_mhz'], d['clocks']['samples'], ' '.join('%s=%.3f'%(k.replace('rasterize','ras').replace('interpolate','ip').replace('texture','tex').replace('antialias','aa'),v['ms']) for k,v in o.items() if v['ms']>0.05))" || tail -5 gpurun_out/b_$1.err
}
tk base ""
for v in early earlyu6 u6 u8b3 b5; do tk $v $V/libb200dr_$v.so; done
tk base2 ""
echo "== early dy: parity + c2 / c4 / c3 op-by-op"
B200DR_LIB=$V/libb200dr_early.so timeout -k 10 600 python -m pytest tests/test_ops_gpu.py tests/test_fusion_gpu.py -x -q -k "interp or chain or fus" --timeout 200 --timeout-method thread 2>&1 | tail -3
run c2_base "" 1 "--workload c2"
run c2_early $V/libb200dr_early.so 1 "--workload c2"
run c4_base "" 1 "--workload c4"
run c4_early $V/libb200dr_early.so 1 "--workload c4"
