#!/bin/bash
# round 2, GPU call 19 (the last 15 GPU-minutes of the round): parity suite of the current build incl. the new tests (inverse-rendering
# loops, CUDA-graph capture), A/B of the fused chain backward's STRIPS schedule / rast_db requests in one process, parity of the winner,
# bench line of the winner, ncu of the winner's fused kernel.  Everything important is written to gpurun_out/call19/ as it is produced.
O=gpurun_out/call19; mkdir -p $O
V=$PWD/nvdiffrast_b200/variants
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader | tee $O/gpu.txt
timeout -k 5 150 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee $O/smoke.txt
echo "== pytest -m gpu (default build)"; date +%s
timeout -k 5 420 python -m pytest tests -m gpu -q -rf --timeout 200 --timeout-method thread 2>&1 | tail -15 | tee $O/pytest_default.txt
echo "== A/B"; date +%s
timeout -k 5 150 python scripts/ab_chain.py --out $O/ab.json base strips8 strips4 strips16 strips32 strips8w4 s8pfdb s8earlydb s8both pfdb earlydb strips8b6 2>&1 | grep -v Warning | tee $O/ab.txt
BEST=$(python -c "import json; print(json.load(open('$O/ab.json'))['best'])" 2>/dev/null || echo base)
echo "best variant: $BEST" | tee $O/best.txt
if [ "$BEST" != "base" ]; then
  export B200DR_LIB=$V/libb200dr_$BEST.so
  echo "== pytest (winner: $BEST): everything that runs the fused chain backward"; date +%s
  timeout -k 5 300 python -m pytest tests/test_fusion_gpu.py tests/test_fullsize_gpu.py tests/test_grad_reference_gpu.py tests/test_fit_gpu.py tests/test_graph_gpu.py tests/test_ops_gpu.py -m gpu -q --timeout 200 --timeout-method thread -k "chain or fus or fit or graph or full or reference or texture" 2>&1 | tail -8 | tee $O/pytest_winner.txt
fi
echo "== bench line ($BEST)"; date +%s
timeout -k 5 330 python bench.py --steps 20 --warmup 5 > $O/bench_ours.json 2> $O/bench_ours.err
python - <<PY | tee $O/bench_summary.txt
import json
try:
    d = json.load(open("$O/bench_ours.json"))
    print("ours", round(d["ms_per_step"], 4), "ms", round(d["value"], 1), "Mpx/s e2e", round(d["e2e"]["value"], 1), "chain", d["ops"].get("texture_chain_grad", {}).get("ms"), "clk", d["clocks"])
    oc = d.get("other_configs", {})
    print({k: (round(v.get("ms_per_step", 0), 4) if "ms_per_step" in v else v) for k, v in oc.items() if k in ("c2", "c4", "c5")})
    print("graph leg", oc.get("cuda_graph_launch_bound"))
except Exception as e:
    print("bench failed", e)
PY
echo "== ncu of the fused kernel ($BEST)"; date +%s
timeout -k 5 150 ncu --set full --clock-control none --import-source on -k regex:texture_chain_grad_fast -s 3 -c 1 -o $O/texture_chain_grad_fast \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-extras --min-seconds 0 --batch 4 > $O/ncu_chain.log 2>&1
ls -la $O | head -30
unset B200DR_LIB
echo "== reference arm"; date +%s
timeout -k 5 200 python bench.py --impl reference --steps 20 --warmup 5 > $O/bench_reference.json 2> $O/bench_reference.err
python -c "
import json; d=json.load(open('$O/bench_reference.json')); print('reference', d.get('ms_per_step'), d.get('value'), (d.get('e2e') or {}).get('value'))" 2>&1 | tee -a $O/bench_summary.txt
date +%s
