#!/bin/bash
# round 2, GPU call 20 (the remaining ~7 GPU-minutes): STRIPS schedule with the next row's rast values loaded into registers one row
# ahead (TCG_PIPE_RAST), 16-row strips, 4 CTAs/SM; parity of the winner (or of the rebuilt default when nothing wins).
O=gpurun_out/call20; mkdir -p $O
V=$PWD/nvdiffrast_b200/variants
timeout -k 5 200 python scripts/ab_chain.py --out $O/ab.json base pipe8 pipe16 rows16 pipe8b4 b4 2>&1 | grep -v Warning | tee $O/ab.txt
BEST=$(python -c "import json; print(json.load(open('$O/ab.json'))['best'])" 2>/dev/null || echo base)
echo "best variant: $BEST" | tee $O/best.txt
if [ "$BEST" != "base" ]; then export B200DR_LIB=$V/libb200dr_$BEST.so; fi
timeout -k 5 240 python -m pytest tests/test_fusion_gpu.py tests/test_fullsize_gpu.py tests/test_grad_reference_gpu.py tests/test_fit_gpu.py tests/test_graph_gpu.py tests/test_ops_gpu.py -m gpu -q --timeout 200 --timeout-method thread -k "chain or fus or fit or graph or full or reference or texture" 2>&1 | tail -4 | tee $O/pytest_winner.txt
timeout -k 5 120 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline > $O/bench_ours.json 2> $O/bench_ours.err
python -c "
import json; d=json.load(open('$O/bench_ours.json')); print('ours', d['ms_per_step'], d['value'], d['e2e']['value'], d['ops']['texture_chain_grad']['ms'], d['clocks'])" 2>&1 | tee $O/bench_summary.txt
