#!/bin/bash
# round 2, GPU call 1: race-free reference gradients (golden second pass + full-size parity), then the three unmeasured experiments.
mkdir -p gpurun_out/golden
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
echo "== noca golden pass"; timeout 300 python tests/golden/make_golden.py gpurun_out/golden noca 2>&1 | tail -3
echo "== full-size gradient parity vs oracle/_ref_noca"; timeout 900 python -m pytest tests/test_grad_reference_gpu.py -q 2>&1 | tail -40
echo "== variants"; bash scripts/round2_first_call.sh
