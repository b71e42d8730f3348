#!/bin/bash
# Round-end evidence on one B200: parity suite, smoke, bench lines (ours / reference / c2 / c4), per-kernel ncu captures, launch list.
mkdir -p gpurun_out/final
timeout -k 5 400 python -m pytest tests -m gpu -q 2>&1 | tail -3 | tee gpurun_out/final/pytest_gpu.txt
grep -q passed gpurun_out/final/pytest_gpu.txt && ! grep -q failed gpurun_out/final/pytest_gpu.txt || exit 1
timeout -k 5 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2 | tee gpurun_out/final/smoke.txt
timeout -k 5 300 python bench.py --steps 20 --warmup 5 > gpurun_out/final/bench_ours.json 2> gpurun_out/final/bench_ours.err
timeout -k 5 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/final/bench_reference.json 2> gpurun_out/final/bench_reference.err
timeout -k 5 200 python bench.py --workload c2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/final/bench_c2.json 2> gpurun_out/final/bench_c2.err
timeout -k 5 200 python bench.py --workload c4 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/final/bench_c4.json 2> gpurun_out/final/bench_c4.err
python - <<'PY'
import json
for n in ("ours", "reference", "c2", "c4"):
    try:
        d = json.load(open("gpurun_out/final/bench_%s.json" % n))
        print(n, d.get("ms_per_step"), d.get("value"), "e2e", (d.get("e2e") or {}).get("value"), "roofline", (d.get("roofline") or {}).get("frac"))
    except Exception as e:
        print(n, "failed", e)
PY
bash scripts/ncu_all.sh
