#!/bin/bash
# Round-end evidence on one B200: parity suite, smoke, per-kernel ncu captures + launch list (first, so that the bench line can carry
# this build's measured DRAM traffic), bench lines of both arms.
#   gpurun --timeout 2400 -- 'bash scripts/final_run.sh'     then here:  python scripts/ncu_summary.py r02 ; cp gpurun_out/final/* profiles/ (see profiles/README.md)
mkdir -p gpurun_out/final
timeout -k 5 600 python -m pytest tests -m gpu -q --timeout 200 --timeout-method thread 2>&1 | tail -3 | tee gpurun_out/final/pytest_gpu.txt
grep -q passed gpurun_out/final/pytest_gpu.txt && ! grep -q failed gpurun_out/final/pytest_gpu.txt || exit 1
timeout -k 5 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2 | tee gpurun_out/final/smoke.txt
bash scripts/ncu_all.sh gpurun_out/ncu
python scripts/ncu_summary.py r02 > /dev/null 2>&1          # writes profiles/r02_ncu_kernels.md + profiles/r02_traffic.json on the box: bench.py reads the latter
mkdir -p gpurun_out/final/profiles; cp profiles/r02_ncu_kernels.md profiles/r02_traffic.json profiles/r02_launches_c3.csv gpurun_out/final/profiles/ 2>/dev/null
timeout -k 5 400 python bench.py --steps 20 --warmup 5 > gpurun_out/final/bench_ours.json 2> gpurun_out/final/bench_ours.err
timeout -k 5 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/final/bench_reference.json 2> gpurun_out/final/bench_reference.err
python - <<'PY'
import json
for n in ("ours", "reference"):
    try:
        d = json.load(open("gpurun_out/final/bench_%s.json" % n))
        print(n, d.get("ms_per_step"), d.get("value"), "e2e", (d.get("e2e") or {}).get("value"), "roofline", (d.get("roofline") or {}), "clocks", d.get("clocks"))
    except Exception as e:
        print(n, "failed", e)
PY
