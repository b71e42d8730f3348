"""Summarise gpurun_out/ncu_c4/*.ncu-rep (kernels that only the c2 / c4 configurations run) into profiles/<tag>_ncu_c2c4.md."""
import csv, glob, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = os.path.join(ROOT, "gpurun_out", "ncu_c4"); tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
PX = {"interpolate_chain_c2": 16 * 1024 * 1024}           # pixels per captured launch; default: c4, 2 images of 2048^2
M = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
     'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
     'smsp__inst_executed.sum', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
     'smsp__thread_inst_executed_per_inst_executed.ratio']
lines = ["# ncu --set full --clock-control none, one kernel per capture: c4 (1 M triangles, 2 images of 2048^2 per launch) and c2 (16 images of 1024^2)\n",
         "| capture | kernel | time us | DRAM read MB | DRAM write MB | DRAM B/px | DRAM % | L1 % | L2 % | issue active % | warps active % | regs | warp-inst / 32 px | threads / inst | top stalls (per issue) |",
         "|---|---|---|---|---|---|---|---|---|---|---|---|---|---|---|"]
for rep in sorted(glob.glob(os.path.join(src, "*.ncu-rep"))):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines())); hdr, units, r = rows[0], rows[1], rows[2]
    def f(n):
        v = float(r[hdr.index(n)].replace(',', '')); u = units[hdr.index(n)]
        if "byte" in u: return v * {"Gbyte": 1e3, "Mbyte": 1.0, "Kbyte": 1e-3, "byte": 1e-6}.get(u, 1.0)
        if u in ("ms", "us", "ns"): return v * {"ms": 1e3, "us": 1.0, "ns": 1e-3}[u]
        return v
    name = os.path.basename(rep)[:-8]
    px = PX.get(name, 2 * 2048 * 2048)
    st = sorted(((float(r[i].replace(',', '') or 0), h) for i, h in enumerate(hdr) if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('_per_issue_active.ratio')), reverse=True)[:3]
    stalls = ", ".join("%s %.1f" % (h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''), v) for v, h in st)
    kern = r[hdr.index('Kernel Name')].replace('<unnamed>::', '').replace('void ', '').split('(')[0]
    lines.append("| %s | %s | %.1f | %.1f | %.1f | %.1f | %.1f | %.1f | %.1f | %.1f | %.1f | %d | %.0f | %.1f | %s |" % (
        name, kern, f(M[0]), f(M[1]), f(M[2]), (f(M[1]) + f(M[2])) * 1e6 / px, f(M[3]), f(M[8]), f(M[9]), f(M[4]), f(M[5]), f(M[6]), f(M[7]) / (px / 32), f(M[10]), stalls))
open(os.path.join(ROOT, "profiles", "%s_ncu_c2c4.md" % tag), "w").write("\n".join(lines) + "\n")
print("\n".join(lines))
