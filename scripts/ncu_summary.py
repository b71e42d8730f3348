"""Summarise gpurun_out/ncu/*.ncu-rep (one kernel each) + the launch list into profiles/ (committed evidence)."""
import csv, glob, json, os, subprocess, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = os.path.join(ROOT, "gpurun_out", "ncu"); dst = os.path.join(ROOT, "profiles")
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
px = 4 * 2048 * 2048
M = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
     'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
     'smsp__inst_executed.sum', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed']
opmap = {"raster_bins_kernel": "rasterize_fwd", "interpolate_fwd_fast": "interpolate_fwd", "texture_fwd_fast": "texture_fwd", "aa_discontinuity_kernel": "antialias_fwd",
         "aa_copy_kernel": "antialias_grad", "texture_chain_grad_fast": "texture_chain_grad", "texture_grad_fast": "texture_grad", "interpolate_grad_fast": "interpolate_grad", "raster_grad_kernel": "rasterize_grad"}
lines = ["# ncu --set full, one kernel per capture (c3 workload, 4 images = 16.8 Mpixel per launch, --clock-control none)\n",
         "| kernel | time us | DRAM read MB | DRAM write MB | DRAM B/px | DRAM % | L1 % | L2 % | issue active % | warps active % | regs | warp-inst / 32 px |", "|---|---|---|---|---|---|---|---|---|---|---|---|"]
traffic = {}
for rep in sorted(glob.glob(os.path.join(src, "*.ncu-rep"))):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines())); hdr = rows[0]; units = rows[1]; r = rows[2]
    def f(n):
        v = float(r[hdr.index(n)].replace(',', '')); u = units[hdr.index(n)]
        return v * {"Gbyte": 1e3, "Mbyte": 1.0, "Kbyte": 1e-3, "byte": 1e-6}.get(u, 1.0) if "byte" in u else v * {"ms": 1e3, "us": 1.0, "ns": 1e-3}.get(u, 1.0) if u in ("ms", "us", "ns") else v
    name = os.path.basename(rep)[:-8]
    bpp = (f(M[1]) + f(M[2])) * 1e6 / px
    lines.append("| %s | %.1f | %.1f | %.1f | %.1f | %.1f | %.1f | %.1f | %.1f | %.1f | %d | %.0f |" % (name, f(M[0]), f(M[1]), f(M[2]), bpp, f(M[3]), f(M[8]), f(M[9]), f(M[4]), f(M[5]), f(M[6]), f(M[7]) / (px / 32)))
    traffic[opmap.get(name, name)] = {"dram_bytes_per_pixel": bpp, "kernel": name, "source": "profiles/%s_ncu_kernels.md (ncu --set full, 16.8 Mpixel launch)" % tag}
# launch list
ll = os.path.join(src, "launches_c3.csv")
if os.path.exists(ll):
    rows = [r for r in csv.reader(open(ll)) if len(r) > 10]; hdr = rows[0]
    ki, vi = hdr.index('Kernel Name'), hdr.index('Metric Value')
    agg = collections.OrderedDict()
    for r in rows[1:]:
        try: v = float(r[vi].replace(',', ''))
        except ValueError: continue
        k = r[ki].replace('<unnamed>::', '').replace('void ', '')[:70]
        agg.setdefault(k, [0.0, 0]); agg[k][0] += v; agg[k][1] += 1
    tot = sum(v[0] for v in agg.values())
    lines += ["", "# launch list of `bench.py --steps 2 --warmup 3` (16 images, gpu__time_duration, serialised / cold cache: compare shares)\n", "| kernel | launches | total us | share % |", "|---|---|---|---|"]
    for k, (t, c) in sorted(agg.items(), key=lambda x: -x[1][0])[:24]:
        lines.append("| %s | %d | %.1f | %.1f |" % (k, c, t / 1e3, 100 * t / tot))
    import shutil; shutil.copy(ll, os.path.join(dst, "%s_launches_c3.csv" % tag))
open(os.path.join(dst, "%s_ncu_kernels.md" % tag), "w").write("\n".join(lines) + "\n")
json.dump(traffic, open(os.path.join(dst, "%s_traffic.json" % tag), "w"), indent=1)
print("\n".join(lines))
