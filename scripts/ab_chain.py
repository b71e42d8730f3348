"""A/B of library variants (scripts/build_variants.py) inside ONE process: the c3 pipeline is built once, the C-ABI library is
swapped between measurements (each variant is its own shared object), every variant is measured `--passes` times in
interleaved order so that box drift shows up as disagreement between passes.

    python scripts/ab_chain.py --out gpurun_out/ab.json base strips8 strips4 ...

Prints ms/step and the time of the fused chain backward per variant and pass; writes the table and the name of the winner
("base" unless a variant beats it by more than --margin in EVERY pass) to --out."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import bench  # noqa: E402
from nvdiffrast_b200 import _lib  # noqa: E402


def use_lib(name):
    _lib._lib = None
    if name == "base":
        os.environ.pop("B200DR_LIB", None)
    else:
        os.environ["B200DR_LIB"] = os.path.join(ROOT, "nvdiffrast_b200", "variants", "libb200dr_%s.so" % name)
    _lib.load()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("variants", nargs="+")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "ab.json"))
    ap.add_argument("--passes", type=int, default=2)
    ap.add_argument("--margin", type=float, default=0.015)
    ap.add_argument("--workload", default="c3")
    ap.add_argument("--seconds", type=float, default=0.6)
    args = ap.parse_args()
    device = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    use_lib("base")
    mod, _ = bench.load_module("ours")
    res, n_img = bench.DEFAULTS[args.workload]
    pipe = bench.Pipeline(mod, args.workload, n_img, res, device, seed=0)
    names = [v for v in args.variants if v == "base" or os.path.exists(os.path.join(ROOT, "nvdiffrast_b200", "variants", "libb200dr_%s.so" % v))]
    rows = {v: [] for v in names}
    for p in range(args.passes):
        for v in (names if p % 2 == 0 else names[::-1]):
            use_lib(v)
            r = bench.timed_run(pipe, 10, 3, 1, 0, 0, device, "ours", args.seconds)
            table, _ = bench.ops_table(r["trace"], 10 * r["rounds"], {}, 1.0)
            big = {k: round(x["ms"], 4) for k, x in table.items() if x["ms"] > 0.05}
            rows[v].append({"ms_per_step": r["ms_per_step"], "ops": big, "clk": (r["clocks"] or {}).get("sm_mhz")})
            print("pass %d %-12s %.4f ms/step  %s" % (p, v, r["ms_per_step"], " ".join("%s=%.3f" % (k.replace("texture", "tex").replace("interpolate", "ip").replace("rasterize", "ras").replace("antialias", "aa"), x) for k, x in big.items())), flush=True)
    use_lib("base")
    best = "base"
    if "base" in rows:
        cand = []
        for v in names:
            if v != "base" and all(a["ms_per_step"] < b["ms_per_step"] * (1.0 - args.margin) for a, b in zip(rows[v], rows["base"])):
                cand.append((sum(a["ms_per_step"] for a in rows[v]), v))
        if cand:
            best = min(cand)[1]
    print("best:", best)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump({"rows": rows, "best": best, "workload": args.workload}, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
