"""Triangle-count sweep of the two rasterizer schedules (BINS: CTA per 64x64 bin over chunked bounding boxes; SCATTER: fused
setup + 64-bit atomicMin key buffer) at 2048^2: where should b200dr_rasterize_pick_path switch?

    gpurun -- python scripts/tri_sweep.py [n_images]      -> table on stdout + gpurun_out/tri_sweep.json

Meshes: the workloads.py lat-long sphere at increasing tessellation (12 = the c2 cube), same cameras as c3.  Times are
CUDA-event averages of rasterize forward alone (ids asserted identical between the two schedules)."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import workloads as wl                     # noqa: E402
import nvdiffrast_b200.torch as dr         # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 16       # enough pixels per launch that the host-side call overhead (~0.1 ms) is hidden
RES = 2048
MESHES = [("cube", None), (24, 22), (48, 62), (78, 78), (124, 122), (224, 224), (400, 376), (1000, 501)]


def mesh(spec):
    if spec[0] == "cube":
        v, t = wl.cube_mesh()
        return v, t, 50.0
    v, t, _, _ = wl.sphere_mesh(spec[0], spec[1])
    return v, t, 200.0


def main():
    dev = torch.device("cuda", 0)
    ctx = dr.RasterizeCudaContext(dev)
    rows = []
    for spec in MESHES:
        v, t, far = mesh(spec)
        pos = torch.from_numpy(wl.transform(wl.camera_batch(N, 0, far=far), v)).to(dev)
        tri = torch.from_numpy(t).to(dev)
        res = {}
        ids = {}
        for path in (1, 2):
            ctx.raster_path = path
            try:
                for _ in range(3):
                    r, _ = dr.rasterize(ctx, pos, tri, (RES, RES))
                torch.cuda.synchronize()
                e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
                from nvdiffrast_b200 import _lib
                trace = []
                _lib.set_trace(trace)                              # CUDA events around the C-ABI call itself: kernel time, not host overhead
                it = 0; t0 = time.perf_counter()
                while it < 5 or (time.perf_counter() - t0 < 0.3 and it < 200):
                    r, _ = dr.rasterize(ctx, pos, tri, (RES, RES)); it += 1
                    if it % 5 == 0:
                        torch.cuda.synchronize()
                torch.cuda.synchronize()
                _lib.set_trace(None)
                res[path] = sum(a.elapsed_time(b) for _, a, b in trace) / len(trace)
                ids[path] = r[..., 3].clone()
            finally:
                ctx.raster_path = 0
        assert torch.equal(ids[1], ids[2]), "schedules disagree at %s" % (spec,)
        cov = float((ids[1] > 0).float().mean())
        row = {"triangles": int(t.shape[0]), "vertices": int(v.shape[0]), "coverage": cov, "bins_ms": res[1], "scatter_ms": res[2],
               "bins_frac": 32 * N * RES * RES / (res[1] * 1e-3) / 1e9 / 6570.0, "scatter_frac": 32 * N * RES * RES / (res[2] * 1e-3) / 1e9 / 6570.0}
        rows.append(row)
        print("%8d tris  cov %.3f  bins %7.3f ms (%.2f)  scatter %7.3f ms (%.2f)  %s" % (row["triangles"], cov, res[1], row["bins_frac"], res[2], row["scatter_frac"],
                                                                                         "BINS" if res[1] <= res[2] else "SCATTER"), flush=True)
        del pos, tri, ids
        ctx._scratch = None
        torch.cuda.empty_cache()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump({"n_images": N, "res": RES, "rows": rows}, open(os.path.join(ROOT, "gpurun_out", "tri_sweep.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
