"""c4 (1M-triangle sphere, 2048^2) parity against the unmodified reference on the same GPU + timing of both."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import workloads as wl
from util import load_reference
import nvdiffrast_b200.torch as dr
ref = load_reference()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
c = wl.config_c4(n=n, res=2048)
cu = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
pos, tri = cu(c["pos"]), cu(c["tri"])
ctx = dr.RasterizeCudaContext()
def timed(f, reps=3):
    f(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): r = f()
    torch.cuda.synchronize()
    return r, (time.perf_counter() - t0) / reps * 1e3
for path in (2, 1):
    ctx.raster_path = path
    (rast, db), ms = timed(lambda: dr.rasterize(ctx, pos, tri, c["res"]))
    print("ours path %d: %.2f ms  covered %.3f" % (path, ms, float((rast[..., 3] > 0).float().mean())))
    if path == 2: rast2 = rast
    else: print("paths identical:", bool(torch.equal(rast, rast2)))
if ref is not None:
    rctx = ref.RasterizeCudaContext()
    (rr, rdb), ms = timed(lambda: ref.rasterize(rctx, pos, tri, c["res"]))
    print("reference: %.2f ms" % ms)
    neq = int((rr[..., 3] != rast2[..., 3]).sum())
    print("id mismatches vs reference: %d of %d" % (neq, rr[..., 3].numel()))
    d = (rr[..., :3] - rast2[..., :3]).abs()
    print("max |uvz| diff %.3g ; frac > 1e-5: %.3g" % (float(d.max()), float((d > 1e-5).float().mean())))
