import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import workloads as wl, nvdiffrast_b200.torch as dr
c = wl.config_c3(n=2, res=1024, tex_res=64)
cu = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
ctx = dr.RasterizeCudaContext()
for i in range(3):
    rast, db = dr.rasterize(ctx, cu(c["pos"]), cu(c["tri"]), c["res"])
    torch.cuda.synchronize()
    print("iter", i, "coverage", float((rast[..., 3] > 0).float().mean()))
