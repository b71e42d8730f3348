"""Numbers for BASELINE.md section 2: the UNMODIFIED reference (oracle/_ref) and this implementation on the same B200, same
inputs, same timing: c2 / c3 / c4 chains through bench.py's Pipeline (ms per step, CUDA events, >= 0.5 s per point) and the
four target kernels (interpolate fwd / bwd at A = 3, antialias fwd / bwd at C = 3, c3 geometry, 2048^2 x 16 images) timed
as op calls through the public API: forward = the op call, backward = out.backward(dy), each bracketed by CUDA events.

    gpurun -- python scripts/baseline_table.py        -> gpurun_out/baseline_table.json + a table on stdout
"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench                      # noqa: E402
import workloads as wl            # noqa: E402

dev = torch.device("cuda", 0)


def chain_ms(mod, name):
    res, n = bench.DEFAULTS[name]
    pipe = bench.Pipeline(mod, name, n, res, dev, seed=0)
    for _ in range(3):
        pipe.step()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    it, t0 = 0, time.perf_counter()
    e0.record()
    while it < 10 or time.perf_counter() - t0 < 0.5:
        pipe.step(); it += 1
        if it % 10 == 0:
            torch.cuda.synchronize()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / it
    del pipe
    torch.cuda.empty_cache()
    return {"ms_per_step": ms, "Mpx_s": n * res * res / (ms * 1e-3) / 1e6}


def targets(mod, n_img=16, res=2048):
    c = wl.config_c3(n=1, res=res, tex_res=8)
    v = wl.sphere_mesh()[0]
    verts = torch.from_numpy(np.concatenate([v, np.ones((v.shape[0], 1), np.float32)], 1)).to(dev)
    mvp = torch.from_numpy(wl.camera_batch(n_img, 0, far=200.0)).to(dev)
    pos = torch.matmul(verts[None], mvp.transpose(1, 2)).contiguous()
    tri = torch.from_numpy(c["tri"]).to(dev)
    ctx = mod.RasterizeCudaContext(dev)
    with torch.no_grad():
        rast, _ = mod.rasterize(ctx, pos, tri, (res, res))
    g = torch.Generator(device=dev); g.manual_seed(7)
    attr = torch.rand((1, v.shape[0], 3), device=dev, generator=g)
    dy = torch.randn((n_img, res, res, 3), device=dev, generator=g)
    topo = mod.antialias_construct_topology_hash(tri)
    P = n_img * res * res
    out = {}

    def run(fwd_fn, name, bytes_fwd, bytes_bwd):
        tf = tb = 0.0; it = 0
        for k in range(3 + 20):
            e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            e[0].record(); o = fwd_fn(); e[1].record(); o.backward(dy); e[2].record()
            torch.cuda.synchronize()
            if k >= 3:
                tf += e[0].elapsed_time(e[1]); tb += e[1].elapsed_time(e[2]); it += 1
        out[name + "_fwd"] = {"ms": tf / it, "frac_of_6570GBs": bytes_fwd * P / (tf / it * 1e-3) / 1e9 / 6570.0}
        out[name + "_bwd"] = {"ms": tb / it, "frac_of_6570GBs": bytes_bwd * P / (tb / it * 1e-3) / 1e9 / 6570.0}

    def interp():
        a = attr.detach().requires_grad_(True); r = rast.detach().requires_grad_(True)
        return mod.interpolate(a, r, tri)[0]

    def aa():
        col = dy.detach().requires_grad_(True); pp = pos.detach().requires_grad_(True)
        return mod.antialias(col, rast, pp, tri, topology_hash=topo)

    run(interp, "interpolate", 28, 44)
    run(aa, "antialias", 40, 24)
    return out


def main():
    res = {}
    for impl in ("reference", "ours"):
        mod, _ = bench.load_module(impl)
        r = {"chains": {k: chain_ms(mod, k) for k in ("c2", "c3", "c4")}, "targets": targets(mod)}
        res[impl] = r
        print(impl, json.dumps(r))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "baseline_table.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
