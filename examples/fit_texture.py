"""Learn a texture from renders -- the reference's samples/torch/earth.py in miniature, on seeded synthetic data, written
against the public API only (swap the import for `nvdiffrast.torch` and it runs on the reference unchanged).

    python examples/fit_texture.py [--res 256] [--steps 200] [--graph]

--graph records one optimisation step (forward, backward, Adam update) into a CUDA graph and replays it: the whole step is a
single launch, which is what a launch-bound loop at low resolution wants (nvdiffrast_b200/graphs.py; the reference's rasterizer
blocks the host every call and cannot be captured).
"""
import argparse
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import workloads as wl  # noqa: E402  (seeded synthetic meshes / cameras)
import nvdiffrast_b200.torch as dr  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--res", type=int, default=256)
    ap.add_argument("--tex", type=int, default=128)
    ap.add_argument("--views", type=int, default=4)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--graph", action="store_true")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)

    v, tri, uv, uv_idx = wl.sphere_mesh(48, 32)
    verts = t(np.concatenate([v, np.ones((v.shape[0], 1), np.float32)], 1))
    tri, uv, uv_idx = t(tri), t(uv), t(uv_idx)
    y, x = np.meshgrid(np.arange(args.tex) / args.tex, np.arange(args.tex) / args.tex, indexing="ij")
    target_tex = t(np.stack([0.5 + 0.5 * np.sin(12.6 * x), 0.5 + 0.5 * np.cos(12.6 * y), (np.floor(8 * x) + np.floor(8 * y)) % 2], -1)[None].astype(np.float32))
    mvp = t(wl.camera_batch(args.views, seed=1, far=200.0))                  # static buffer: new cameras are copied into it

    ctx = dr.RasterizeCudaContext(dev)
    topo = dr.antialias_construct_topology_hash(tri)

    def render(tex):
        pos = torch.matmul(verts[None], mvp.transpose(1, 2)).contiguous()
        rast, db = dr.rasterize(ctx, pos, tri, (args.res, args.res))
        texc, texd = dr.interpolate(uv[None], rast, uv_idx, rast_db=db, diff_attrs="all")
        color = dr.texture(tex, texc, texd, filter_mode="linear-mipmap-linear", max_mip_level=4)
        return dr.antialias(color, rast, pos, tri, topology_hash=topo)

    tex = torch.full_like(target_tex, 0.5).requires_grad_(True)
    opt = torch.optim.Adam([tex], lr=0.03, capturable=args.graph)
    loss_buf = torch.zeros((), device=dev)

    def step():
        with torch.no_grad():
            target = render(target_tex)
        opt.zero_grad(set_to_none=True)
        loss = torch.mean((render(tex) - target) ** 2)
        loss.backward()
        opt.step()
        loss_buf.copy_(loss.detach())
        return loss_buf

    replay = step
    if args.graph:
        from nvdiffrast_b200 import graphs
        replay = graphs.capture(step).replay
    rng = np.random.default_rng(2)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for it in range(args.steps):
        mvp.copy_(t(wl.camera_batch(args.views, seed=int(rng.integers(1 << 30)), far=200.0)))
        replay()
        if it % 50 == 0 or it == args.steps - 1:
            print("step %4d  loss %.6f" % (it, float(loss_buf)))
    torch.cuda.synchronize()
    print("%.2f ms per step (%s), texture error %.4f" % ((time.perf_counter() - t0) / args.steps * 1e3, "CUDA graph" if args.graph else "eager",
                                                        float((tex.detach() - target_tex).abs().mean())))


if __name__ == "__main__":
    main()
