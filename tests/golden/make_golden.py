"""Generate the golden vectors that pin the oracle: outputs of the UNMODIFIED reference (oracle/_ref, built by
oracle/build_ref.sh from /root/reference) on seeded inputs, run on a B200.

    gpurun -- python tests/golden/make_golden.py gpurun_out/golden [cube]   # on the GPU box ("cube": only texcube_*.npz)
    cp gpurun_out/golden/*.npz tests/golden/                               # here; small fixtures, committed
    gpurun -- python tests/golden/make_golden.py gpurun_out/golden noca     # second pass, separate process (see below)

Second pass ("noca"): the four gradients the reference accumulates through its warp-coalesced atomics are corrupted on
sm_100a in the unmodified build (tests/golden_check.py), so they are re-computed from the stored inputs with the
plain-atomics build of the same sources (oracle/_ref_noca, oracle/build_ref.sh) and ADDED to the existing files as
`g_pos_rf`, `ip_g_attr_rf`, `aa_g_pos_rf` (raster_*.npz) and `g_tex_rf` (texture_*.npz, texcube_*.npz); the forward
outputs of that build are asserted bitwise equal to the stored ones first (only the accumulation differs).

Only reference OUTPUTS (and the seeded inputs) are stored, no reference code.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import workloads as wl          # noqa: E402
from util import load_reference  # noqa: E402


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def npy(t):
    return t.detach().cpu().numpy()


def rnd(shape, seed):
    return np.random.default_rng(seed).normal(size=shape).astype(np.float32)


def texture_golden(ref, path, tex, uv, uv_da, bias, fm, bm):
    ttex = cu(tex).requires_grad_(True); tuv = cu(uv).requires_grad_(True)
    tda = cu(uv_da).requires_grad_(True) if uv_da is not None else None
    tb = cu(bias).requires_grad_(True) if bias is not None else None
    out = ref.texture(ttex, tuv, tda, tb, filter_mode=fm, boundary_mode=bm)
    dy = rnd(out.shape, 17)
    (out * cu(dy)).sum().backward()
    d = dict(kind="texture", filter_mode=fm, boundary_mode=bm, tex=tex, uv=uv, dy=dy, out=npy(out), g_tex=npy(ttex.grad))
    if uv_da is not None: d.update(uv_da=uv_da)
    if bias is not None: d.update(bias=bias)
    if tuv.grad is not None: d.update(g_uv=npy(tuv.grad))
    if tda is not None and tda.grad is not None: d.update(g_uv_da=npy(tda.grad))
    if tb is not None and tb.grad is not None: d.update(g_bias=npy(tb.grad))
    np.savez_compressed(path, **d)


def cube_goldens(ref, outdir):
    """Cube-map texture cases (workloads.CUBE_CASES) -> texcube_<i>.npz"""
    for i in range(len(wl.CUBE_CASES)):
        c = wl.cube_texture_case(i)
        texture_golden(ref, os.path.join(outdir, "texcube_%d.npz" % i), c["tex"], c["uv"], c["uv_da"], c["bias"], c["filter_mode"], "cube")


def noca_pass(outdir):
    """Add the race-free reference gradients (`*_rf`) to every golden file of tests/golden/; results go to outdir."""
    ref = load_reference("noca")
    assert ref is not None, "oracle/_ref_noca missing"
    os.makedirs(outdir, exist_ok=True)
    ctx = ref.RasterizeCudaContext()
    src = os.path.dirname(os.path.abspath(__file__))
    for fn in sorted(os.listdir(src)):
        if not fn.endswith(".npz"):
            continue
        g = dict(np.load(os.path.join(src, fn)))
        kind = str(g["kind"])
        if kind == "raster":
            pos, tri, res = g["pos"], g["tri"], tuple(int(v) for v in g["res"])
            tpos = cu(pos).requires_grad_(True)
            rast, db = ref.rasterize(ctx, tpos, cu(tri), res)
            assert np.array_equal(npy(rast), g["rast"]) and np.array_equal(npy(db), g["rast_db"]), fn
            ((rast * cu(g["dy"])).sum() + (db * cu(g["ddb"])).sum()).backward()
            g["g_pos_rf"] = npy(tpos.grad)
            tattr = cu(g["attr"]).requires_grad_(True)
            out, out_da = ref.interpolate(tattr, rast.detach(), cu(tri), rast_db=db.detach(), diff_attrs="all")
            assert np.array_equal(npy(out), g["ip_out"]), fn
            ((out * cu(g["ip_dy"])).sum() + (out_da * cu(g["ip_dda"])).sum()).backward()
            g["ip_g_attr_rf"] = npy(tattr.grad)
            tcol = cu(g["color"]).requires_grad_(True); tpos2 = cu(pos).requires_grad_(True)
            aa = ref.antialias(tcol, rast.detach(), tpos2, cu(tri))
            (aa * cu(g["aa_dy"])).sum().backward()
            g["aa_g_pos_rf"] = npy(tpos2.grad)
            g["aa_g_color_rf"] = npy(tcol.grad)
        elif kind == "texture":
            ttex = cu(g["tex"]).requires_grad_(True)
            tda = cu(g["uv_da"]) if "uv_da" in g else None
            tb = cu(g["bias"]) if "bias" in g else None
            out = ref.texture(ttex, cu(g["uv"]), tda, tb, filter_mode=str(g["filter_mode"]), boundary_mode=str(g["boundary_mode"]))
            assert np.array_equal(npy(out), g["out"]), fn
            (out * cu(g["dy"])).sum().backward()
            g["g_tex_rf"] = npy(ttex.grad)
        else:
            continue
        np.savez_compressed(os.path.join(outdir, fn), **g)
        print("noca pass:", fn)


def main(outdir, only=None):
    if only == "noca":
        return noca_pass(outdir)
    ref = load_reference()
    assert ref is not None, "oracle/_ref missing"
    os.makedirs(outdir, exist_ok=True)
    if only == "cube":
        cube_goldens(ref, outdir)
        print("cube golden vectors written to", outdir)
        return
    ctx = ref.RasterizeCudaContext()
    scenes = {
        "triangle": (wl.triangle_scene()[0], wl.triangle_scene()[1], (64, 64)),
        "cube": (wl.config_c2(n=2, res=64)["pos"], wl.cube_mesh()[1], (64, 64)),
        "cube_odd": (wl.config_c2(n=1, res=64)["pos"], wl.cube_mesh()[1], (45, 77)),
        "sphere": (wl.config_c3(n=1, res=96, tex_res=8)["pos"], wl.sphere_mesh()[1], (96, 96)),
        "soup": wl.soup_scene(2, 300, 1) + ((96, 96),),
        "grid": wl.grid_aligned_scene() + ((32, 32),),
        "zfight": wl.zfight_scene() + ((64, 64),),
    }
    for name, (pos, tri, res) in scenes.items():
        d = dict(kind="raster", pos=pos, tri=tri, res=np.array(res))
        tpos = cu(pos).requires_grad_(True)
        rast, db = ref.rasterize(ctx, tpos, cu(tri), res)
        dy, ddb = rnd(rast.shape, 5), rnd(rast.shape, 6)
        ((rast * cu(dy)).sum() + (db * cu(ddb)).sum()).backward()
        d.update(rast=npy(rast), rast_db=npy(db), dy=dy, ddb=ddb, g_pos=npy(tpos.grad))
        # interpolate with pixel differentials of all attributes
        V = pos.shape[1]
        attr = rnd((pos.shape[0], V, 3), 7)
        tattr = cu(attr).requires_grad_(True); trast = rast.detach().clone().requires_grad_(True); tdb = db.detach().clone().requires_grad_(True)
        out, out_da = ref.interpolate(tattr, trast, cu(tri), rast_db=tdb, diff_attrs="all")
        idy, idda = rnd(out.shape, 8), rnd(out_da.shape, 9)
        ((out * cu(idy)).sum() + (out_da * cu(idda)).sum()).backward()
        d.update(attr=attr, ip_out=npy(out), ip_out_da=npy(out_da), ip_dy=idy, ip_dda=idda,
                 ip_g_attr=npy(tattr.grad), ip_g_rast=npy(trast.grad), ip_g_rast_db=npy(tdb.grad))
        # antialias
        color = np.random.default_rng(14).uniform(size=rast.shape[:3] + (3,)).astype(np.float32)
        tcol = cu(color).requires_grad_(True); tpos2 = cu(pos).requires_grad_(True)
        aa = ref.antialias(tcol, rast.detach(), tpos2, cu(tri))
        ady = rnd(aa.shape, 15)
        (aa * cu(ady)).sum().backward()
        d.update(color=color, aa_out=npy(aa), aa_dy=ady, aa_g_color=npy(tcol.grad), aa_g_pos=npy(tpos2.grad))
        # depth peeling, three layers (ids only)
        with ref.DepthPeeler(ctx, cu(pos), cu(tri), res) as peeler:
            layers = [npy(peeler.rasterize_next_layer()[0])[..., 3].astype(np.int32) for _ in range(3)]
        d.update(peel_ids=np.stack(layers))
        np.savez_compressed(os.path.join(outdir, "raster_%s.npz" % name), **d)

    # range mode
    pos, tri = wl.soup_scene(1, 300, 1)
    ranges = np.array([[0, 100], [100, 150], [37, 1], [250, 50]], np.int32)
    rast, _ = ref.rasterize(ctx, cu(pos[0]), cu(tri), (64, 64), ranges=torch.from_numpy(ranges))
    np.savez_compressed(os.path.join(outdir, "range_soup.npz"), kind="range", pos=pos[0], tri=tri, ranges=ranges, res=np.array((64, 64)), rast=npy(rast))

    # texture
    rng = np.random.default_rng(16)
    N, H, W = 2, 40, 56
    cases = [("nearest", "wrap", 0, 0), ("linear", "wrap", 0, 0), ("linear", "clamp", 0, 0), ("linear", "zero", 0, 0),
             ("linear-mipmap-nearest", "wrap", 1, 0), ("linear-mipmap-linear", "wrap", 1, 0),
             ("linear-mipmap-linear", "clamp", 1, 1), ("linear-mipmap-linear", "zero", 0, 1)]
    for i, (fm, bm, use_da, use_bias) in enumerate(cases):
        C = (1, 3, 4)[i % 3]
        tex = rng.uniform(size=(1 if C != 3 else 2, 32, 64, C)).astype(np.float32)
        uv = rng.uniform(-0.3, 1.3, size=(N, H, W, 2)).astype(np.float32)
        uv_da = (rng.normal(size=(N, H, W, 4)) * 0.03).astype(np.float32) if use_da else None
        bias = rng.uniform(-1, 3, size=(N, H, W)).astype(np.float32) if use_bias else None
        ttex = cu(tex).requires_grad_(True); tuv = cu(uv).requires_grad_(True)
        tda = cu(uv_da).requires_grad_(True) if use_da else None
        tb = cu(bias).requires_grad_(True) if use_bias else None
        out = ref.texture(ttex, tuv, tda, tb, filter_mode=fm, boundary_mode=bm)
        dy = rnd(out.shape, 17)
        (out * cu(dy)).sum().backward()
        d = dict(kind="texture", filter_mode=fm, boundary_mode=bm, tex=tex, uv=uv, dy=dy, out=npy(out), g_tex=npy(ttex.grad))
        if use_da: d.update(uv_da=uv_da)
        if use_bias: d.update(bias=bias)
        if tuv.grad is not None: d.update(g_uv=npy(tuv.grad))
        if use_da and tda.grad is not None: d.update(g_uv_da=npy(tda.grad))
        if use_bias and tb.grad is not None: d.update(g_bias=npy(tb.grad))
        np.savez_compressed(os.path.join(outdir, "texture_%d.npz" % i), **d)
    cube_goldens(ref, outdir)
    print("golden vectors written to", outdir)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "golden"), sys.argv[2] if len(sys.argv) > 2 else None)
