"""Child process of tests/test_grad_reference_gpu.py: full-size gradient parity of libb200dr against the RACE-FREE
reference build (oracle/_ref_noca: the reference sources with common.h's own plain-atomicAdd branch selected, see
oracle/build_ref.sh).  A separate process because the unmodified build (oracle/_ref, used by the other GPU tests for the
forward parity) defines the same pybind module and C++ types, and only one of the two can live in a process.

    python tests/ref_noca_compare.py <cfg> <n_images> <res> [tex_res]        -> one JSON line on stdout

cfg: c2 (cube, rasterize+interpolate), c3 (sphere, rasterize+interpolate(da)+texture+antialias), c4 (1M triangles,
rasterize+interpolate).  For every op the two implementations are fed IDENTICAL input tensors (the reference's own
intermediate results), so that a difference is the op's own and not drift of an upstream knife edge; then the whole chain
is differentiated end to end in both.  Reported per tensor: max |ours - ref|, |ref|_inf, and for per-pixel tensors the
fraction of elements outside abs max(1e-5, 2^-22 |ref|_inf) + rel 1e-4 (sums of O(|ref|_inf) terms that cancel carry that
much rounding in either implementation) together with the fraction removed by the knife-edge predicates.
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import workloads as wl           # noqa: E402
from util import load_reference  # noqa: E402


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def randn(shape, seed):
    g = torch.Generator(device="cuda"); g.manual_seed(seed)
    return torch.randn(tuple(shape), device="cuda", generator=g)


def leaf(t):
    return t.detach().clone().requires_grad_(True)


def entry(name, ours, ref, kind="sum", exclude=None):
    """kind 'sum': accumulated gradient, judged by |d|_inf / |ref|_inf.  kind 'pixel': per-pixel tensor, judged element-wise
    (abs 1e-5 + rel 1e-4); `exclude` = bool mask of pixels removed by a stated knife-edge predicate."""
    o = ours.detach().double(); r = ref.detach().double()
    d = (o - r).abs()
    e = {"name": name, "kind": kind, "scale": float(r.abs().max()), "numel": int(r.numel())}
    if kind == "sum":
        e["err"] = float(d.max()); e["rel"] = e["err"] / max(e["scale"], 1e-30)
    else:
        bad = d > (max(1e-5, 2.0 ** -22 * e["scale"]) + 1e-4 * r.abs())      # abs part: 1e-5, or 4 ulp of the tensor's magnitude
        if exclude is not None:
            ex = exclude
            while ex.dim() < bad.dim():
                ex = ex[..., None]
            e["excluded_frac"] = float(exclude.double().mean())
            bad = bad & ~ex
            d = d * (~ex)
        e["err"] = float(d.max()); e["rel"] = e["err"] / max(e["scale"], 1e-30)
        e["bad_frac"] = float(bad.double().mean())
    return e


def main():
    cfg, n, res = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
    tex_res = int(sys.argv[4]) if len(sys.argv) > 4 else 1024
    ref = load_reference("noca")
    assert ref is not None, "oracle/_ref_noca missing"
    import nvdiffrast_b200.torch as dr
    if cfg == "c2":
        c = wl.config_c2(n=n, res=res)
    elif cfg == "c3":
        c = wl.config_c3(n=n, res=res, tex_res=tex_res)
    else:
        c = wl.config_c4(n=n, res=res)
    with_tex = cfg == "c3"
    tri = cu(c["tri"])
    uv_idx = cu(c["uv_idx"]) if with_tex else None
    mods = {"ours": (dr, dr.RasterizeCudaContext()), "ref": (ref, ref.RasterizeCudaContext())}
    maxmip = int(np.log2(tex_res)) - 1
    report = []

    # ---- reference forward: the intermediates every per-op comparison starts from
    with torch.no_grad():
        r_rast, r_db = ref.rasterize(mods["ref"][1], cu(c["pos"]), tri, c["res"])
        o_rast, o_db = dr.rasterize(mods["ours"][1], cu(c["pos"]), tri, c["res"])
        assert torch.equal(r_rast[..., 3], o_rast[..., 3]), "triangle ids differ"
        report.append(entry("fwd rast uvz", o_rast[..., :3], r_rast[..., :3], "pixel"))
        del o_rast, o_db
        if with_tex:
            r_uv, r_uvda = ref.interpolate(cu(c["uv"])[None], r_rast, uv_idx, rast_db=r_db, diff_attrs="all")
            r_color = ref.texture(cu(c["tex"]), r_uv, r_uvda, filter_mode="linear-mipmap-linear", max_mip_level=maxmip)
        else:
            r_color, _ = ref.interpolate(cu(c["attr"])[None], r_rast, tri)
    covered = r_rast[..., 3] > 0

    # ---- per op, identical inputs
    res_op = {}
    for who, (m, ctx) in mods.items():
        g = {}
        pos = leaf(cu(c["pos"]))
        rast, db = m.rasterize(ctx, pos, tri, c["res"])
        torch.autograd.backward([rast, db], [randn(rast.shape, 5), randn(db.shape, 6)])
        g["rasterize g_pos (dy+ddb)"] = pos.grad
        pos = leaf(cu(c["pos"]))
        rast, db = m.rasterize(ctx, pos, tri, c["res"], grad_db=False)
        rast.backward(randn(rast.shape, 5))
        g["rasterize g_pos (dy)"] = pos.grad
        del rast, db
        if with_tex:
            uv = leaf(cu(c["uv"])); rr = leaf(r_rast); rdb = leaf(r_db)
            o, oda = m.interpolate(uv[None], rr, uv_idx, rast_db=rdb, diff_attrs="all")
            torch.autograd.backward([o, oda], [randn(o.shape, 8), randn(oda.shape, 9)])
            g["interpolate g_attr"] = uv.grad; g["interpolate g_rast|px"] = rr.grad; g["interpolate g_rast_db|px"] = rdb.grad
            del o, oda
            tex = leaf(cu(c["tex"])); tuv = leaf(r_uv); tda = leaf(r_uvda)
            col = m.texture(tex, tuv, tda, filter_mode="linear-mipmap-linear", max_mip_level=maxmip)
            col.backward(randn(col.shape, 17))
            g["texture g_tex"] = tex.grad; g["texture g_uv|px"] = tuv.grad; g["texture g_uv_da|px"] = tda.grad
            del col
        else:
            attr = leaf(cu(c["attr"])); rr = leaf(r_rast)
            o, _ = m.interpolate(attr[None], rr, tri)
            o.backward(randn(o.shape, 8))
            g["interpolate g_attr"] = attr.grad; g["interpolate g_rast|px"] = rr.grad
            del o
        if cfg != "c4":
            col = leaf(r_color); pos = leaf(cu(c["pos"]))
            aa = m.antialias(col, r_rast, pos, tri)
            aa.backward(randn(aa.shape, 15))
            g["antialias g_pos"] = pos.grad; g["antialias g_color|px"] = col.grad
            del aa
        res_op[who] = g
    # knife-edge predicates for per-pixel tensors
    excl = {}
    if with_tex:
        # mip level = 0.5*__log2f(..) floors to an integer level: pixels whose continuous level sits within 1e-3 of an integer
        # may pick the neighbouring level pair when the ellipse terms differ in the last bit; evaluated from
        # the reference's own uv_da in float64
        q = r_uvda.double() * float(tex_res)
        A = q[..., 0] ** 2 + q[..., 2] ** 2; B = q[..., 1] ** 2 + q[..., 3] ** 2; Cc = q[..., 0] * q[..., 1] + q[..., 2] * q[..., 3]
        lam = 0.5 * (A + B) + torch.sqrt(0.25 * (A - B) ** 2 + Cc ** 2)
        lvl = 0.5 * torch.log2(torch.clamp(lam, min=1e-300))
        frac = (lvl - torch.round(lvl)).abs()
        excl["texture g_uv|px"] = excl["texture g_uv_da|px"] = (frac < 1e-3) & (lvl > -1e-3) & (lvl < maxmip + 1e-3)
    for k in res_op["ref"]:
        kind = "pixel" if k.endswith("|px") else "sum"
        report.append(entry(k.replace("|px", ""), res_op["ours"][k], res_op["ref"][k], kind, excl.get(k)))
    del res_op

    # ---- whole chain, each implementation on its own intermediates
    chain = {}
    for who, (m, ctx) in mods.items():
        pos = leaf(cu(c["pos"]))
        rast, db = m.rasterize(ctx, pos, tri, c["res"])
        if with_tex:
            uv = leaf(cu(c["uv"])); tex = leaf(cu(c["tex"]))
            tc, td = m.interpolate(uv[None], rast, uv_idx, rast_db=db, diff_attrs="all")
            col = m.texture(tex, tc, td, filter_mode="linear-mipmap-linear", max_mip_level=maxmip)
            out = m.antialias(col, rast, pos, tri)
            leaves = {"pos": pos, "uv": uv, "tex": tex}
        else:
            attr = leaf(cu(c["attr"]))
            out, _ = m.interpolate(attr[None], rast, tri)
            if cfg == "c2":
                out = m.antialias(out, rast, pos, tri)
            leaves = {"pos": pos, "attr": attr}
        out.backward(randn(out.shape, 19))
        chain[who] = ({k: v.grad for k, v in leaves.items()}, out.detach())
        del out
    report.append(entry("chain out", chain["ours"][1], chain["ref"][1], "pixel"))
    for k in chain["ref"][0]:
        report.append(entry("chain g_" + k, chain["ours"][0][k], chain["ref"][0][k], "sum"))
    torch.cuda.synchronize()
    print(json.dumps({"cfg": cfg, "n": n, "res": res, "coverage": float(covered.double().mean()), "report": report}))


if __name__ == "__main__":
    main()
