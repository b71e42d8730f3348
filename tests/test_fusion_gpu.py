"""Fused chain backwards (texture -> interpolate -> rasterize in one kernel, nvdiffrast_b200/torch/ops.py `_fusion_plan`;
interpolate -> rasterize in one kernel, `_interp_fusion_plan`):
same results as the op-by-op path, engaged only when the provenance of uv / uv_da / rast / rast_db is proven, and summed
correctly with every other gradient path.  Tolerance: forward 1e-6 (same kernels; antialias' atomic blend order varies); accumulated gradients
|d|_inf <= 2e-5 |g|_inf (only the summation order differs); per-pixel gradients abs 1e-6 |g|_inf + rel 1e-5."""
import numpy as np
import pytest
import torch

import workloads as wl

pytestmark = pytest.mark.gpu


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.fixture(scope="module")
def dr():
    import nvdiffrast_b200.torch as dr
    return dr


@pytest.fixture(scope="module")
def ctx(dr):
    return dr.RasterizeCudaContext()


def _ops_called(fn):
    from nvdiffrast_b200 import _lib
    trace = []
    _lib.set_trace(trace)
    try:
        r = fn()
        torch.cuda.synchronize()
    finally:
        _lib.set_trace(None)
    return r, [t[0].replace("b200dr_", "") for t in trace]


def _chain(dr, ctx, c, n, *, filter_mode="linear-mipmap-linear", boundary_mode="wrap", need=("pos", "uv", "tex"), attr_mode="broadcast",
           extra_uv_loss=False, antialias=True, grad_db=True, tamper=None, range_mode=False):
    """One forward + backward of the c3 chain; returns (out, grads, ops called during backward)."""
    pos_np = c["pos"][:n]
    tri = cu(c["tri"]); uv_idx = cu(c["uv_idx"])
    if range_mode:                                              # range mode: one shared vertex buffer, the same triangles for every image
        pos = cu(pos_np[0]).requires_grad_("pos" in need)
        ranges = torch.tensor([[0, c["tri"].shape[0]]] * n, dtype=torch.int32)
        rast, db = dr.rasterize(ctx, pos, tri, c["res"], ranges=ranges, grad_db=grad_db)
    else:
        pos = cu(pos_np).requires_grad_("pos" in need)
        rast, db = dr.rasterize(ctx, pos, tri, c["res"], grad_db=grad_db)
    uv = cu(c["uv"])
    if attr_mode == "per_image":
        uv = uv[None].repeat(n, 1, 1).contiguous() * torch.linspace(1.0, 1.1, n, device="cuda")[:, None, None]
    elif attr_mode == "broadcast":
        uv = uv[None].contiguous()
    uv = uv.requires_grad_("uv" in need)                        # range mode with attr_mode == "range": [V,2]
    tex = cu(c["tex"]).requires_grad_("tex" in need)
    mip = "mipmap" in filter_mode
    texc, texd = dr.interpolate(uv, rast, uv_idx, rast_db=db if mip else None, diff_attrs="all" if mip else None)
    if tamper == "view":
        texc = texc[..., :2]
    elif tamper == "inplace":
        texc.mul_(1.0)
    elif tamper == "retain":
        texc.retain_grad()
    elif tamper == "hook":
        texd.register_hook(lambda g: g)
    color = dr.texture(tex, texc, texd if mip else None, filter_mode=filter_mode, boundary_mode=boundary_mode, max_mip_level=5 if mip else None)
    out = dr.antialias(color, rast, pos, tri) if (antialias and not range_mode) else color
    g = torch.Generator(device="cuda"); g.manual_seed(19)
    dy = torch.randn(out.shape, device="cuda", generator=g)
    loss = (out * dy).sum()
    if extra_uv_loss:
        loss = loss + (texc * texc).sum() * 3.0 + (rast[..., :2] * 0.5).sum()      # second consumers of the intermediates
    _, ops = _ops_called(lambda: loss.backward())
    grads = {"pos": pos.grad, "uv": uv.grad, "tex": tex.grad}
    return out.detach(), {k: v for k, v in grads.items() if v is not None}, ops


def _compare(dr, ctx, c, n, expect_fused=True, expect_raster_fused=True, **kw):
    dr.set_fusion(True)
    try:
        out_f, g_f, ops_f = _chain(dr, ctx, c, n, **kw)
    finally:
        dr.set_fusion(True)
    dr.set_fusion(False)
    try:
        out_u, g_u, ops_u = _chain(dr, ctx, c, n, **kw)
    finally:
        dr.set_fusion(True)
    assert "texture_chain_grad" not in ops_u and "texture_grad" in ops_u
    assert ("texture_chain_grad" in ops_f) == expect_fused, ops_f
    if expect_fused:
        assert "texture_grad" not in ops_f
        if expect_raster_fused and not kw.get("extra_uv_loss"):
            assert "interpolate_grad" not in ops_f and "rasterize_grad" not in ops_f, ops_f
    # the forward kernels are the same in both modes; antialias blends with fp32 atomics whose order varies from run to run
    assert float((out_f - out_u).abs().max()) <= 1e-6
    assert set(g_f) == set(g_u) and len(g_u) > 0
    for k in g_u:
        d = float((g_f[k].double() - g_u[k].double()).abs().max()); sc = float(g_u[k].abs().max())
        assert sc > 0 and d <= 2e-5 * sc, "%s: |d|_inf %.3g vs |g|_inf %.3g" % (k, d, sc)
    return ops_f


@pytest.fixture(scope="module")
def c3():
    return wl.config_c3(n=2, res=384, tex_res=64)


def test_fused_chain_matches_op_by_op(dr, ctx, c3):
    ops = _compare(dr, ctx, c3, 2)
    assert ops.count("texture_chain_grad") == 1


@pytest.mark.parametrize("kw", [dict(filter_mode="linear"), dict(filter_mode="linear-mipmap-nearest"), dict(boundary_mode="clamp"),
                                dict(boundary_mode="zero"), dict(grad_db=False), dict(attr_mode="per_image"), dict(antialias=False)])
def test_fused_chain_modes(dr, ctx, c3, kw):
    _compare(dr, ctx, c3, 2, **kw)


@pytest.mark.parametrize("need", [("pos",), ("uv",), ("tex",), ("pos", "tex"), ("uv", "tex")])
def test_fused_chain_partial_gradients(dr, ctx, c3, need):
    # only the texture needs a gradient: nothing flows past the texture op, so the intermediates carry no provenance and the
    # plain texture backward runs (without writing the unused g_uv / g_uv_da)
    _compare(dr, ctx, c3, 2, need=need, antialias=False, expect_fused=(need != ("tex",)))


def test_fused_chain_with_other_consumers_of_the_intermediates(dr, ctx, c3):
    """uv and rast feed the loss directly as well: their own backward nodes run for those paths, the fused node for the
    texture path, autograd adds them up."""
    ops = _compare(dr, ctx, c3, 2, extra_uv_loss=True)
    # texc's own path goes through the fused interpolate -> rasterize node, rast's direct consumer through rasterize's node
    assert "interpolate_chain_grad" in ops and "rasterize_grad" in ops


def test_fused_chain_stops_at_rast_in_range_mode(dr, ctx, c3):
    """Range-mode geometry (pos [V,4] + ranges): the fused kernel covers texture + interpolate and hands g_rast / g_rast_db to
    rasterize's own backward."""
    ops = _compare(dr, ctx, c3, 2, range_mode=True, attr_mode="range", expect_raster_fused=False)
    assert "rasterize_grad" in ops and "interpolate_grad" not in ops


@pytest.mark.parametrize("tamper", ["view", "inplace", "retain", "hook"])
def test_unproven_provenance_takes_the_op_by_op_path(dr, ctx, c3, tamper):
    _compare(dr, ctx, c3, 2, tamper=tamper, expect_fused=False)


def test_fullsize_fused_chain_c3(dr, ctx):
    """BASELINE's c3 size (2048^2, 4 images, 1024^2 texture): fused == op-by-op."""
    c = wl.config_c3(n=4, res=2048, tex_res=1024)
    dr.set_fusion(True)
    _, g_f, ops = _chain(dr, ctx, c, 4)
    assert "texture_chain_grad" in ops and "interpolate_grad" not in ops and "rasterize_grad" not in ops
    dr.set_fusion(False)
    try:
        _, g_u, _ = _chain(dr, ctx, c, 4)
    finally:
        dr.set_fusion(True)
    for k in g_u:
        d = float((g_f[k].double() - g_u[k].double()).abs().max()); sc = float(g_u[k].abs().max())
        assert d <= 5e-5 * sc, "%s: |d|_inf %.3g vs |g|_inf %.3g" % (k, d, sc)


# ----------------------------------------------------------------------------------------------
# interpolate -> rasterize in one kernel (b200dr_interpolate_chain_grad)


def _interp_chain(dr, ctx, pos_np, tri_np, attr_np, res, *, da=False, attr_mode="broadcast", grad_db=True, need=("pos", "attr"), tamper=None,
                  peel_layers=0, use_db_loss=False):
    tri = cu(tri_np)
    pos = cu(pos_np).requires_grad_("pos" in need)
    n = pos.shape[0]
    if attr_mode == "per_image":
        attr = cu(attr_np)[None].repeat(n, 1, 1).contiguous() * torch.linspace(1.0, 1.3, n, device="cuda")[:, None, None]
    elif attr_mode == "broadcast":
        attr = cu(attr_np)[None].contiguous()
    else:                                                       # range-mode attributes [V, A] with instanced-mode positions
        attr = cu(attr_np)
    attr = attr.requires_grad_("attr" in need)
    gen = torch.Generator(device="cuda"); gen.manual_seed(23)
    outs = []

    def shade(rast, db):
        if tamper == "view":
            rast = rast[:, :, :, :]
        elif tamper == "retain":
            rast.retain_grad()
        elif tamper == "foreign_db":
            db = db * 1.0
        o, oda = dr.interpolate(attr, rast, tri, rast_db=db if da else None, diff_attrs="all" if da else None)
        outs.append(o); loss = (o * torch.randn(o.shape, device="cuda", generator=gen)).sum()
        if da:
            outs.append(oda); loss = loss + (oda * torch.randn(oda.shape, device="cuda", generator=gen)).sum()
        if use_db_loss:
            loss = loss + (db * torch.randn(db.shape, device="cuda", generator=gen)).sum() + (rast[..., :3] * 0.25).sum()
        return loss

    if peel_layers:
        loss = 0.0
        with dr.DepthPeeler(ctx, pos, tri, res, grad_db=grad_db) as peeler:
            for _ in range(peel_layers):
                rast, db = peeler.rasterize_next_layer()
                loss = loss + shade(rast, db)
    else:
        rast, db = dr.rasterize(ctx, pos, tri, res, grad_db=grad_db)
        loss = shade(rast, db)
    _, ops = _ops_called(lambda: loss.backward())
    grads = {"pos": pos.grad, "attr": attr.grad}
    return [o.detach() for o in outs], {k: v for k, v in grads.items() if v is not None}, ops


def _compare_interp(dr, ctx, *args, expect_fused=True, tol=2e-5, **kw):
    dr.set_fusion(True)
    outs_f, g_f, ops_f = _interp_chain(dr, ctx, *args, **kw)
    dr.set_fusion(False)
    try:
        outs_u, g_u, ops_u = _interp_chain(dr, ctx, *args, **kw)
    finally:
        dr.set_fusion(True)
    assert "interpolate_chain_grad" not in ops_u and "interpolate_grad" in ops_u
    assert ("interpolate_chain_grad" in ops_f) == expect_fused, ops_f
    if expect_fused and not kw.get("use_db_loss"):
        assert "interpolate_grad" not in ops_f and "rasterize_grad" not in ops_f, ops_f
    for a, b in zip(outs_f, outs_u):
        assert torch.equal(a, b)                                # the forward kernel is the same
    assert set(g_f) == set(g_u) and len(g_u) > 0
    for k in g_u:
        d = float((g_f[k].double() - g_u[k].double()).abs().max()); sc = float(g_u[k].abs().max())
        assert sc > 0 and d <= tol * sc, "%s: |d|_inf %.3g vs |g|_inf %.3g" % (k, d, sc)
    return ops_f


@pytest.fixture(scope="module")
def c2s():
    c = wl.config_c2(n=3, res=320)
    return c["pos"], c["tri"], c["attr"], c["res"]


@pytest.fixture(scope="module")
def sph():
    """c3's sphere (2 930 vertices: position gradients accumulate in global memory), 3 attributes from the uv table."""
    c = wl.config_c3(n=2, res=384, tex_res=64)
    rng = np.random.RandomState(5)
    return c["pos"], c["tri"], rng.randn(c["pos"].shape[1], 4).astype(np.float32), c["res"]


@pytest.mark.parametrize("A", [1, 2, 3, 4])
@pytest.mark.parametrize("da", [False, True])
def test_interp_chain_matches_op_by_op_small_mesh(dr, ctx, c2s, A, da):
    """c2's cube (8 vertices: shared-memory accumulators for both gradients)."""
    pos, tri, attr, res = c2s
    rng = np.random.RandomState(A)
    _compare_interp(dr, ctx, pos, tri, rng.randn(attr.shape[0], A).astype(np.float32), res, da=da)


@pytest.mark.parametrize("A", [1, 2, 3, 4])
@pytest.mark.parametrize("da", [False, True])
def test_interp_chain_matches_op_by_op_sphere(dr, ctx, sph, A, da):
    pos, tri, attr, res = sph
    _compare_interp(dr, ctx, pos, tri, np.ascontiguousarray(attr[:, :A]), res, da=da)


@pytest.mark.parametrize("kw", [dict(attr_mode="per_image"), dict(attr_mode="range"), dict(grad_db=False, da=True), dict(need=("pos",)),
                                dict(attr_mode="per_image", da=True), dict(peel_layers=3, da=True), dict(use_db_loss=True, da=True)])
def test_interp_chain_modes(dr, ctx, sph, kw):
    pos, tri, attr, res = sph
    _compare_interp(dr, ctx, pos, tri, np.ascontiguousarray(attr[:, :3]), res, **kw)


@pytest.mark.parametrize("kw", [dict(tamper="view"), dict(tamper="retain"), dict(tamper="foreign_db", da=True), dict(need=("attr",))])
def test_interp_chain_unproven_provenance_takes_the_op_by_op_path(dr, ctx, sph, kw):
    pos, tri, attr, res = sph
    _compare_interp(dr, ctx, pos, tri, np.ascontiguousarray(attr[:, :3]), res, expect_fused=False, **kw)


def test_interp_chain_partial_diff_list_takes_the_op_by_op_path(dr, ctx, sph):
    pos_np, tri_np, attr_np, res = sph
    pos = cu(pos_np).requires_grad_(True); attr = cu(attr_np[:, :3].copy())[None].requires_grad_(True)
    rast, db = dr.rasterize(ctx, pos, cu(tri_np), res)
    o, oda = dr.interpolate(attr, rast, cu(tri_np), rast_db=db, diff_attrs=[0, 2])
    _, ops = _ops_called(lambda: (o.sum() + oda.sum()).backward())
    assert "interpolate_chain_grad" not in ops and "interpolate_grad" in ops and "rasterize_grad" in ops


def test_fullsize_interp_chain_c2_and_c4(dr, ctx):
    """BASELINE's c2 (1024^2 x 8 here) and c4 (1M triangles, 2048^2 x 1): fused == op-by-op."""
    c = wl.config_c2(n=8, res=1024)
    _compare_interp(dr, ctx, c["pos"], c["tri"], c["attr"], c["res"], tol=5e-5)
    c = wl.config_c4(n=1, res=2048)
    _compare_interp(dr, ctx, c["pos"], c["tri"], c["attr"], c["res"], tol=5e-5)
