"""Fused chain backward (texture -> interpolate -> rasterize in one kernel, nvdiffrast_b200/torch/ops.py `_fusion_plan`):
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
    assert "interpolate_grad" in ops and "rasterize_grad" in ops


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
