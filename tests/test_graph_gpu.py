"""A whole forward + backward step records into ONE CUDA graph (nvdiffrast_b200/graphs.py): nothing in the op surface
synchronises, reads device values back or allocates outside torch's allocator.  Replays on new input VALUES must reproduce the
eager results (forward 1e-6; accumulated gradients 1e-4 |g|_inf -- only the atomic summation order differs between runs)."""
import numpy as np
import pytest
import torch

import workloads as wl

pytestmark = pytest.mark.gpu


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _rel(a, b):
    return float((a - b).abs().max()) / max(float(b.abs().max()), 1e-30)


@pytest.mark.parametrize("fusion", [True, False])
def test_full_chain_replays_from_a_cuda_graph(fusion):
    import nvdiffrast_b200.torch as dr
    from nvdiffrast_b200 import graphs
    prev = dr.get_fusion()
    dr.set_fusion(fusion)
    try:
        c = wl.config_c3(n=3, res=96, tex_res=64)
        c2 = wl.config_c3(n=3, res=96, tex_res=64, seed=11)            # other cameras: different coverage, same shapes
        ctx = dr.RasterizeCudaContext()
        tri, uv_idx = cu(c["tri"]), cu(c["uv_idx"])
        topo = dr.antialias_construct_topology_hash(tri)
        pos = cu(c["pos"]).requires_grad_(True)
        uv = cu(c["uv"]).requires_grad_(True)
        tex = cu(c["tex"]).requires_grad_(True)
        dy = torch.randn((3, 96, 96, 3), device="cuda", generator=torch.Generator(device="cuda").manual_seed(1))

        def step():
            rast, db = dr.rasterize(ctx, pos, tri, c["res"])
            texc, texd = dr.interpolate(uv[None], rast, uv_idx, rast_db=db, diff_attrs="all")
            color = dr.texture(tex, texc, texd, filter_mode="linear-mipmap-linear")
            out = dr.antialias(color, rast, pos, tri, topology_hash=topo)
            g_pos, g_uv, g_tex = torch.autograd.grad(out, [pos, uv, tex], dy)
            return out, rast, g_pos, g_uv, g_tex

        captured = graphs.capture(step)
        for scene, texmul in ((c, 1.0), (c2, 0.5), (c, 0.25)):
            with torch.no_grad():
                pos.copy_(cu(scene["pos"])); tex.copy_(cu(scene["tex"]) * texmul)
            got = [t.clone() for t in captured.replay()]
            want = step()
            torch.cuda.synchronize()
            assert torch.equal(got[1][..., 3], want[1][..., 3])                           # triangle ids
            assert float((got[0] - want[0]).abs().max()) <= 1e-6 * max(1.0, float(want[0].abs().max()))
            for name, a, b in zip(("g_pos", "g_uv", "g_tex"), got[2:], want[2:]):
                assert _rel(a, b) <= 1e-4, (name, _rel(a, b))
            assert float(got[4].abs().sum()) > 0
    finally:
        dr.set_fusion(prev)


def test_interpolate_chain_and_depth_peeling_replay_from_a_cuda_graph():
    """c2-shaped chain (rasterize -> interpolate, fused backward) and two peeled layers in one graph."""
    import nvdiffrast_b200.torch as dr
    from nvdiffrast_b200 import graphs
    c = wl.config_c2(n=4, res=64)
    ctx = dr.RasterizeCudaContext()
    tri = cu(c["tri"])
    pos = cu(c["pos"]).requires_grad_(True)
    attr = cu(c["attr"]).requires_grad_(True)
    dy = torch.randn((4, 64, 64, 3), device="cuda", generator=torch.Generator(device="cuda").manual_seed(2))

    def step():
        rast, _ = dr.rasterize(ctx, pos, tri, c["res"])
        out, _ = dr.interpolate(attr[None], rast, tri)
        g_pos, g_attr = torch.autograd.grad(out, [pos, attr], dy)
        with dr.DepthPeeler(ctx, pos.detach(), tri, c["res"]) as peeler:
            l0, _ = peeler.rasterize_next_layer()
            l1, _ = peeler.rasterize_next_layer()
        return out, g_pos, g_attr, l0, l1

    captured = graphs.capture(step)
    with torch.no_grad():
        pos.copy_(cu(wl.config_c2(n=4, res=64, seed=3)["pos"]))
    got = [t.clone() for t in captured.replay()]
    want = step()
    torch.cuda.synchronize()
    assert float((got[0] - want[0]).abs().max()) <= 1e-6
    assert _rel(got[1], want[1]) <= 1e-4 and _rel(got[2], want[2]) <= 1e-4
    assert torch.equal(got[3][..., 3], want[3][..., 3]) and torch.equal(got[4][..., 3], want[4][..., 3])
    assert float((got[4][..., 3] > 0).float().sum()) > 0                                   # the second layer is not empty
