"""Small inverse-rendering loops, the way a caller uses the op surface (the reference's samples/torch/earth.py and cube.py in
miniature, on seeded synthetic data): a texture learned from renders, and vertex positions + colours learned from renders.

The same loop runs on two back ends: `B200Backend` (the public API on the GPU, gradients by autograd) and `OracleBackend`
(the CPU oracle, gradients by its restated backward passes).  The CPU run pins what the loop is expected to do without a
GPU (tests/test_fit_cpu.py); the GPU run must converge the same way (tests/test_fit_gpu.py).  Adam is torch's own on both.
"""
import numpy as np
import torch

import workloads as wl


# ----------------------------------------------------------------------------------------------
# scenes


def texture_scene(res=64, tex_res=32, n_views=8, seed=5):
    """Bumped sphere (720 triangles) with a seam-duplicated uv index buffer; target texture = smooth colour pattern."""
    v, tri, uv, uv_idx = wl.sphere_mesh(24, 16)
    mvp = wl.camera_batch(n_views, seed, far=200.0)
    y, x = np.meshgrid(np.arange(tex_res) / tex_res, np.arange(tex_res) / tex_res, indexing="ij")
    tex = np.stack([0.5 + 0.5 * np.sin(6.28 * (2 * x + y)), 0.5 + 0.5 * np.cos(6.28 * (x - 2 * y)), 0.25 + 0.5 * ((np.floor(4 * x) + np.floor(4 * y)) % 2)], -1)
    return dict(verts=v.astype(np.float32), tri=tri, uv=uv, uv_idx=uv_idx, mvp=mvp, res=(res, res),
                tex_target=tex[None].astype(np.float32), tex_init=np.full((1, tex_res, tex_res, 3), 0.5, np.float32))


def cube_scene(res=64, n_views=8, seed=7, noise=0.08):
    """The cube of samples/torch/cube.py: vertex positions and vertex colours are both unknown."""
    v, tri = wl.cube_mesh()
    rng = np.random.default_rng(seed)
    mvp = wl.camera_batch(n_views, seed)
    col = rng.uniform(size=(8, 3)).astype(np.float32)
    return dict(verts_target=v.astype(np.float32), col_target=col, tri=tri, mvp=mvp, res=(res, res),
                verts_init=(v + rng.normal(scale=noise, size=v.shape)).astype(np.float32),
                col_init=rng.uniform(size=(8, 3)).astype(np.float32))


def _clip(verts, mvp):
    """[V,3] x [n,4,4] -> clip-space positions [n,V,4]; torch tensors (autograd-friendly on the GPU back end)."""
    vh = torch.cat([verts, torch.ones_like(verts[:, :1])], 1)
    return torch.matmul(vh[None], mvp.transpose(1, 2)).contiguous()


# ----------------------------------------------------------------------------------------------
# back ends: render(params) -> image, and loss_and_grads(params, target) -> (loss, grads)


class B200Backend:
    """The product path: nvdiffrast_b200.torch on cuda:0."""

    def __init__(self, dr, device):
        self.dr, self.dev = dr, device
        self.ctx = dr.RasterizeCudaContext(device)

    def t(self, a):
        return torch.from_numpy(np.ascontiguousarray(a)).to(self.dev)

    def render_texture(self, s, tex, mvp):
        dr = self.dr
        pos = _clip(s["verts_t"], mvp)
        rast, db = dr.rasterize(self.ctx, pos, s["tri_t"], s["res"])
        uv, uvd = dr.interpolate(s["uv_t"][None], rast, s["uv_idx_t"], rast_db=db, diff_attrs="all")
        col = dr.texture(tex, uv, uvd, filter_mode="linear-mipmap-linear", max_mip_level=3)
        return dr.antialias(col, rast, pos, s["tri_t"])

    def render_cube(self, s, verts, col, mvp):
        dr = self.dr
        pos = _clip(verts, mvp)
        rast, _ = dr.rasterize(self.ctx, pos, s["tri_t"], s["res"])
        c, _ = dr.interpolate(col[None], rast, s["tri_t"])
        return dr.antialias(c, rast, pos, s["tri_t"])


class OracleBackend:
    """CPU oracle with hand-chained backward passes (test infrastructure)."""

    def __init__(self):
        import oracle
        oracle.build()
        self.o = oracle

    @staticmethod
    def _pos(verts, mvp):
        return wl.transform(mvp, verts)

    def texture_fwd_bwd(self, s, tex, mvp, target=None):
        o = self.o
        pos = self._pos(s["verts"], mvp)
        r = o.rasterize_fwd(pos, s["tri"], s["res"])
        rast, db = r["rast"], r["rast_db"]
        uv, uvd = o.interpolate_fwd(s["uv"][None], rast, s["uv_idx"], db, "all")
        mip = o.texture_construct_mip(tex, 3)
        col = o.texture_fwd(tex, uv, uvd, None, mip=mip, filter_mode="linear-mipmap-linear")
        out, items = o.antialias_fwd(col, rast, pos, s["tri"])
        if target is None:
            return out
        d = out - target
        loss = float(np.mean(d.astype(np.float64) ** 2))
        dy = (2.0 / d.size * d).astype(np.float32)
        g_col, _ = o.antialias_grad(col, rast, pos, s["tri"], dy, items)
        tg = o.texture_grad(tex, uv, g_col, uvd, None, mip=mip, filter_mode="linear-mipmap-linear")
        return loss, tg["g_tex"]

    def cube_fwd_bwd(self, s, verts, col, mvp, target=None):
        o = self.o
        pos = self._pos(verts, mvp)
        r = o.rasterize_fwd(pos, s["tri"], s["res"])
        rast = r["rast"]
        c, _ = o.interpolate_fwd(col[None], rast, s["tri"])
        out, items = o.antialias_fwd(c, rast, pos, s["tri"])
        if target is None:
            return out
        d = out - target
        loss = float(np.mean(d.astype(np.float64) ** 2))
        dy = (2.0 / d.size * d).astype(np.float32)
        g_c, g_pos = o.antialias_grad(c, rast, pos, s["tri"], dy, items)
        g_col, g_rast, _ = o.interpolate_grad(col[None], rast, s["tri"], g_c)
        g_pos = g_pos + o.rasterize_grad(pos, s["tri"], rast, g_rast, None)
        # pos[n] = [verts, 1] mvp[n]^T  ->  g_verts = sum_n (g_pos[n] mvp[n])[:, :3]
        g_verts = np.einsum("nvi,nij->vj", g_pos.astype(np.float64), mvp.astype(np.float64))[:, :3].astype(np.float32)
        return loss, g_verts, g_col[0]


# ----------------------------------------------------------------------------------------------
# the loops (identical hyper-parameters on both back ends)

TEXTURE_FIT = dict(steps=60, views_per_step=4, lr=0.05)
CUBE_FIT = dict(steps=120, views_per_step=4, lr_verts=0.01, lr_col=0.03)


def _views(step, k, n):
    return [(step * k + i) % n for i in range(k)]


def fit_texture_oracle(s, cfg=TEXTURE_FIT):
    be = OracleBackend()
    tex = torch.from_numpy(s["tex_init"].copy())
    opt = torch.optim.Adam([tex], lr=cfg["lr"])
    targets = be.texture_fwd_bwd(s, s["tex_target"], s["mvp"])
    losses = []
    for it in range(cfg["steps"]):
        vi = _views(it, cfg["views_per_step"], len(s["mvp"]))
        loss, g = be.texture_fwd_bwd(s, tex.numpy(), s["mvp"][vi], targets[vi])
        tex.grad = torch.from_numpy(np.ascontiguousarray(g))
        opt.step()
        losses.append(loss)
    return losses, tex.numpy()


def fit_texture_b200(s, dr, device, cfg=TEXTURE_FIT):
    be = B200Backend(dr, device)
    s = dict(s, verts_t=be.t(s["verts"]), tri_t=be.t(s["tri"]), uv_t=be.t(s["uv"]), uv_idx_t=be.t(s["uv_idx"]))
    mvp = be.t(s["mvp"])
    with torch.no_grad():
        targets = be.render_texture(s, be.t(s["tex_target"]), mvp)
    tex = be.t(s["tex_init"]).requires_grad_(True)
    opt = torch.optim.Adam([tex], lr=cfg["lr"])
    losses = []
    for it in range(cfg["steps"]):
        vi = torch.tensor(_views(it, cfg["views_per_step"], len(s["mvp"])), device=device)
        opt.zero_grad(set_to_none=True)
        loss = torch.mean((be.render_texture(s, tex, mvp[vi]) - targets[vi]) ** 2)
        loss.backward()
        opt.step()
        losses.append(float(loss))
    return losses, tex.detach().cpu().numpy()


def fit_cube_oracle(s, cfg=CUBE_FIT):
    be = OracleBackend()
    verts = torch.from_numpy(s["verts_init"].copy()); col = torch.from_numpy(s["col_init"].copy())
    opt = torch.optim.Adam([{"params": [verts], "lr": cfg["lr_verts"]}, {"params": [col], "lr": cfg["lr_col"]}])
    targets = be.cube_fwd_bwd(s, s["verts_target"], s["col_target"], s["mvp"])
    losses = []
    for it in range(cfg["steps"]):
        vi = _views(it, cfg["views_per_step"], len(s["mvp"]))
        loss, gv, gc = be.cube_fwd_bwd(s, verts.numpy(), col.numpy(), s["mvp"][vi], targets[vi])
        verts.grad = torch.from_numpy(np.ascontiguousarray(gv)); col.grad = torch.from_numpy(np.ascontiguousarray(gc))
        opt.step()
        losses.append(loss)
    return losses, verts.numpy(), col.numpy()


def fit_cube_b200(s, dr, device, cfg=CUBE_FIT):
    be = B200Backend(dr, device)
    s = dict(s, tri_t=be.t(s["tri"]))
    mvp = be.t(s["mvp"])
    with torch.no_grad():
        targets = be.render_cube(s, be.t(s["verts_target"]), be.t(s["col_target"]), mvp)
    verts = be.t(s["verts_init"]).requires_grad_(True); col = be.t(s["col_init"]).requires_grad_(True)
    opt = torch.optim.Adam([{"params": [verts], "lr": cfg["lr_verts"]}, {"params": [col], "lr": cfg["lr_col"]}])
    losses = []
    for it in range(cfg["steps"]):
        vi = torch.tensor(_views(it, cfg["views_per_step"], len(s["mvp"])), device=device)
        opt.zero_grad(set_to_none=True)
        loss = torch.mean((be.render_cube(s, verts, col, mvp[vi]) - targets[vi]) ** 2)
        loss.backward()
        opt.step()
        losses.append(float(loss))
    return losses, verts.detach().cpu().numpy(), col.detach().cpu().numpy()


def geometric_error(verts, target):
    return float(np.abs(np.asarray(verts, np.float64) - target).mean())
