"""Seeded synthetic scenes shared by tests/, bench.py and tests/golden/make_golden.py (numpy only, no data files).

Recipes follow SURVEY.md section 8(d): the reference ships no benchmark inputs, its sample meshes are third-party
data, and "Spot" is not in the repository, so every configuration named in BASELINE.json is regenerated here.
"""
import numpy as np


# ----------------------------------------------------------------------------------------------
# cameras


def projection(x=0.4, n=1.0, f=50.0):
    """OpenGL-style perspective matrix with half-width x at the near plane (same convention as the reference samples)."""
    return np.array([[n / x, 0, 0, 0],
                     [0, n / x, 0, 0],
                     [0, 0, -(f + n) / (f - n), -(2 * f * n) / (f - n)],
                     [0, 0, -1, 0]], np.float32)


def translate(x, y, z):
    m = np.eye(4, dtype=np.float32)
    m[:3, 3] = (x, y, z)
    return m


def random_rotations(n, rng):
    """n random rotation matrices (QR of a normal 3x3, determinant fixed to +1) as 4x4."""
    out = np.zeros((n, 4, 4), np.float32)
    for i in range(n):
        q, r = np.linalg.qr(rng.normal(size=(3, 3)))
        q = q * np.sign(np.diag(r))
        if np.linalg.det(q) < 0:
            q[:, 0] = -q[:, 0]
        out[i, :3, :3] = q
        out[i, 3, 3] = 1
    return out


def transform(mvp, verts):
    """mvp [N,4,4], verts [V,3] -> clip-space positions [N,V,4] float32."""
    v = np.concatenate([verts, np.ones((verts.shape[0], 1), verts.dtype)], 1).astype(np.float32)
    return np.ascontiguousarray(np.einsum('nij,vj->nvi', mvp.astype(np.float32), v), dtype=np.float32)


def camera_batch(n, seed=0, far=50.0, dist=3.5, x=0.4):
    rng = np.random.default_rng(seed)
    rot = random_rotations(n, rng)
    return np.stack([projection(x, 1.0, far) @ translate(0, 0, -dist) @ rot[i] for i in range(n)]).astype(np.float32)


# ----------------------------------------------------------------------------------------------
# meshes


def cube_mesh():
    """Unit cube +-0.5, 8 welded vertices, 12 CCW-outward triangles."""
    v = np.array([[x, y, z] for z in (-.5, .5) for y in (-.5, .5) for x in (-.5, .5)], np.float32)
    quads = [(0, 2, 3, 1), (4, 5, 7, 6), (0, 1, 5, 4), (2, 6, 7, 3), (0, 4, 6, 2), (1, 3, 7, 5)]
    tri = []
    for a, b, c, d in quads:
        tri += [(a, b, c), (a, c, d)]
    return v, np.array(tri, np.int32)


def sphere_mesh(slices=48, stacks=62, bump=0.15):
    """Lat-long sphere with a smooth radial bump.  slices x stacks -> T = 2*slices*(stacks-1), V = slices*(stacks-1)+2.
    Also returns a separate uv index buffer with a duplicated seam column (tri != uv_idx, as in the reference's earth.py)."""
    rings = stacks - 1
    th = np.arange(slices) * (2 * np.pi / slices)
    ph = (np.arange(rings) + 1) * (np.pi / stacks)
    T, P = np.meshgrid(th, ph)                                     # [rings, slices]
    r = 1.0 + bump * np.sin(3 * T) * np.sin(2 * P)
    xyz = np.stack([r * np.sin(P) * np.cos(T), r * np.cos(P), r * np.sin(P) * np.sin(T)], -1).reshape(-1, 3)
    verts = np.concatenate([xyz, [[0, 1, 0]], [[0, -1, 0]]], 0).astype(np.float32)
    top, bot = rings * slices, rings * slices + 1

    # uv vertices: (slices+1) per ring (seam duplicated) + one pole vertex per slice for each cap
    su = np.arange(slices + 1) / slices
    uvr = np.stack(np.meshgrid(su, ph / np.pi), -1).reshape(-1, 2)  # [rings*(slices+1), 2]
    uvt = np.stack([(np.arange(slices) + .5) / slices, np.zeros(slices)], -1)
    uvb = np.stack([(np.arange(slices) + .5) / slices, np.ones(slices)], -1)
    uvs = np.concatenate([uvr, uvt, uvb], 0).astype(np.float32)
    uv_top, uv_bot = rings * (slices + 1), rings * (slices + 1) + slices

    tri, uvi = [], []
    for i in range(rings - 1):
        for j in range(slices):
            j1 = (j + 1) % slices
            a, b, c, d = i * slices + j, i * slices + j1, (i + 1) * slices + j1, (i + 1) * slices + j
            ua, ub, uc, ud = i * (slices + 1) + j, i * (slices + 1) + j + 1, (i + 1) * (slices + 1) + j + 1, (i + 1) * (slices + 1) + j
            tri += [(a, b, c), (a, c, d)]
            uvi += [(ua, ub, uc), (ua, uc, ud)]
    for j in range(slices):
        j1 = (j + 1) % slices
        tri.append((top, j1, j)); uvi.append((uv_top + j, j + 1, j))
        lo = (rings - 1) * slices
        ulo = (rings - 1) * (slices + 1)
        tri.append((bot, lo + j, lo + j1)); uvi.append((uv_bot + j, ulo + j, ulo + j + 1))
    return verts, np.array(tri, np.int32), uvs, np.array(uvi, np.int32)


def triangle_scene():
    """The literal tensors of samples/torch/triangle.py:19-21."""
    pos = np.array([[[-0.8, -0.8, 0, 1], [0.8, -0.8, 0, 1], [-0.8, 0.8, 0, 1]]], np.float32)
    col = np.array([[[1, 0, 0], [0, 1, 0], [0, 0, 1]]], np.float32)
    tri = np.array([[0, 1, 2]], np.int32)
    return pos, tri, col


def soup_scene(n_img=2, n_tri=200, seed=0, spread=1.6, wrange=(0.5, 2.0), size=0.5):
    """Random triangle soup in clip space: intersecting, partly outside the frustum (exercises culling and the clipper),
    some behind the camera, some degenerate.  Un-welded: every triangle has its own three vertices."""
    rng = np.random.default_rng(seed)
    ctr = rng.uniform(-spread, spread, size=(n_img, n_tri, 1, 3))
    v = ctr + rng.normal(scale=size, size=(n_img, n_tri, 3, 3))
    w = rng.uniform(wrange[0], wrange[1], size=(n_img, n_tri, 3, 1))
    pos = np.concatenate([v * w, w], -1).astype(np.float32)
    k = max(n_tri // 20, 1)
    pos[:, :k, :, 3] *= rng.choice([1.0, -1.0], size=(n_img, k, 3))     # some vertices behind the camera (w < 0)
    pos[:, k:2 * k, 2, :] = pos[:, k:2 * k, 1, :]                        # degenerate (two equal vertices)
    pos = pos.reshape(n_img, n_tri * 3, 4)
    tri = np.arange(n_tri * 3, dtype=np.int32).reshape(n_tri, 3)
    return np.ascontiguousarray(pos), tri


def grid_aligned_scene(res=32):
    """Vertices exactly on pixel centres / pixel corners and shared edges through pixel centres: pins the fill rule."""
    W = res
    def ndc(px):       # NDC of a position given in pixels (px may be fractional)
        return 2.0 * px / W - 1.0
    quads = []
    # axis-aligned quads whose edges pass exactly through pixel centres (x.5) and pixel corners
    for (x0, y0, x1, y1) in [(2.5, 2.5, 10.5, 10.5), (10.5, 2.5, 18.5, 10.5), (2.5, 10.5, 10.5, 18.5), (12, 12, 20, 20), (20, 12, 28, 20)]:
        quads.append([(x0, y0), (x1, y0), (x1, y1), (x0, y1)])
    verts, tri = [], []
    for q in quads:
        b = len(verts)
        for (x, y) in q:
            verts.append([ndc(x), ndc(y), 0.0, 1.0])
        tri += [(b, b + 1, b + 2), (b, b + 2, b + 3)]
    # diagonal fan sharing one vertex at a pixel centre
    b = len(verts)
    c = (24.5, 4.5)
    ring = [(28.5, 4.5), (28.5, 8.5), (24.5, 8.5), (20.5, 8.5), (20.5, 4.5), (20.5, 0.5), (24.5, 0.5), (28.5, 0.5)]
    verts.append([ndc(c[0]), ndc(c[1]), 0.2, 1.0])
    for (x, y) in ring:
        verts.append([ndc(x), ndc(y), 0.2, 1.0])
    for i in range(len(ring)):
        tri.append((b, b + 1 + i, b + 1 + (i + 1) % len(ring)))
    pos = np.array(verts, np.float32)[None]
    return pos, np.array(tri, np.int32)


def zfight_scene():
    """Coplanar duplicates and equal-depth overlaps: the later triangle must win (impl/FineRaster.inl:352-361)."""
    base = np.array([[-0.6, -0.6, 0.3, 1], [0.6, -0.6, 0.3, 1], [0.6, 0.6, 0.3, 1], [-0.6, 0.6, 0.3, 1]], np.float32)
    shifted = base.copy(); shifted[:, 0] += 0.25
    pos = np.concatenate([base, base, shifted], 0)[None]
    tri = np.array([[0, 1, 2], [0, 2, 3], [4, 5, 6], [4, 6, 7], [8, 9, 10], [8, 10, 11]], np.int32)
    return np.ascontiguousarray(pos), tri


# ----------------------------------------------------------------------------------------------
# BASELINE.json configurations


def config_c2(n=64, res=1024, seed=0):
    """cube (12 tris), res x res, batch n, rasterize + interpolate; instanced mode pos [n,8,4]."""
    v, tri = cube_mesh()
    pos = transform(camera_batch(n, seed), v)
    col = np.random.default_rng(1).uniform(size=(v.shape[0], 3)).astype(np.float32)
    return dict(pos=pos, tri=tri, attr=col, res=(res, res))


def config_c3(n=16, res=2048, seed=0, tex_res=1024, slices=48, stacks=62):
    """'Spot-like' bumped sphere (5856 tris, 2930 verts), full chain rasterize -> interpolate(uv, diff all) ->
    texture (trilinear) -> antialias."""
    v, tri, uv, uv_idx = sphere_mesh(slices, stacks)
    pos = transform(camera_batch(n, seed, far=200.0), v)
    tex = np.random.default_rng(3).uniform(size=(1, tex_res, tex_res, 3)).astype(np.float32)
    return dict(pos=pos, tri=tri, uv=uv, uv_idx=uv_idx, tex=tex, res=(res, res))


def config_c4(n=8, res=2048, seed=0, slices=1000, stacks=501):
    """1M-triangle sphere, rasterize + interpolate (heavy-geometry stress)."""
    v, tri, _, _ = sphere_mesh(slices, stacks)
    pos = transform(camera_batch(n, seed, far=200.0), v)
    col = np.random.default_rng(1).uniform(size=(v.shape[0], 3)).astype(np.float32)
    return dict(pos=pos, tri=tri, attr=col, res=(res, res))


# ----------------------------------------------------------------------------------------------
# Cube-map texture cases (shared by tests/golden/make_golden.py and the GPU parity tests)

CUBE_CASES = [  # filter_mode, use_da, use_bias, channels, texture minibatch, face size
    ("nearest", False, False, 3, 1, 8),
    ("linear", False, False, 1, 2, 8),
    ("linear", False, False, 4, 1, 4),
    ("linear-mipmap-nearest", True, False, 3, 1, 16),
    ("linear-mipmap-linear", True, False, 4, 2, 16),
    ("linear-mipmap-linear", True, True, 1, 1, 8),
    ("linear-mipmap-linear", False, True, 3, 1, 16),
    ("linear", False, False, 5, 1, 8),                  # more than 4 channels: several 128-bit groups per texel
    ("linear-mipmap-linear", True, True, 8, 2, 16),
]


def cube_texture_case(i, N=2, H=40, W=56):
    """Seeded inputs of cube-map case i: directions over the whole sphere (small faces, so seams and corners are hit all the
    time), plus exact face-diagonal ties, axis directions, and a few invalid (zero) directions."""
    fm, use_da, use_bias, C, D, w = CUBE_CASES[i]
    rng = np.random.default_rng(100 + i)
    tex = rng.uniform(size=(D, 6, w, w, C)).astype(np.float32)
    uv = rng.normal(size=(N, H, W, 3)).astype(np.float32)
    uv[0, 0, :8] = [[1, 1, 0.5], [1, -1, 0.2], [-1, 0.3, 1], [0.1, 1, 1], [1, 1, 1], [-1, -1, -1], [1, -1, 1], [-1, 1, -1]]
    uv[0, 1, :6] = [[1, 0, 0], [-1, 0, 0], [0, 1, 0], [0, -1, 0], [0, 0, 1], [0, 0, -1]]
    uv[0, 2, :3] = 0.0
    uv[1] *= rng.uniform(0.05, 20.0, size=(H, W, 1)).astype(np.float32)
    uv_da = (rng.normal(size=(N, H, W, 6)) * 0.05).astype(np.float32) if use_da else None
    if use_da:
        uv_da[1] *= np.linalg.norm(uv[1], axis=-1, keepdims=True)
    bias = rng.uniform(-1, 3, size=(N, H, W)).astype(np.float32) if use_bias else None
    return dict(filter_mode=fm, boundary_mode="cube", tex=tex, uv=uv, uv_da=uv_da, bias=bias)
