"""CPU tests of the oracle itself: closed-form cases that pin the conventions of SURVEY.md Appendix A, internal
consistency (finite differences in float64-accumulated gradients), and the golden vectors produced by the real
reference on a B200 (tests/golden/*.npz, see tests/golden/make_golden.py)."""
import glob
import os

import numpy as np
import pytest

import oracle
import workloads as wl

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "*.npz")))


def test_triangle_sample_closed_form():
    pos, tri, col = wl.triangle_scene()
    r = oracle.rasterize_fwd(pos, tri, (256, 256))
    ids = r["tri_id"][0]
    # right triangle, legs 204.8 px from 25.6: rows of 203, 202, ... 1 covered pixel centres -> 203*204/2
    assert int((ids > 0).sum()) == 203 * 204 // 2
    assert ids[26, 26] == 1 and ids[25, 26] == 0 and ids[26, 25] == 0      # rows are bottom-up; first covered pixel centre is 26.5
    # barycentrics: u = weight of vertex 0 at (-0.8,-0.8): pixel (26,26) is almost that corner
    u, v, zw, idf = r["rast"][0, 26, 26]
    assert u > 0.98 and v < 0.02 and zw == 0.0 and idf == 1.0
    # du/dX = -1 / 204.8 per pixel
    assert np.allclose(r["rast_db"][0, 100, 60], [-1 / 204.8, -1 / 204.8, 1 / 204.8, 0.0], atol=1e-6)
    out, _ = oracle.interpolate_fwd(col, r["rast"], tri)
    assert np.allclose(out[0].sum(-1)[ids > 0], 1.0, atol=1e-6) and np.all(out[0][ids == 0] == 0)


def test_fill_rule_shared_edges_and_pixel_centres():
    pos, tri = wl.grid_aligned_scene(32)
    ids = oracle.rasterize_fwd(pos, tri, (32, 32))["tri_id"][0]
    # quad [2.5,10.5)^2: vertices exactly on pixel centres -> left/bottom inclusive, right/top exclusive
    assert np.all(ids[2:10, 2:10] > 0) and np.all(ids[2:10, 10] != ids[2:10, 9]) and np.all(ids[1, 2:10] == 0)
    assert set(np.unique(ids[2:10, 2:10])) == {1, 2} and set(np.unique(ids[2:10, 10:18])) == {3, 4}
    # every pixel inside the union is covered exactly once (no gaps on shared edges): area 8x8 per quad
    for a, b in [((2, 10), (2, 10)), ((2, 10), (10, 18)), ((10, 18), (2, 10)), ((12, 20), (12, 20)), ((12, 20), (20, 28))]:
        assert np.all(ids[a[0]:a[1], b[0]:b[1]] > 0)
    assert (ids > 0).sum() == 5 * 64 + 64       # five quads + the 8x8 fan around (24.5, 4.5)


def test_depth_rule_and_ties():
    pos, tri = wl.zfight_scene()
    r = oracle.rasterize_fwd(pos, tri, (64, 64))
    ids = r["tri_id"][0]
    assert not np.any((ids == 1) | (ids == 2))             # exact duplicates: the later copy (3,4) wins everywhere
    assert np.all(ids[30, 30:40] >= 5)                     # equal depth, partial overlap: later quad (5,6) on top
    # nearer surface wins regardless of order
    near = pos.copy(); near[0, :4, 2] = -0.5
    ids2 = oracle.rasterize_fwd(near, tri, (64, 64))["tri_id"][0]
    assert np.all(np.isin(ids2[(ids > 0) & np.isin(ids, [3, 4])], [1, 2]))


def test_peeling_layers_are_ordered():
    pos, tri = wl.soup_scene(1, 60, 3)
    prev = None
    last = None
    for layer in range(3):
        r = oracle.rasterize_fwd(pos, tri, (48, 48), peel_depth=prev)
        if last is not None:
            both = (r["tri_id"] > 0)
            assert np.all(r["depth"][both] > last[both])
        prev = last = r["depth"]


def test_clipped_triangle_matches_unclipped_interior():
    # a triangle poking out of the frustum renders the same interior pixels as the oracle's float edge test predicts
    pos = np.array([[[-3.0, -0.5, 0.0, 1.0], [0.5, -0.5, 0.0, 1.0], [0.5, 3.0, 0.0, 1.0]]], np.float32)
    tri = np.array([[0, 1, 2]], np.int32)
    ids = oracle.rasterize_fwd(pos, tri, (64, 64))["tri_id"][0]
    ys, xs = np.mgrid[0:64, 0:64]
    fx, fy = (xs + .5) / 32 - 1, (ys + .5) / 32 - 1
    inside = (fy > -0.5) & (fx < 0.5) & ((fy + 0.5) * 3.5 < (fx + 3.0) * 3.5)
    assert (ids > 0).sum() > 500 and np.mean((ids > 0) != inside) < 0.02


def _fd(f, x, eps=1e-3):
    g = np.zeros_like(x, np.float64)
    it = np.nditer(x, flags=["multi_index"])
    for _ in it:
        i = it.multi_index
        xp = x.copy(); xp[i] += eps
        xm = x.copy(); xm[i] -= eps
        g[i] = (f(xp) - f(xm)) / (2 * eps)
    return g


def test_interpolate_gradients_finite_difference():
    c = wl.config_c2(n=1, res=24)
    r = oracle.rasterize_fwd(c["pos"], c["tri"], (24, 24))
    attr = np.random.default_rng(0).normal(size=(1, 8, 3)).astype(np.float32)
    dy = np.random.default_rng(1).normal(size=(1, 24, 24, 3)).astype(np.float32)
    ga, gr, _ = oracle.interpolate_grad(attr, r["rast"], c["tri"], dy)
    fd = _fd(lambda a: float((oracle.interpolate_fwd(a.astype(np.float32), r["rast"], c["tri"])[0].astype(np.float64) * dy).sum()), attr.astype(np.float64), 1e-2)
    assert np.allclose(ga, fd, rtol=2e-3, atol=2e-3)


def test_texture_gradients_finite_difference():
    rng = np.random.default_rng(2)
    tex = rng.uniform(size=(1, 8, 8, 2)).astype(np.float32)
    uv = rng.uniform(0.1, 0.9, size=(1, 6, 6, 2)).astype(np.float32)
    bias = rng.uniform(0.2, 1.8, size=(1, 6, 6)).astype(np.float32)
    dy = rng.normal(size=(1, 6, 6, 2)).astype(np.float32)
    kw = dict(filter_mode="linear-mipmap-linear", boundary_mode="wrap")
    g = oracle.texture_grad(tex, uv, dy, None, bias, **kw)
    fd = _fd(lambda t: float((oracle.texture_fwd(t.astype(np.float32), uv, None, bias, **kw).astype(np.float64) * dy).sum()), tex.astype(np.float64), 1e-2)
    assert np.allclose(g["g_tex"], fd, rtol=2e-3, atol=2e-3)


def test_antialias_blends_only_silhouettes():
    c = wl.config_c2(n=1, res=64)
    r = oracle.rasterize_fwd(c["pos"], c["tri"], (64, 64))
    color, _ = oracle.interpolate_fwd(np.ones((1, 8, 3), np.float32), r["rast"], c["tri"])   # white cube on black
    out, items = oracle.antialias_fwd(color, r["rast"], c["pos"], c["tri"])
    changed = np.abs(out - color).max(-1)[0] > 0
    ids = r["tri_id"][0]
    sil = np.zeros_like(changed)
    sil[:, :-1] |= (ids[:, :-1] > 0) != (ids[:, 1:] > 0); sil[:, 1:] |= (ids[:, :-1] > 0) != (ids[:, 1:] > 0)
    sil[:-1, :] |= (ids[:-1, :] > 0) != (ids[1:, :] > 0); sil[1:, :] |= (ids[:-1, :] > 0) != (ids[1:, :] > 0)
    assert changed.sum() > 20 and not np.any(changed & ~sil)      # interior edges of the welded cube are never blended
    assert np.all(out >= -1e-6) and np.all(out <= 1 + 1e-6)


@pytest.mark.skipif(not GOLDEN, reason="golden vectors not generated yet")
@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p) for p in GOLDEN])
def test_oracle_matches_reference_golden(path):
    """The pin: outputs of the unmodified reference CUDA extension on a B200 (see tests/golden/make_golden.py)."""
    from golden_check import check_oracle_against_golden
    check_oracle_against_golden(np.load(path))
