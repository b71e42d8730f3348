"""The per-pixel arithmetic of the rasterize / antialias kernels (nvdiffrast_b200/csrc/pixel_math.cuh, __host__ __device__)
executed on the CPU through tests/native/pixel_math_harness.cu and compared with the oracle: the same expressions the GPU
runs, checked without a GPU.  (The kernels around them -- scheduling, reductions, memory movement -- are covered by the
-m gpu tests.)"""
import ctypes
import os
import subprocess

import numpy as np
import pytest

import oracle
import workloads as wl
from util import ROOT, assert_close, assert_grad_close

_SRC = os.path.join(ROOT, "tests", "native", "pixel_math_harness.cu")
_HDR = os.path.join(ROOT, "nvdiffrast_b200", "csrc", "pixel_math.cuh")
_SO = os.path.join(ROOT, "tests", "native", "_build", "libpmh.so")
c_f = ctypes.POINTER(ctypes.c_float); c_i = ctypes.POINTER(ctypes.c_int32); c_d = ctypes.POINTER(ctypes.c_double)


@pytest.fixture(scope="module")
def pmh():
    nvcc = "/usr/local/cuda/bin/nvcc"
    if not os.access(nvcc, os.X_OK):
        pytest.skip("nvcc not available")
    if not os.path.exists(_SO) or os.path.getmtime(_SO) < max(os.path.getmtime(_SRC), os.path.getmtime(_HDR)):
        os.makedirs(os.path.dirname(_SO), exist_ok=True)
        subprocess.check_call([nvcc, "-O2", "-shared", "-Xcompiler", "-fPIC", "-Wno-deprecated-gpu-targets",
                               "-I", os.path.dirname(_HDR), "-o", _SO, _SRC])
    lib = ctypes.CDLL(_SO)
    lib.pmh_antialias.restype = ctypes.c_int
    return lib


def _p(a, t):
    return None if a is None else a.ctypes.data_as(t)


SCENES = {
    "cube": lambda: (wl.config_c2(n=2, res=96)["pos"], wl.cube_mesh()[1], (96, 96)),
    "sphere": lambda: (wl.config_c3(n=1, res=128, tex_res=8)["pos"], wl.sphere_mesh()[1], (128, 128)),
    "soup": lambda: wl.soup_scene(2, 300, 1) + ((96, 80),),
}


@pytest.mark.parametrize("name", sorted(SCENES))
@pytest.mark.parametrize("with_db", [False, True])
def test_shader_and_raster_grad_match_oracle(pmh, name, with_db):
    pos, tri, res = SCENES[name]()
    H, W = res
    o = oracle.rasterize_fwd(pos, tri, res)
    rng = np.random.default_rng(5)
    dy = rng.normal(size=o["rast"].shape).astype(np.float32); ddb = rng.normal(size=o["rast"].shape).astype(np.float32)
    want = oracle.rasterize_grad(pos, tri, o["rast"], dy, ddb if with_db else None)
    got = np.zeros(pos.shape, np.float64)
    for n in range(pos.shape[0]):
        rast = np.ascontiguousarray(o["rast"][n]); ro = np.empty_like(rast); dbo = np.empty_like(rast)
        g = np.zeros((pos.shape[1], 4), np.float64)
        pmh.pmh_raster(_p(np.ascontiguousarray(pos[n]), c_f), _p(tri, c_i), _p(rast, c_f), _p(np.ascontiguousarray(dy[n]), c_f),
                       _p(np.ascontiguousarray(ddb[n]), c_f) if with_db else None, H, W, int(with_db), _p(ro, c_f), _p(dbo, c_f), _p(g, c_d))
        got[n] = g
        # forward shader: same values as the oracle's pixel shader (both uncontracted on the host)
        assert_close(ro[..., :3], rast[..., :3], 1e-6, 1e-5, "uvz")
        assert np.array_equal(ro[..., 3], rast[..., 3])
        assert_close(dbo, o["rast_db"][n], 1e-6, 1e-4, "rast_db")
    assert_grad_close(got, want, 1e-4, name + " g_pos")


def _wings(tri, V):
    """per triangle: the vertex opposite each edge in the neighbouring triangle (edges with a third triangle keep the first
    two, as the reference's hash does), -1 for open edges"""
    edge = {}
    for t, (a, b, c) in enumerate(tri):
        if a == b or b == c or c == a:
            continue
        for (u, v, o) in ((b, c, a), (c, a, b), (a, b, c)):
            edge.setdefault((min(u, v), max(u, v)), [])
            lst = edge[(min(u, v), max(u, v))]
            if o not in lst and len(lst) < 2:
                lst.append(o)
    w = np.full((len(tri), 3), -1, np.int32)
    for t, (a, b, c) in enumerate(tri):
        for i, (u, v, o) in enumerate(((b, c, a), (c, a, b), (a, b, c))):
            lst = edge.get((min(u, v), max(u, v)), [])
            if o in lst and len(lst) == 2:
                w[t, i] = lst[1] if lst[0] == o else lst[0]
    return w


@pytest.mark.parametrize("name", ["cube", "sphere"])
def test_antialias_edge_math_matches_oracle(pmh, name):
    pos, tri, res = SCENES[name]()
    H, W = res
    C = 3
    o = oracle.rasterize_fwd(pos, tri, res)
    rng = np.random.default_rng(14)
    color = rng.uniform(size=o["rast"].shape[:3] + (C,)).astype(np.float32)
    dy = rng.normal(size=color.shape).astype(np.float32)
    want_out, want_items = oracle.antialias_fwd(color, o["rast"], pos, tri)
    want_gc, want_gp = oracle.antialias_grad(color, o["rast"], pos, tri, dy, want_items)
    wings = _wings(tri, pos.shape[1])
    got_gp = np.zeros(pos.shape, np.float64)
    for n in range(pos.shape[0]):
        rast = np.ascontiguousarray(o["rast"][n])
        items = np.zeros((2 * H * W, 6), np.int32)
        acc = np.zeros((H, W, C), np.float64); gc = np.zeros((H, W, C), np.float64); gp = np.zeros((pos.shape[1], 4), np.float64)
        cnt = pmh.pmh_antialias(_p(rast, c_f), _p(np.ascontiguousarray(pos[n]), c_f), _p(tri, c_i), _p(wings, c_i),
                                _p(np.ascontiguousarray(color[n]), c_f), _p(np.ascontiguousarray(dy[n]), c_f), H, W, C, pos.shape[1], tri.shape[0],
                                _p(items, c_i), _p(acc, c_d), _p(gc, c_d), _p(gp, c_d))
        mine = items[:cnt]
        ref = want_items[want_items[:, 2] == n]
        assert cnt == ref.shape[0] and cnt > 50
        # same pixel pairs, same deciding edge, same owner; alpha to the last bits
        assert np.array_equal(mine[:, :2], ref[:, :2]) and np.array_equal(mine[:, 2], ref[:, 3])
        assert np.array_equal(mine[:, 3], ref[:, 4]) and np.array_equal(mine[:, 4], ref[:, 5])
        a_mine = mine[:, 5].copy().view(np.float32); a_ref = ref[:, 6].copy().view(np.float32)
        assert np.abs(a_mine - a_ref).max() <= 1e-6
        assert_close((color[n] + acc).astype(np.float32), want_out[n], 1e-6, 1e-6, "aa out")
        assert_close((dy[n] + gc).astype(np.float32), want_gc[n], 1e-6, 1e-5, "aa g_color")
        got_gp[n] = gp
    assert_grad_close(got_gp, want_gp, 1e-4, name + " aa g_pos")
