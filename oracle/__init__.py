"""TEST INFRASTRUCTURE ONLY -- numpy front end of the CPU oracle (oracle/oracle.c).

The oracle is a from-scratch CPU restatement of the nvdiffrast v0.4.0 hot path used as the parity checker.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
package; the product package (nvdiffrast_b200) never does.

Function names follow the reference's native op names (csrc/torch/torch_bindings.cpp:43-71).
"""
import ctypes
import os
import subprocess
import numpy as np

_DIR = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_DIR, "_build", "liboracle.so")
_lib = None

c_f = ctypes.POINTER(ctypes.c_float)
c_i = ctypes.POINTER(ctypes.c_int32)
c_u = ctypes.POINTER(ctypes.c_uint32)


def build(force=False):
    """Compile oracle.c with the host gcc (OpenMP when available)."""
    src = os.path.join(_DIR, "oracle.c")
    if not force and os.path.exists(_SO) and os.path.getmtime(_SO) >= os.path.getmtime(src):
        return _SO
    os.makedirs(os.path.dirname(_SO), exist_ok=True)
    cc = "/usr/bin/gcc" if os.access("/usr/bin/gcc", os.X_OK) else "gcc"
    base = [cc, "-O2", "-ffp-contract=off", "-fno-fast-math", "-fvisibility=hidden", "-shared", "-fPIC",
            "-o", _SO, src, "-lm"]
    try:
        subprocess.check_call(base[:1] + ["-fopenmp"] + base[1:], stderr=subprocess.DEVNULL)
    except Exception:
        subprocess.check_call(base)
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
        _lib.orc_mip_info.restype = ctypes.c_longlong
    return _lib


def _f(a):
    return None if a is None else a.ctypes.data_as(c_f)


def _i(a):
    return None if a is None else a.ctypes.data_as(c_i)


def _c32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float32)


def _ci32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.int32)


def num_threads():
    return int(lib().orc_num_threads())


# ----------------------------------------------------------------------------------------------
# rasterize

def rasterize_fwd(pos, tri, resolution, ranges=None, peel_depth=None):
    """-> dict(rast, rast_db, tri_id (int32 [N,H,W]), depth (uint32 [N,H,W])).  torch_rasterize.cpp:43-166."""
    pos = _c32(pos); tri = _ci32(tri)
    H, W = int(resolution[0]), int(resolution[1])
    inst = pos.ndim == 3
    if inst:
        N, V = pos.shape[0], pos.shape[1]
        rng = None
    else:
        rng = _ci32(ranges); N, V = rng.shape[0], pos.shape[0]
    T = tri.shape[0]
    out_id = np.empty((N, H, W), np.int32); out_depth = np.empty((N, H, W), np.uint32)
    rast = np.empty((N, H, W, 4), np.float32); rast_db = np.empty((N, H, W, 4), np.float32)
    peel = None if peel_depth is None else np.ascontiguousarray(peel_depth, dtype=np.uint32)
    lib().orc_rasterize_fwd(_f(pos), _i(tri), _i(rng), N, V, T, H, W, int(inst),
                            None if peel is None else peel.ctypes.data_as(c_u),
                            _i(out_id), out_depth.ctypes.data_as(c_u), _f(rast), _f(rast_db))
    return dict(rast=rast, rast_db=rast_db, tri_id=out_id, depth=out_depth)


def rasterize_grad(pos, tri, rast, dy, ddb=None):
    """-> g_pos.  torch_rasterize.cpp:171-256 (rasterize_grad / rasterize_grad_db)."""
    pos = _c32(pos); tri = _ci32(tri); rast = _c32(rast); dy = _c32(dy); ddb = _c32(ddb)
    inst = pos.ndim == 3
    N, H, W = rast.shape[:3]
    V = pos.shape[1] if inst else pos.shape[0]
    g = np.empty_like(pos)
    lib().orc_rasterize_grad(_f(pos), _i(tri), _f(rast), _f(dy), _f(ddb), N, V, tri.shape[0], H, W, int(inst), _f(g))
    return g


# ----------------------------------------------------------------------------------------------
# interpolate

def _attr_dims(attr, rast):
    inst = attr.ndim > 2
    N = rast.shape[0]
    if inst:
        V, A = attr.shape[1], attr.shape[2]
        bc = int(attr.shape[0] == 1 and N > 1)
    else:
        V, A = attr.shape
        bc = 0
    return inst, V, A, bc


def _diff(diff_attrs, A):
    if diff_attrs is None or (not isinstance(diff_attrs, str) and len(diff_attrs) == 0):
        return 0, None, 0
    if isinstance(diff_attrs, str) and diff_attrs == 'all':
        return 1, None, A
    lst = np.asarray(diff_attrs, np.int32)
    return 0, lst, int(lst.shape[0])


def interpolate_fwd(attr, rast, tri, rast_db=None, diff_attrs=None):
    """-> (out, out_da).  torch_interpolate.cpp:42-132."""
    attr = _c32(attr); rast = _c32(rast); tri = _ci32(tri); rast_db = _c32(rast_db)
    inst, V, A, bc = _attr_dims(attr, rast)
    N, H, W = rast.shape[:3]
    dall, lst, D = _diff(diff_attrs, A)
    out = np.empty((N, H, W, A), np.float32); out_da = np.empty((N, H, W, 2 * D), np.float32)
    lib().orc_interpolate_fwd(_f(attr), _f(rast), _i(tri), _f(rast_db), N, H, W, V, tri.shape[0], A, bc, int(inst),
                              dall, _i(lst), D, _f(out), _f(out_da) if D else None)
    return out, out_da


def interpolate_grad(attr, rast, tri, dy, rast_db=None, dda=None, diff_attrs=None):
    """-> (g_attr, g_rast, g_rast_db or None).  torch_interpolate.cpp:137-248."""
    attr = _c32(attr); rast = _c32(rast); tri = _ci32(tri); dy = _c32(dy); rast_db = _c32(rast_db); dda = _c32(dda)
    inst, V, A, bc = _attr_dims(attr, rast)
    N, H, W = rast.shape[:3]
    dall, lst, D = _diff(diff_attrs, A)
    g_attr = np.empty_like(attr); g_rast = np.empty((N, H, W, 4), np.float32)
    g_db = np.empty((N, H, W, 4), np.float32) if D else None
    lib().orc_interpolate_grad(_f(attr), _f(rast), _i(tri), _f(dy), _f(rast_db), _f(dda), N, H, W, V, tri.shape[0], A, bc,
                               int(inst), dall, _i(lst), D, _f(g_attr), _f(g_rast), _f(g_db))
    return g_attr, g_rast, g_db


# ----------------------------------------------------------------------------------------------
# antialias

def antialias_fwd(color, rast, pos, tri):
    """-> (out, items) ; items is the oracle's own work list consumed by antialias_grad.  torch_antialias.cpp:68-155."""
    color = _c32(color); rast = _c32(rast); pos = _c32(pos); tri = _ci32(tri)
    N, H, W, C = color.shape
    inst = pos.ndim > 2
    V = pos.shape[1] if inst else pos.shape[0]
    out = np.empty_like(color)
    items = np.zeros((2 * N * H * W, 7), np.int32)
    n = ctypes.c_int64(0)
    lib().orc_antialias_fwd(_f(color), _f(rast), _f(pos), _i(tri), N, H, W, C, V, tri.shape[0], int(inst),
                            _f(out), _i(items), ctypes.byref(n))
    return out, items[:n.value].copy()


def antialias_knife_pixels(rast, pos, tri, tau=1e-3):
    """Test support: bool [N,H,W] mask of the pixels that belong to a pixel pair whose antialias analysis sits within `tau` of
    flipping one of its discrete decisions (silhouette sign tests, axis crossings, candidate choice, |dy| >= |dx|, acceptance
    window, clamps; see aa_margin in oracle.c).  Only there can two correct fp32 implementations disagree by a finite amount."""
    rast = _c32(rast); pos = _c32(pos); tri = _ci32(tri)
    N, H, W = rast.shape[:3]
    inst = pos.ndim > 2
    V = pos.shape[1] if inst else pos.shape[0]
    m = np.empty((N, H, W, 2), np.float32)
    lib().orc_antialias_margins(_f(rast), _f(pos), _i(tri), N, H, W, V, tri.shape[0], int(inst), _f(m))
    k = m < tau
    mask = k[..., 0] | k[..., 1]
    mask[:, :, 1:] |= k[:, :, :-1, 0]            # right partner of a horizontal pair
    mask[:, 1:, :] |= k[:, :-1, :, 1]            # upper partner of a vertical pair
    return mask


def antialias_grad(color, rast, pos, tri, dy, items):
    """-> (g_color, g_pos).  torch_antialias.cpp:160-241."""
    color = _c32(color); rast = _c32(rast); pos = _c32(pos); tri = _ci32(tri); dy = _c32(dy); items = _ci32(items)
    N, H, W, C = color.shape
    inst = pos.ndim > 2
    V = pos.shape[1] if inst else pos.shape[0]
    g_color = np.empty_like(color); g_pos = np.empty_like(pos)
    lib().orc_antialias_grad(_f(color), _f(rast), _f(pos), _i(tri), _f(dy), _i(items), ctypes.c_int64(items.shape[0]),
                             N, H, W, C, V, tri.shape[0], int(inst), _f(g_color), _f(g_pos))
    return g_color, g_pos


# ----------------------------------------------------------------------------------------------
# texture (2-D)

FILTER = {'nearest': 0, 'linear': 1, 'linear-mipmap-nearest': 2, 'linear-mipmap-linear': 3}
BOUNDARY = {'cube': 0, 'wrap': 1, 'clamp': 2, 'zero': 3}


def mip_info(tex_shape, max_mip_level=-1):
    """-> (total floats, offsets[17], max_level).  texture.cpp:62-102."""
    D, Ht, Wt, C = tex_shape
    offs = (ctypes.c_int * 17)()
    ml = ctypes.c_int(0)
    tot = lib().orc_mip_info(Wt, Ht, D, C, int(max_mip_level), offs, ctypes.byref(ml))
    if tot < 0:
        raise RuntimeError("texture size is not compatible with mipmapping")
    return int(tot), np.array(list(offs), np.int32), int(ml.value)


def _as_slices(tex):
    """cube maps [D,6,H,W,C] are handled as 6*D two-dimensional slices"""
    return tex.reshape((-1,) + tex.shape[-3:]) if tex.ndim == 5 else tex


def texture_construct_mip(tex, max_mip_level=-1):
    """-> dict(mip, offsets, max_level).  torch_texture.cpp:98-169."""
    tex = _as_slices(_c32(tex))
    tot, offs, ml = mip_info(tex.shape, max_mip_level)
    mip = np.zeros(max(tot, 1), np.float32)
    D, Ht, Wt, C = tex.shape
    lib().orc_mip_build(_f(tex), _f(mip), Wt, Ht, D, C, _i(offs), ml)
    return dict(mip=mip, offsets=offs, max_level=ml, total=tot)


def _mipargs(tex, filter_mode, mip, max_mip_level):
    if 'mipmap' in filter_mode:
        if mip is None:
            mip = texture_construct_mip(tex, max_mip_level)
        return mip['mip'], mip['offsets'], mip['max_level'], mip['total']
    return np.zeros(1, np.float32), np.zeros(17, np.int32), 0, 0


def texture_fwd(tex, uv, uv_da=None, mip_level_bias=None, mip=None, filter_mode='linear', boundary_mode='wrap', max_mip_level=-1):
    """-> out [N,H,W,C].  torch_texture.cpp:174-416 (2-D modes)."""
    tex = _c32(tex); uv = _c32(uv); uv_da = _c32(uv_da); bias = _c32(mip_level_bias)
    m, offs, ml, _ = _mipargs(tex, filter_mode, mip, max_mip_level)
    N, H, W = uv.shape[:3]
    D, Ht, Wt, C = tex.shape[0], tex.shape[-3], tex.shape[-2], tex.shape[-1]
    out = np.empty((N, H, W, C), np.float32)
    lib().orc_texture_fwd(_f(tex), _f(m), _i(offs), ml, _f(uv), _f(uv_da), _f(bias), N, H, W, D, Ht, Wt, C,
                          FILTER[filter_mode], BOUNDARY[boundary_mode], _f(out))
    return out


def texture_grad(tex, uv, dy, uv_da=None, mip_level_bias=None, mip=None, filter_mode='linear', boundary_mode='wrap', max_mip_level=-1):
    """-> dict(g_tex, g_uv, g_uv_da, g_bias).  torch_texture.cpp:421-716 (2-D modes, internal mip stack)."""
    tex = _c32(tex); uv = _c32(uv); dy = _c32(dy); uv_da = _c32(uv_da); bias = _c32(mip_level_bias)
    m, offs, ml, tot = _mipargs(tex, filter_mode, mip, max_mip_level)
    N, H, W = uv.shape[:3]
    D, Ht, Wt, C = tex.shape[0], tex.shape[-3], tex.shape[-2], tex.shape[-1]
    g_tex = np.empty_like(tex)
    g_uv = np.zeros(uv.shape, np.float32) if filter_mode != 'nearest' else None
    trilinear = filter_mode == 'linear-mipmap-linear'
    g_uv_da = np.zeros(uv_da.shape, np.float32) if (trilinear and uv_da is not None) else None
    g_bias = np.zeros((N, H, W), np.float32) if (trilinear and bias is not None) else None
    lib().orc_texture_grad(_f(tex), _f(m), _i(offs), ml, ctypes.c_longlong(tot), _f(uv), _f(uv_da), _f(bias), _f(dy),
                           N, H, W, D, Ht, Wt, C, FILTER[filter_mode], BOUNDARY[boundary_mode],
                           _f(g_tex), None, _f(g_uv), _f(g_uv_da), _f(g_bias))
    return dict(g_tex=g_tex, g_uv=g_uv, g_uv_da=g_uv_da, g_bias=g_bias)
