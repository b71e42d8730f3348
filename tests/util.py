"""Helpers shared by the parity tests."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(ROOT, "oracle", "_ref")
# Reference sources with the header's own coalesced-atomics switch forced to its plain-atomicAdd branch
# (oracle/build_ref.sh): the only reference-computed values of raster g_pos / interpolate g_attr / linear-mode g_tex /
# antialias g_pos that are not corrupted by the caAtomicAdd race on sm_100a.  Both builds define the pybind module
# `_nvdiffrast_c` (and register the same C++ types), so a process can load only ONE of them: parity tests import the
# unmodified build in the pytest process and drive the race-free build in a child process (tests/ref_noca_compare.py).
REF_NOCA_DIR = os.path.join(ROOT, "oracle", "_ref_noca")
_loaded_variant = None


def reference_available(variant="stock"):
    d = REF_DIR if variant == "stock" else REF_NOCA_DIR
    return os.path.isdir(os.path.join(d, "nvdiffrast"))


def load_reference(variant="stock"):
    """Import the reference built by oracle/build_ref.sh: variant 'stock' = unmodified (oracle/_ref), 'noca' = plain-atomics
    gradient build (oracle/_ref_noca).  None if absent/unloadable."""
    global _loaded_variant
    d = REF_DIR if variant == "stock" else REF_NOCA_DIR
    if not os.path.isdir(os.path.join(d, "nvdiffrast")):
        return None
    assert _loaded_variant in (None, variant), "only one reference build per process (already loaded: %s)" % _loaded_variant
    if d not in sys.path:
        sys.path.insert(0, d)
    try:
        import nvdiffrast.torch as ref_dr
        _loaded_variant = variant
        return ref_dr
    except Exception as e:  # e.g. ABI mismatch on another image
        print("reference not loadable:", e)
        return None


def close_stats(a, b, atol, rtol):
    """-> (max abs err, fraction of elements violating |a-b| <= atol + rtol*|b|)."""
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    err = np.abs(a - b)
    bad = err > (atol + rtol * np.abs(b))
    return (float(err.max()) if err.size else 0.0), (float(bad.mean()) if err.size else 0.0)


def assert_close(a, b, atol, rtol, name="", max_bad_frac=0.0, hard_atol=None):
    mx, frac = close_stats(a, b, atol, rtol)
    assert frac <= max_bad_frac, "%s: %.3g of elements outside atol=%g rtol=%g (max abs err %.3g)" % (name, frac, atol, rtol, mx)
    if hard_atol is not None:
        assert mx <= hard_atol, "%s: max abs err %.3g > %g" % (name, mx, hard_atol)


def assert_grad_close(g, g_ref, rel=1e-4, name=""):
    """|g - g_ref|_inf <= rel * |g_ref|_inf  (atomic summation order differs between implementations)."""
    g = np.asarray(g, np.float64); g_ref = np.asarray(g_ref, np.float64)
    scale = max(float(np.abs(g_ref).max()), 1e-30)
    err = float(np.abs(g - g_ref).max())
    assert err <= rel * scale, "%s: max err %.3g vs scale %.3g (rel %.3g > %g)" % (name, err, scale, err / scale, rel)


def oracle_chain(c, n, with_tex, dy, max_mip_level=-1):
    """Whole pipeline on the CPU oracle: rasterize -> interpolate -> (texture) -> antialias, forward and backward.
    Returns (rast, color, out, grads dict) with the same leaves as the torch pipelines in the tests / bench."""
    import oracle
    pos, tri = c["pos"][:n], c["tri"]
    r = oracle.rasterize_fwd(pos, tri, c["res"])
    rast, db = r["rast"], r["rast_db"]
    if with_tex:
        uv = c["uv"][None]
        texc, texd = oracle.interpolate_fwd(uv, rast, c["uv_idx"], db, "all")
        mip = oracle.texture_construct_mip(c["tex"], max_mip_level)
        color = oracle.texture_fwd(c["tex"], texc, texd, None, mip=mip, filter_mode="linear-mipmap-linear")
    else:
        attr = c["attr"][None]
        color, _ = oracle.interpolate_fwd(attr, rast, tri)
    out, items = oracle.antialias_fwd(color, rast, pos, tri)
    g_color, g_pos = oracle.antialias_grad(color, rast, pos, tri, dy, items)
    grads = {}
    if with_tex:
        tg = oracle.texture_grad(c["tex"], texc, g_color, texd, None, mip=mip, filter_mode="linear-mipmap-linear")
        grads["tex"] = tg["g_tex"]
        g_uv, g_rast, g_db = oracle.interpolate_grad(uv, rast, c["uv_idx"], tg["g_uv"], db, tg["g_uv_da"], "all")
        grads["uv"] = g_uv[0]
        g_pos = g_pos + oracle.rasterize_grad(pos, tri, rast, g_rast, g_db)
    else:
        g_attr, g_rast, _ = oracle.interpolate_grad(attr, rast, tri, g_color)
        grads["attr"] = g_attr[0]
        g_pos = g_pos + oracle.rasterize_grad(pos, tri, rast, g_rast, None)
    grads["pos"] = g_pos
    return rast, color, out, grads
