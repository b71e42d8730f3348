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


# ----------------------------------------------------------------------------------------------
# Knife-edge predicates and condition-scaled tolerances (no anonymous outlier budgets: every pixel is either held to the
# stated tolerance or removed by one of these stated predicates)

EPS32 = 2.0 ** -24


def raster_condition(pos, tri, rast, res):
    """Condition numbers of the shaded values of every covered pixel, from a float64 evaluation of the edge functions
    (units of 2^-24: multiply by a small constant k for the tolerance).
      barycentrics: u, v = e_i / sum(e); each e_i = q_j x q_k carries an absolute rounding error of about 2^-24 m_i, whatever
                    the operation order, with m_i = |qx_j qy_k| + |qy_j qx_k| + (rounding of q = x - fx w, which cancels O(1)
                    clip coordinates down to a few pixel widths) -> cond_uv = sum(m) / |sum(e)|: small and sliver triangles
                    have u, v accurate to 1e-5..1e-4 only, in ANY fp32 evaluation of the reference formula;
      depth:        z/w = sum(z_i e_i) / sum(w_i e_i) -> cond_zw = (sum |z_i| m_i + |z/w| sum |w_i| m_i) / |sum(w_i e_i)|;
      rast_db:      d = k (u sn - n_0) with k = (2/W) / sum(e), n_i = y_k w_j - y_j w_k (error 2^-24 mn_i):
                    cond_db = |k| (cond_uv |sn| + |u| sum(mn) + mn_0 + |u sn| + |n_0|) + cond_uv |d|   (same in y and for v).
    Returns (cond [N,H,W,3] for (u, v, z/w), cond_db [N,H,W,4], covered mask)."""
    pos = np.asarray(pos, np.float64); tri = np.asarray(tri)
    H, W = res
    rast = np.asarray(rast, np.float64)
    ids = rast[..., 3].astype(np.int64)
    N = ids.shape[0]
    cond = np.zeros(ids.shape + (3,), np.float64)
    cond_db = np.zeros(ids.shape + (4,), np.float64)
    fx = (2.0 / W) * np.arange(W) + (1.0 / W - 1.0)
    fy = (2.0 / H) * np.arange(H) + (1.0 / H - 1.0)
    for n in range(N):
        ys, xs = np.nonzero(ids[n] > 0)
        t = tri[ids[n, ys, xs] - 1]
        vb = pos[n] if pos.ndim == 3 else pos
        p = vb[t]                                               # [K,3,4]
        X, Y, Z, Wc = p[:, :, 0], p[:, :, 1], p[:, :, 2], p[:, :, 3]
        qx = X - fx[xs, None] * Wc; qy = Y - fy[ys, None] * Wc
        dqx = np.abs(X) + np.abs(fx[xs, None] * Wc); dqy = np.abs(Y) + np.abs(fy[ys, None] * Wc)
        K = len(ys)
        m = np.zeros((K, 3)); e = np.zeros((K, 3)); nx = np.zeros((K, 3)); ny = np.zeros((K, 3)); mnx = np.zeros((K, 3)); mny = np.zeros((K, 3))
        for i in range(3):
            j, k = (i + 1) % 3, (i + 2) % 3
            # rounding of the products, plus the rounding of q itself: q = x - fx w cancels O(1) clip coordinates down to a few
            # pixel widths, so its absolute error 2^-24 (|x| + |fx w|) is LARGE relative to q (this term dominates)
            m[:, i] = (np.abs(qx[:, j] * qy[:, k]) + np.abs(qy[:, j] * qx[:, k])
                       + dqx[:, j] * np.abs(qy[:, k]) + np.abs(qx[:, j]) * dqy[:, k] + dqy[:, j] * np.abs(qx[:, k]) + np.abs(qy[:, j]) * dqx[:, k])
            e[:, i] = qx[:, j] * qy[:, k] - qy[:, j] * qx[:, k]
            nx[:, i] = Y[:, k] * Wc[:, j] - Y[:, j] * Wc[:, k]; mnx[:, i] = np.abs(Y[:, k] * Wc[:, j]) + np.abs(Y[:, j] * Wc[:, k])
            ny[:, i] = X[:, j] * Wc[:, k] - X[:, k] * Wc[:, j]; mny[:, i] = np.abs(X[:, j] * Wc[:, k]) + np.abs(X[:, k] * Wc[:, j])
        S = np.maximum(np.abs(e.sum(1)), 1e-300)
        wn = (Wc * e).sum(1); zn = (Z * e).sum(1)
        awn = np.maximum(np.abs(wn), 1e-300)
        zw = np.clip(zn / np.where(wn == 0, 1e-300, wn), -1.0, 1.0)
        cuv = m.sum(1) / S
        czw = ((np.abs(Z) * m).sum(1) + np.abs(zw) * (np.abs(Wc) * m).sum(1)) / awn
        cond[n, ys, xs, 0] = cuv; cond[n, ys, xs, 1] = cuv; cond[n, ys, xs, 2] = czw
        uvv = rast[n, ys, xs, :2]
        for c, (nn, mn, sc) in enumerate(((nx, mnx, 2.0 / W), (ny, mny, 2.0 / H))):
            kk = sc / S
            sn = nn.sum(1)
            for b in range(2):                                  # b = 0: u (vertex 0), b = 1: v (vertex 1)
                d = kk * np.abs(uvv[:, b] * sn - nn[:, b])
                cond_db[n, ys, xs, 2 * b + c] = kk * (cuv * np.abs(sn) + np.abs(uvv[:, b]) * mn.sum(1) + mn[:, b] + np.abs(uvv[:, b] * sn) + np.abs(nn[:, b])) + cuv * d
    return cond, cond_db, ids > 0


def assert_close_cond(a, b, cond, atol, rtol, k, name=""):
    """|a - b| <= atol + rtol |b| + k 2^-24 cond  for every element (cond per element, or per pixel and broadcast)."""
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    c = cond if cond.ndim == a.ndim else cond[..., None]
    tol = atol + rtol * np.abs(b) + k * EPS32 * c
    bad = np.abs(a - b) > tol
    if bad.any():
        i = np.unravel_index(np.argmax(np.abs(a - b) - tol), a.shape)
        raise AssertionError("%s: %d elements outside atol=%g rtol=%g k=%g; worst at %s: got %.9g want %.9g tol %.3g cond %.3g (needs k >= %.1f)" % (
            name, int(bad.sum()), atol, rtol, k, i, a[i], b[i], tol[i], np.broadcast_to(c, a.shape)[i],
            float(((np.abs(a - b) - atol - rtol * np.abs(b)) / (EPS32 * np.maximum(np.broadcast_to(c, a.shape), 1e-300)))[bad].max())))


def mip_level_f64(uv_da, bias, tex_w, tex_h, max_level):
    """Continuous (unclamped) mip level of every pixel in float64 (texture_kernel.cu:509-568), for knife-edge predicates."""
    shape = (uv_da if uv_da is not None else bias).shape[:3]
    lvl = np.zeros(shape, np.float64)
    if uv_da is not None:
        q = np.asarray(uv_da, np.float64)
        dsdx, dsdy, dtdx, dtdy = q[..., 0] * tex_w, q[..., 1] * tex_w, q[..., 2] * tex_h, q[..., 3] * tex_h
        A = dsdx ** 2 + dtdx ** 2; B = dsdy ** 2 + dtdy ** 2; C = dsdx * dsdy + dtdx * dtdy
        lam = 0.5 * (A + B) + np.sqrt(0.25 * (A - B) ** 2 + C ** 2)
        with np.errstate(divide="ignore"):
            lvl = 0.5 * np.log2(np.maximum(lam, 1e-300))
    if bias is not None:
        lvl = lvl + np.asarray(bias, np.float64)
    return lvl


def texture_knife_edges(uv, uv_da, bias, tex_shape, filter_mode, boundary_mode, max_level, level_eps=1e-3, texel_eps=2e-4):
    """Pixels that sit on a discontinuity of the texture op, where a last-bit difference upstream legitimately changes the
    result by a finite amount:
      level: continuous mip level within `level_eps` of an integer inside [0, max_level], or of the clamp ends (the GPU
             evaluates it with the approximate __log2f); relevant to every output of the mipmapped modes;
      texel: u*w - 0.5 (or v*h - 0.5) within `texel_eps` of -- but not exactly on -- an integer on a level the pixel uses (an
             exact hit is computed exactly by both implementations: fp32 products are correctly rounded): the bilinear footprint
             moves to the next texel quad there; the colour is continuous but d(colour)/d(uv) is not, relevant to g_uv,
             g_uv_da, and -- through which texels receive the (vanishing) weight -- irrelevant to g_tex.
    Returns (level_mask, texel_mask), each [N,H,W] bool."""
    D, th, tw = tex_shape[0], tex_shape[-3], tex_shape[-2]
    uv = np.asarray(uv, np.float64)
    shape = uv.shape[:3]
    mip = "mipmap" in filter_mode
    lvl = mip_level_f64(uv_da, bias, tw, th, max_level) if mip else np.zeros(shape)
    level_mask = np.zeros(shape, bool)
    if mip:
        inside = (lvl > -level_eps) & (lvl < max_level + level_eps)
        level_mask = inside & (np.abs(lvl - np.round(lvl)) < level_eps)
    texel_mask = np.zeros(shape, bool)
    if filter_mode != "nearest":
        lc = np.clip(lvl, 0, max_level)
        for L in sorted(set(np.floor(lc).astype(int).ravel().tolist()) | set(np.minimum(np.floor(lc).astype(int) + 1, max_level).ravel().tolist())):
            uses = (np.floor(lc).astype(int) == L) | ((np.minimum(np.floor(lc).astype(int) + 1, max_level) == L) & (lc > np.floor(lc)) & (filter_mode == "linear-mipmap-linear"))
            w, h = max(tw >> L, 1), max(th >> L, 1)
            for c, s in ((0, w), (1, h)):
                x = uv[..., c]
                if boundary_mode == "wrap":
                    x = x - np.floor(x)
                x = x * s - 0.5
                d = np.abs(x - np.round(x))
                texel_mask |= uses & (d > 0) & (d < texel_eps * max(1.0, s / 64.0))
    else:
        for c, s in ((0, tw), (1, th)):
            x = uv[..., c]
            if boundary_mode == "wrap":
                x = x - np.floor(x)
            x = x * s
            d = np.abs(x - np.round(x))
            texel_mask |= (d > 0) & (d < texel_eps)
    return level_mask, texel_mask
