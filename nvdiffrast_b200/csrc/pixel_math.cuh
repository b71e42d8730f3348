// Per-pixel / per-edge arithmetic of the rasterize and antialias ops, written once as __host__ __device__ code so that the
// exact expressions the kernels execute can also be run on the CPU by tests/test_pixel_math_cpu.py (against the oracle).
//
// Everything here is derived from the geometry, in array form over the three vertices:
//
//   A pixel centre (fx, fy) in NDC and a clip-space vertex (x, y, z, w) give the homogeneous 2-D offset
//       q = (x - fx w, y - fy w).
//   The three edge functions e_i = q_j x q_k  (j = i+1, k = i+2 mod 3, 2-D cross product) are w-scaled areas; e_i / sum(e)
//   are the perspective-correct barycentrics of the pixel centre (what csrc/common/rasterize.cu:63-94 evaluates).
//   d e_i / d fx = y_j w_k - y_k w_j and d e_i / d fy = x_k w_j - x_j w_k do not depend on the pixel (the fx, fy terms
//   cancel), which gives the screen-space derivatives of the barycentrics in closed form (rast_db).
//
// Forward shader: operation order is kept compatible with the reference build's FMA contraction so that u, v, z/w and
// rast_db agree to the last bits (they feed bit-exact consumers downstream).  Gradients use their own factorisation:
// back-propagation through (u, v, du/dX, ...) -> e_i -> q -> (x, y, w); only the regularisation constants and the clamping
// rules are the reference's (Appendix A of SURVEY.md).
#pragma once
#include <cuda_runtime.h>
#include <math.h>

#define PM_HD __host__ __device__ __forceinline__

namespace pm {

PM_HD float sat01(float v)
{
#ifdef __CUDA_ARCH__
    return __saturatef(v);
#else
    return v != v ? 0.f : fminf(fmaxf(v, 0.f), 1.f);
#endif
}
PM_HD bool sign_bits_agree(float a, float b)
{
#ifdef __CUDA_ARCH__
    return (__float_as_int(a) ^ __float_as_int(b)) >= 0;
#else
    return (signbit(a) != 0) == (signbit(b) != 0);
#endif
}
PM_HD float cross2(float ax, float ay, float bx, float by) { return ax * by - ay * bx; }
// element i of a 3-array by selects (keeps the array in registers: no dynamically indexed local memory)
template <class T> PM_HD T pick3(const T (&a)[3], int i) { return i == 0 ? a[0] : (i == 1 ? a[1] : a[2]); }

// ------------------------------------------------------------------------------------------------
// Triangle as seen from one pixel centre.

struct PixelTri
{
    float x[3], y[3], z[3], w[3];   // clip-space vertices
    float qx[3], qy[3];             // homogeneous offsets from the pixel centre
    float e[3];                     // edge functions; e[i] belongs to vertex i
    float esum;
};

PM_HD void pixel_tri(PixelTri& t, float4 v0, float4 v1, float4 v2, float fx, float fy)
{
    t.x[0] = v0.x; t.y[0] = v0.y; t.z[0] = v0.z; t.w[0] = v0.w;
    t.x[1] = v1.x; t.y[1] = v1.y; t.z[1] = v1.z; t.w[1] = v1.w;
    t.x[2] = v2.x; t.y[2] = v2.y; t.z[2] = v2.z; t.w[2] = v2.w;
    #pragma unroll
    for (int i = 0; i < 3; i++) { t.qx[i] = t.x[i] - fx * t.w[i]; t.qy[i] = t.y[i] - fy * t.w[i]; }
    #pragma unroll
    for (int i = 0; i < 3; i++)
    {
        const int j = (i + 1) % 3, k = (i + 2) % 3;
        t.e[i] = cross2(t.qx[j], t.qy[j], t.qx[k], t.qy[k]);
    }
    t.esum = t.e[0] + t.e[1] + t.e[2];
}

// Pixel-independent derivatives of the edge functions w.r.t. the pixel centre, NEGATED (the forward shader's operand
// order): nx[i] = -d e_i / d fx,  ny[i] = -d e_i / d fy.
PM_HD void edge_slopes_neg(const PixelTri& t, float (&nx)[3], float (&ny)[3])
{
    #pragma unroll
    for (int i = 0; i < 3; i++)
    {
        const int j = (i + 1) % 3, k = (i + 2) % 3;
        nx[i] = t.y[k] * t.w[j] - t.y[j] * t.w[k];
        ny[i] = t.x[j] * t.w[k] - t.x[k] * t.w[j];
    }
}

// rast = (u, v, z/w, id) with u, v clamped to the triangle, rast_db = d(u, v) / d(pixel X, Y).   xs = 2/W, ys = 2/H.
PM_HD void shade(const PixelTri& t, float xs, float ys, float id_as_float, float4& rast, float4& rast_db)
{
    const float inv = 1.f / t.esum;
    float u = t.e[0] * inv;
    float v = t.e[1] * inv;
    const float zn = t.z[0] * t.e[0] + t.z[1] * t.e[1] + t.z[2] * t.e[2];
    const float wn = t.w[0] * t.e[0] + t.w[1] * t.e[1] + t.w[2] * t.e[2];
    float depth = zn / wn;

    u = sat01(u);
    v = sat01(v);
    const float shrink = 1.f / fmaxf(u + v, 1.f);       // pull (u, v) back onto the edge u + v = 1
    u *= shrink;
    v *= shrink;
    depth = fmaxf(fminf(depth, 1.f), -1.f);
    rast = make_float4(u, v, depth, id_as_float);

    float nx[3], ny[3];
    edge_slopes_neg(t, nx, ny);
    const float snx = nx[0] + nx[1] + nx[2], sny = ny[0] + ny[1] + ny[2];
    const float kx = xs * inv, ky = ys * inv;
    rast_db = make_float4(kx * (u * snx - nx[0]), ky * (u * sny - ny[0]), kx * (v * snx - nx[1]), ky * (v * sny - ny[1]));
}

// Gradient of one pixel w.r.t. the clip-space (x, y, w) of its three vertices (z receives none).
//   in : gu, gv = dL/d(u, v); h = dL/d(du/dX, du/dY, dv/dX, dv/dY) (used when WITH_DB && has_h)
//   out: gx, gy, gw per vertex
// Unclamped barycentrics u = e0 s, v = e1 s with s = 1 / (sum e + copysign(1e-6, sum e)).
//   dL/de_i from (u, v):      E_i = s (g_i - (gu u + gv v)),  g = (gu, gv, 0)
//   screen derivatives:       F = s P,  P = sum_i (ax_i cx_i + ay_i cy_i)  with c = d e / d(fx, fy),
//                             ax_i = kx0 (d_i0 - u) + kx1 (d_i1 - v)  (kx0 = h.x xs, kx1 = h.z xs; same in y with h.y, h.w)
//                             -> E_i += -s^2 (P + Mu (d_i0 - u) + Mv (d_i1 - v)),  Mu = kx0 Scx + ky0 Scy,  Mv = kx1 Scx + ky1 Scy
//                             and directly through c: cx_i = y_j w_k - y_k w_j,  cy_i = x_k w_j - x_j w_k
//   e_i = q_j x q_k:          dL/dq_m = (E_a q_b.y - E_b q_a.y,  E_b q_a.x - E_a q_b.x),  a = m-1, b = m+1
//   q = (x - fx w, y - fy w): dL/dw_m = -fx dL/dq_m.x - fy dL/dq_m.y
template <bool WITH_DB>
PM_HD void raster_grad(const PixelTri& t, float fx, float fy, float xs, float ys, float gu, float gv, float4 h, bool has_h,
                       float (&gx)[3], float (&gy)[3], float (&gw)[3])
{
    const float s = 1.f / (t.esum + copysignf(1e-6f, t.esum));
    const float u = t.e[0] * s, v = t.e[1] * s;
    const float mix = gu * u + gv * v;
    float E[3] = { (gu - mix) * s, (gv - mix) * s, -mix * s };
    #pragma unroll
    for (int m = 0; m < 3; m++) { gx[m] = 0.f; gy[m] = 0.f; gw[m] = 0.f; }

    if (WITH_DB && has_h)
    {
        float cx[3], cy[3];
        #pragma unroll
        for (int i = 0; i < 3; i++)
        {
            const int j = (i + 1) % 3, k = (i + 2) % 3;
            cx[i] = t.y[j] * t.w[k] - t.y[k] * t.w[j];
            cy[i] = t.x[k] * t.w[j] - t.x[j] * t.w[k];
        }
        const float scx = cx[0] + cx[1] + cx[2], scy = cy[0] + cy[1] + cy[2];
        const float kx0 = h.x * xs, ky0 = h.y * ys, kx1 = h.z * xs, ky1 = h.w * ys;
        const float du[3] = { 1.f - u, -u, -u }, dv[3] = { -v, 1.f - v, -v };     // d_i0 - u, d_i1 - v
        float ax[3], ay[3], P = 0.f;
        #pragma unroll
        for (int i = 0; i < 3; i++)
        {
            ax[i] = kx0 * du[i] + kx1 * dv[i];
            ay[i] = ky0 * du[i] + ky1 * dv[i];
            P += ax[i] * cx[i] + ay[i] * cy[i];
        }
        const float Mu = kx0 * scx + ky0 * scy, Mv = kx1 * scx + ky1 * scy;
        const float s2 = s * s;
        #pragma unroll
        for (int i = 0; i < 3; i++) E[i] -= s2 * (P + Mu * du[i] + Mv * dv[i]);
        #pragma unroll
        for (int i = 0; i < 3; i++)
        {
            const int j = (i + 1) % 3, k = (i + 2) % 3;
            const float bx = s * ax[i], by = s * ay[i];
            gy[j] += bx * t.w[k]; gy[k] -= bx * t.w[j];
            gx[k] += by * t.w[j]; gx[j] -= by * t.w[k];
            gw[k] += bx * t.y[j] - by * t.x[j];
            gw[j] += by * t.x[k] - bx * t.y[k];
        }
    }
    #pragma unroll
    for (int m = 0; m < 3; m++)
    {
        const int a = (m + 2) % 3, b = (m + 1) % 3;
        const float dqx = E[a] * t.qy[b] - E[b] * t.qy[a];
        const float dqy = E[b] * t.qx[a] - E[a] * t.qx[b];
        gx[m] += dqx;
        gy[m] += dqy;
        gw[m] -= fx * dqx + fy * dqy;
    }
}

// ------------------------------------------------------------------------------------------------
// Antialias: silhouette-edge coverage of a pixel pair (csrc/common/antialias.cu:219-382 computes the same quantity).
//
// The pair is (p, p + step) with step = +x or +y; "hx, hy" = half the image size in pixels.  All geometry is taken in pixel
// units relative to the centre of the pixel that owns the nearer surface; for a vertical pair the axes are exchanged, so
// that the partner pixel always lies on the local +x or -x axis at distance 1.
//   * an edge is a silhouette edge if the triangle across it (its "wing" vertex) lies on the same side as the triangle
//     itself, or if there is no wing (the wing then defaults to the triangle's own vertex);
//   * every edge that crosses the local x axis does so at  xc_i = x_j - y_j (x_k - x_j) / (y_k - y_j); the crossing furthest
//     towards the partner pixel decides, candidates are compared as fractions (no division);
//   * coverage c = dir * xc in [0, 1] (accepted from -1/16 to 1 + 1/16, then clamped) -> blend weight alpha = dir (1/2 - c).

struct Frac { float num, den; };
// a > b for fractions whose denominators may be negative
PM_HD bool frac_greater(Frac a, Frac b)
{
    const bool straight = sign_bits_agree(a.den, b.den);
    return (a.num * b.den > b.num * a.den) == straight;
}

struct AAEdgeResult { int edge; float alpha; bool blend; };

// v*: triangle vertices, o*: wing vertex opposite each vertex (= the vertex itself when the edge is open);
// px, py: integer pixel that owns the triangle; vertical: pair is (p, p + y); toward_plus: partner is at +step (owner is the pair's first pixel)
PM_HD AAEdgeResult aa_edge(const float4 (&vtx)[3], const float4 (&wing)[3], int px, int py, float hx, float hy, bool vertical, bool toward_plus)
{
    const float cxp = (float)px + .5f - hx, cyp = (float)py + .5f - hy;
    float X[3], Y[3], OX[3], OY[3];
    #pragma unroll
    for (int i = 0; i < 3; i++)
    {
        const float rw = 1.f / vtx[i].w, row = 1.f / wing[i].w;
        X[i] = vtx[i].x * rw * hx - cxp;   Y[i] = vtx[i].y * rw * hy - cyp;
        OX[i] = wing[i].x * row * hx - cxp; OY[i] = wing[i].y * row * hy - cyp;
    }
    const float own = cross2(X[1] - X[0], Y[1] - Y[0], X[2] - X[0], Y[2] - Y[0]);
    bool sil[3];
    #pragma unroll
    for (int i = 0; i < 3; i++)
    {
        const int j = (i + 1) % 3, k = (i + 2) % 3;
        sil[i] = sign_bits_agree(cross2(X[j] - OX[i], Y[j] - OY[i], X[k] - OX[i], Y[k] - OY[i]), own);
    }
    AAEdgeResult r; r.edge = 0; r.alpha = 0.f; r.blend = false;
    if (!(sil[0] || sil[1] || sil[2])) return r;

    if (vertical)
    {
        #pragma unroll
        for (int i = 0; i < 3; i++) { float s = X[i]; X[i] = Y[i]; Y[i] = s; }
    }
    const float dir = toward_plus ? 1.f : -1.f;
    const float NEG_HUGE = -3.402823466e+38f;
    float ex[3], ey[3];
    Frac cand[3];
    #pragma unroll
    for (int i = 0; i < 3; i++)
    {
        const int j = (i + 1) % 3, k = (i + 2) % 3;
        ex[i] = X[k] - X[j]; ey[i] = Y[k] - Y[j];
        cand[i].num = dir * (X[j] * ey[i] - Y[j] * ex[i]);
        cand[i].den = ey[i];
        if (sign_bits_agree(Y[j], Y[k])) { cand[i].num = NEG_HUGE; cand[i].den = 1.f; }   // edge does not cross the axis
    }
    int best = frac_greater(cand[1], cand[0]) ? 1 : 0;
    if (frac_greater(cand[2], cand[0]) && frac_greater(cand[2], cand[1])) best = 2;
    r.edge = best;
    float c = NEG_HUGE;
    // only silhouette edges that are steeper than 45 degrees in the local frame are resolved by this pair
    const float bnum = best == 0 ? cand[0].num : (best == 1 ? cand[1].num : cand[2].num);
    const float bden = best == 0 ? cand[0].den : (best == 1 ? cand[1].den : cand[2].den);
    if (pick3(sil, best) && fabsf(bden) >= fabsf(pick3(ex, best))) c = bnum / bden;
    const float slack = .0625f;
    if (c > -slack && c < 1.f + slack)
    {
        c = fminf(fmaxf(c, 0.f), 1.f);
        r.alpha = dir * (.5f - c);
        r.blend = true;
    }
    return r;
}

// Gradient of the blend weight w.r.t. the two end points (clip space) of the deciding edge.
//   dd = dL/d(alpha) (= sum_c dy_c (color1_c - color0_c)); alpha = dir (1/2 - dir xc) -> dL/d(xc) = -dd.
//   xc = x1 - y1 ex / ey,  ex = x2 - x1, ey = y2 - y1, with 1/ey regularised as 1 / (ey + copysign(1e-3, ey)) everywhere.
// px, py: the owning pixel; p1, p2: edge end points; g1, g2 receive (d/dx, d/dy, d/dw).  No gradient at |alpha| >= 1/2
// (the crossing was clamped).
PM_HD void aa_edge_grad(float4 p1, float4 p2, int px, int py, float hx, float hy, bool vertical, float alpha, float dd,
                        float (&g1)[3], float (&g2)[3])
{
    float cxp = (float)px + .5f - hx, cyp = (float)py + .5f - hy;
    if (vertical)
    {
        float s;
        s = p1.x; p1.x = p1.y; p1.y = s;
        s = p2.x; p2.x = p2.y; p2.y = s;
        s = hx; hx = hy; hy = s;
        s = cxp; cxp = cyp; cyp = s;
    }
    const float rw1 = 1.f / p1.w, rw2 = 1.f / p2.w;
    const float x1 = p1.x * rw1 * hx - cxp, y1 = p1.y * rw1 * hy - cyp;
    const float x2 = p2.x * rw2 * hx - cxp, y2 = p2.y * rw2 * hy - cyp;
    const float ex = x2 - x1, ey = y2 - y1;
    const float rey = 1.f / (ey + copysignf(1e-3f, ey));
    const float xc = (x1 * ey - y1 * ex) * rey;            // regularised crossing
    const float k1 = -dd * rey * rw1, k2 = dd * rey * rw2;   // dL/d(xc) folded with d(xc)/d(.) 's common factor and the 1/w of the projection
    float a1 = k1 * hx * y2, b1 = k1 * hy * (xc - x2);
    float a2 = k2 * hx * y1, b2 = k2 * hy * (xc - x1);
    float c1 = -(p1.x * a1 + p1.y * b1) * rw1;
    float c2 = -(p2.x * a2 + p2.y * b2) * rw2;
    if (vertical) { float s; s = a1; a1 = b1; b1 = s; s = a2; a2 = b2; b2 = s; }
    const bool clamped = fabsf(alpha) >= 0.5f;
    g1[0] = clamped ? 0.f : a1; g1[1] = clamped ? 0.f : b1; g1[2] = clamped ? 0.f : c1;
    g2[0] = clamped ? 0.f : a2; g2[1] = clamped ? 0.f : b2; g2[2] = clamped ? 0.f : c2;
}

} // namespace pm
