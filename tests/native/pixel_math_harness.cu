// TEST INFRASTRUCTURE.  Runs the __host__ side of nvdiffrast_b200/csrc/pixel_math.cuh (the exact per-pixel expressions the
// CUDA kernels execute) over whole images on the CPU, so that tests/test_pixel_math_cpu.py can compare them with the oracle
// without a GPU.  Built by the test with:  nvcc -shared -Xcompiler -fPIC (host code only is called).
#include "pixel_math.cuh"
#include <stdint.h>
#include <string.h>

static int tri_id_of(float x) { if (x <= 16777216.f) return (int)x; int i; memcpy(&i, &x, 4); return i - 0x4a800000; }

// one image: rast ids are taken from `rast`; writes the re-shaded rast/db and accumulates g_pos (double, [V,4])
extern "C" void pmh_raster(const float* pos, const int32_t* tri, const float* rast, const float* dy, const float* ddb,
                           int H, int W, int with_db, float* rast_out, float* db_out, double* g_pos)
{
    const float xs = 2.f / (float)W, xo = 1.f / (float)W - 1.f, ys = 2.f / (float)H, yo = 1.f / (float)H - 1.f;
    for (int py = 0; py < H; py++) for (int px = 0; px < W; px++)
    {
        size_t pi = (size_t)py * W + px;
        int t = tri_id_of(rast[pi * 4 + 3]) - 1;
        float4 o = make_float4(0.f, 0.f, 0.f, 0.f), db = o;
        if (t >= 0)
        {
            const int32_t* vi = tri + 3 * t;
            const float4* vb = (const float4*)pos;
            float fx = xs * (float)px + xo, fy = ys * (float)py + yo;
            pm::PixelTri T;
            pm::pixel_tri(T, vb[vi[0]], vb[vi[1]], vb[vi[2]], fx, fy);
            pm::shade(T, xs, ys, rast[pi * 4 + 3], o, db);
            float gx[3], gy[3], gw[3];
            float4 h = ddb ? make_float4(ddb[pi * 4], ddb[pi * 4 + 1], ddb[pi * 4 + 2], ddb[pi * 4 + 3]) : make_float4(0.f, 0.f, 0.f, 0.f);
            if (with_db) pm::raster_grad<true>(T, fx, fy, xs, ys, dy[pi * 4], dy[pi * 4 + 1], h, true, gx, gy, gw);
            else         pm::raster_grad<false>(T, fx, fy, xs, ys, dy[pi * 4], dy[pi * 4 + 1], h, false, gx, gy, gw);
            for (int m = 0; m < 3; m++) { g_pos[4 * vi[m] + 0] += gx[m]; g_pos[4 * vi[m] + 1] += gy[m]; g_pos[4 * vi[m] + 3] += gw[m]; }
        }
        memcpy(rast_out + pi * 4, &o, 16); memcpy(db_out + pi * 4, &db, 16);
    }
}

// one image: analysis of every id discontinuity + gradients.  wings[T,3] = opposite vertex per edge or -1.
// items[k] = (px, py, d, edge, owner_is_second, alpha bits); returns the number of blending items.
extern "C" int pmh_antialias(const float* rast, const float* pos, const int32_t* tri, const int32_t* wings,
                             const float* color, const float* dy, int H, int W, int C, int V, int T,
                             int32_t* items, double* out_acc, double* g_color_acc, double* g_pos)
{
    const float hx = .5f * (float)W, hy = .5f * (float)H;
    const float4* vb = (const float4*)pos;
    int cnt = 0;
    for (int py = 0; py < H; py++) for (int px = 0; px < W; px++) for (int d = 0; d < 2; d++)
    {
        if (d == 0 && px >= W - 1) continue;
        if (d == 1 && py >= H - 1) continue;
        size_t p0 = (size_t)py * W + px, p1 = p0 + (d ? W : 1);
        if (rast[p0 * 4 + 3] == rast[p1 * 4 + 3]) continue;
        int t0 = tri_id_of(rast[p0 * 4 + 3]) - 1, t1 = tri_id_of(rast[p1 * 4 + 3]) - 1;
        int t = (t0 >= 0) ? t0 : t1;
        if (t0 >= 0 && t1 >= 0) t = (rast[p0 * 4 + 2] < rast[p1 * 4 + 2]) ? t0 : t1;
        bool second = (t == t1);
        int qx = px + (second ? 1 - d : 0), qy = py + (second ? d : 0);
        if (t < 0 || t >= T) continue;
        float4 v[3], o[3];
        for (int i = 0; i < 3; i++) { v[i] = vb[tri[3 * t + i]]; int w = wings[3 * t + i]; o[i] = (w < 0) ? v[i] : vb[w]; }
        pm::AAEdgeResult r = pm::aa_edge(v, o, qx, qy, hx, hy, d != 0, !second);
        if (!r.blend) continue;
        int32_t ab; memcpy(&ab, &r.alpha, 4);
        int32_t* q = items + 6 * cnt++;
        q[0] = px; q[1] = py; q[2] = d; q[3] = r.edge; q[4] = second; q[5] = ab;
        size_t dst = (r.alpha > 0.f) ? p0 : p1;
        float dd = 0.f;
        for (int c = 0; c < C; c++)
        {
            out_acc[dst * C + c] += (double)(r.alpha * (color[p1 * C + c] - color[p0 * C + c]));
            float y = dy[dst * C + c];
            if (y != 0.f) { dd += y * (color[p1 * C + c] - color[p0 * C + c]); g_color_acc[p0 * C + c] -= r.alpha * y; g_color_acc[p1 * C + c] += r.alpha * y; }
        }
        if (dd == 0.f || ab == 0) continue;
        int i1 = (r.edge + 1) % 3, i2 = (r.edge + 2) % 3;
        int v1 = tri[3 * t + i1], v2 = tri[3 * t + i2];
        float g1[3], g2[3];
        pm::aa_edge_grad(vb[v1], vb[v2], qx, qy, hx, hy, d != 0, r.alpha, dd, g1, g2);
        g_pos[4 * v1 + 0] += g1[0]; g_pos[4 * v1 + 1] += g1[1]; g_pos[4 * v1 + 3] += g1[2];
        g_pos[4 * v2 + 0] += g2[0]; g_pos[4 * v2 + 1] += g2[1]; g_pos[4 * v2 + 3] += g2[2];
    }
    return cnt;
}
