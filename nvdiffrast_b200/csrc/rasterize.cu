// rasterize.cu -- B200-native replacement of the reference's rasterize op:
//   rasterize_fwd_cuda  (csrc/torch/torch_rasterize.cpp:43-166)  = CudaRaster 4-stage pipeline
//                        (csrc/common/cudaraster/impl/*.inl) + pixel shader (csrc/common/rasterize.cu:15-114)
//   rasterize_grad[_db] (csrc/torch/torch_rasterize.cpp:171-262, csrc/common/rasterize.cu:119-277)
//
// Design (not a port): visibility is order-free.  The reference's ordered bin/tile queues exist only so the ROP can
// resolve "min depth, ties -> later triangle" sequentially (impl/FineRaster.inl:344-361,152-172).  Here the same rule
// is a total order on the key (depth_u32, ~triangle_id), so
//   * BINS path   : one CTA per 64x64-pixel bin scans the image's compact bbox array (L2 resident), keeps the best
//                   key of 16 pixels per thread in registers (no atomics, no queues, no host sync), and runs the
//                   pixel shader on the winners before anything is written: rast/rast_db leave the SM once, as
//                   full 512-byte warp stores.  The id/depth buffers of the reference never exist in HBM.
//   * SCATTER path: for heavy geometry (many tiny triangles) triangle setup and coverage are fused, fragments go to
//                   a 64-bit atomicMin key buffer, a resolve pass shades.
// Snapping, culling, clipping, fill rule and the fixed-point depth plane reproduce the reference bit for bit
// (see oracle/oracle.c for the scalar restatement with citations).
#include "common.cuh"
#include "pixel_math.cuh"

namespace {

// ------------------------------------------------------------------------------------------------
// Parameters

struct RastParams
{
    const float*    pos;
    const int32_t*  tri;
    const int32_t*  ranges;         // device [N,2] or NULL
    int             N, V, T, H, W;          // H, W: full output resolution
    int             vpW, vpH, offX, offY;   // current viewport (<= 2048^2) inside the output; larger images are tiled
    float           vxs, vxo, vys, vyo;     // clip-space remap into the viewport: x' = x*vxs + w*vxo (impl/RasterImpl.cpp:295-298)
    int             instance_mode;
    int             max_tris;       // per image
    int             cap;            // record slots per image (7 * max_tris, rounded up to even)
    const uint32_t* peel_in;
    uint32_t*       depth_out;
    // BINS scratch
    int*            extra;          // [N] extra (clipped) sub-triangle counters
    uint32_t*       binmask;        // [N, 32] row y, bit x: some triangle's bounding box touches bin (x, y) of the current viewport
    uint2*          bbox;           // [N, cap]
    uint2*          cbox;           // [N, nchunk]: union of the bounding boxes of 32 consecutive triangle slots ("chunk")
    int             nchunk;         // chunk boxes per image (even, so that every image's array is a 16-byte aligned TMA source)
    uint4*          rec;            // [N, cap, 2]
    // SCATTER scratch
    unsigned long long* keys;       // [N,H,W]
    float*          rast;
    float*          rast_db;
    float           xs, xo, ys, yo;
};

// One snapped, CCW-normalised sub-triangle with its fixed-point depth plane.
struct SubTri
{
    int      x0, y0, x1, y1, x2, y2;    // 1/16 px, origin = viewport centre (fits S16)
    uint32_t zx, zy, zb;
    int      bx0, bx1, by0, by1;        // inclusive pixel bbox of candidate samples, clipped to the image
};

// ------------------------------------------------------------------------------------------------
// Triangle setup (bit-exact with impl/TriangleSetup.inl + impl/Util.inl:184-210).

static __device__ __forceinline__ int f2i_rni_sat(float a) { int v; asm("cvt.rni.sat.s32.f32 %0, %1;" : "=r"(v) : "f"(a)); return v; }

static __device__ __forceinline__ void depth_plane(float zv0, float zv1, float zv2, int v0x, int v0y,
                                                   int d1x, int d1y, int d2x, int d2y, float areaRcp,
                                                   uint32_t& zx, uint32_t& zy, uint32_t& zb)
{
    float mx = fmaxf(fmaxf(zv0, zv1), zv2);
    int sh = min(max((__float_as_int(mx) >> 23) - (127 + 22), 0), 8);
    int t0 = (int)(__float2uint_rz(zv0) >> sh);
    int t1 = (int)((__float2uint_rz(zv1) >> sh) - (uint32_t)t0);
    int t2 = (int)((__float2uint_rz(zv2) >> sh) - (uint32_t)t0);

    uint32_t rcpMant = ((uint32_t)__float_as_int(areaRcp) & 0x007FFFFFu) | 0x00800000u;
    int rcpShift = (23 + 127) - (__float_as_int(areaRcp) >> 23);

    long long xc = ((long long)t1 * d2y - (long long)t2 * d1y) * (long long)rcpMant;
    long long yc = ((long long)t2 * d1x - (long long)t1 * d2x) * (long long)rcpMant;
    zx = (uint32_t)(xc >> (rcpShift - (sh + 4)));
    zy = (uint32_t)(yc >> (rcpShift - (sh + 4)));

    int centerX = (v0x * 2 + min(min(d1x, d2x), 0) + max(max(d1x, d2x), 0)) >> 5;
    int centerY = (v0y * 2 + min(min(d1y, d2y), 0) + max(max(d1y, d2y), 0)) >> 5;
    int vcx = v0x - (int)((uint32_t)centerX << 4);
    int vcy = v0y - (int)((uint32_t)centerY << 4);

    uint32_t pz = (uint32_t)t0 << sh;
    pz -= (uint32_t)(((xc >> 13) * vcx + (yc >> 13) * vcy) >> (rcpShift - (sh + 13)));
    pz -= zx * (uint32_t)centerX + zy * (uint32_t)centerY;
    zb = pz;
}

// Snap three clip-space vertices, run the degenerate / between-samples culls, normalise winding, build the depth plane.
static __device__ __forceinline__ bool snap_and_setup(float4 v0, float4 v1, float4 v2, int W, int H, SubTri& s)
{
    const float sx = (float)(W << 3), sy = (float)(H << 3);
    float rw0 = __frcp_rn(v0.w), rw1 = __frcp_rn(v1.w), rw2 = __frcp_rn(v2.w);
    int p0x = f2i_rni_sat(__fmul_rn(__fmul_rn(v0.x, rw0), sx)), p0y = f2i_rni_sat(__fmul_rn(__fmul_rn(v0.y, rw0), sy));
    int p1x = f2i_rni_sat(__fmul_rn(__fmul_rn(v1.x, rw1), sx)), p1y = f2i_rni_sat(__fmul_rn(__fmul_rn(v1.y, rw1), sy));
    int p2x = f2i_rni_sat(__fmul_rn(__fmul_rn(v2.x, rw2), sx)), p2y = f2i_rni_sat(__fmul_rn(__fmul_rn(v2.y, rw2), sy));
    int lox = min(min(p0x, p1x), p2x), loy = min(min(p0y, p1y), p2y);
    int hix = max(max(p0x, p1x), p2x), hiy = max(max(p0y, p1y), p2y);

    int d1x = p1x - p0x, d1y = p1y - p0y, d2x = p2x - p0x, d2y = p2y - p0y;
    int area = d1x * d2y - d1y * d2x;
    if (area == 0) return false;

    const int biasX = (W << 3) - 8, biasY = (H << 3) - 8;
    int sxlo = (lox + 15 + biasX) & -16, sylo = (loy + 15 + biasY) & -16;
    int sxhi = (hix + biasX) & -16,      syhi = (hiy + biasY) & -16;
    if (sxlo > sxhi || sylo > syhi) return false;

    int diff = (sxhi + syhi - sxlo) - sylo;
    if (diff <= 16)
    {
        int t0x = p0x + biasX - sxlo, t0y = p0y + biasY - sylo;
        int t1x = p1x + biasX - sxlo, t1y = p1y + biasY - sylo;
        int t2x = p2x + biasX - sxlo, t2y = p2y + biasY - sylo;
        int e0 = t0x * t1y - t0y * t1x, e1 = t1x * t2y - t1y * t2x, e2 = t2x * t0y - t2y * t0x;
        if (area < 0) { e0 = -e0; e1 = -e1; e2 = -e2; }
        if (e0 < 0 || e1 < 0 || e2 < 0)
        {
            if (diff == 0) return false;
            t0x = p0x + biasX - sxhi; t0y = p0y + biasY - syhi;
            t1x = p1x + biasX - sxhi; t1y = p1y + biasY - syhi;
            t2x = p2x + biasX - sxhi; t2y = p2y + biasY - syhi;
            e0 = t0x * t1y - t0y * t1x; e1 = t1x * t2y - t1y * t2x; e2 = t2x * t0y - t2y * t0x;
            if (area < 0) { e0 = -e0; e1 = -e1; e2 = -e2; }
            if (e0 < 0 || e1 < 0 || e2 < 0) return false;
        }
    }

    float z0 = v0.z, z1 = v1.z, z2 = v2.z;
    if (area < 0)
    {
        int t;
        t = d1x; d1x = d2x; d2x = t;  t = d1y; d1y = d2y; d2y = t;
        t = p1x; p1x = p2x; p2x = t;  t = p1y; p1y = p2y; p2y = t;
        float f = z1; z1 = z2; z2 = f;  f = rw1; rw1 = rw2; rw2 = f;
        area = -area;
    }

    // zcoef = (F32)(CR_DEPTH_MAX - CR_DEPTH_MIN) * 0.5f = 0x4EFFFF76, zbias = 2^31; mul then fma as in the reference build.
    const float zcoef = __int_as_float(0x4EFFFF76), zbias = 2147483648.0f;
    float zv0 = __fmaf_rn(__fmul_rn(z0, zcoef), rw0, zbias);
    float zv1 = __fmaf_rn(__fmul_rn(z1, zcoef), rw1, zbias);
    float zv2 = __fmaf_rn(__fmul_rn(z2, zcoef), rw2, zbias);
    depth_plane(zv0, zv1, zv2, p0x + (W << 3) - 8, p0y + (H << 3) - 8, d1x, d1y, d2x, d2y,
                __frcp_rn((float)area), s.zx, s.zy, s.zb);

    s.x0 = p0x; s.y0 = p0y; s.x1 = p1x; s.y1 = p1y; s.x2 = p2x; s.y2 = p2y;
    s.bx0 = max(sxlo >> 4, 0); s.bx1 = min(sxhi >> 4, W - 1);
    s.by0 = max(sylo >> 4, 0); s.by1 = min(syhi >> 4, H - 1);
    return (s.bx0 <= s.bx1) && (s.by0 <= s.by1);
}

// Sutherland-Hodgman against one plane in barycentric space.  FMA shapes match the reference's PTX:
// plane value fma(c2,b1,fma(c1,b0,c0)), interpolated bary fma(in_a, ac, in_b*bc)  (impl/Util.inl:101-130).
static __device__ __noinline__ int clip_plane(float* out, const float* in, int n, float c0, float c1, float c2)
{
    int m = 0;
    if (n >= 3)
    {
        int ai = (n - 1) * 2;
        float av = __fmaf_rn(c2, in[ai + 1], __fmaf_rn(c1, in[ai], c0));
        for (int bi = 0; bi < n * 2; bi += 2)
        {
            float bv = __fmaf_rn(c2, in[bi + 1], __fmaf_rn(c1, in[bi], c0));
            if (__fmul_rn(av, bv) < 0.0f)
            {
                float bc = __fdiv_rn(av, __fsub_rn(av, bv));
                float ac = __fsub_rn(1.0f, bc);
                out[m + 0] = __fmaf_rn(in[ai + 0], ac, __fmul_rn(in[bi + 0], bc));
                out[m + 1] = __fmaf_rn(in[ai + 1], ac, __fmul_rn(in[bi + 1], bc));
                m += 2;
            }
            if (bv >= 0.0f) { out[m + 0] = in[bi + 0]; out[m + 1] = in[bi + 1]; m += 2; }
            ai = bi; av = bv;
        }
    }
    return m >> 1;
}

// Clip a triangle that is neither trivially rejected nor fully inside (impl/Util.inl:134-160).  Returns the number of
// polygon vertices (<= 9) and their barycentrics.
static __device__ __noinline__ int clip_frustum(float* bary, float4 v0, float4 v1, float4 v2, float4 d1, float4 d2)
{
    int n = 3;
    float tmp[18];
    bary[0] = 0.f; bary[1] = 0.f; bary[2] = 1.f; bary[3] = 0.f; bary[4] = 0.f; bary[5] = 1.f;
    if ((v0.w < fabsf(v0.x)) | (v1.w < fabsf(v1.x)) | (v2.w < fabsf(v2.x)))
    {
        n = clip_plane(tmp, bary, n, __fadd_rn(v0.w, v0.x), __fadd_rn(d1.w, d1.x), __fadd_rn(d2.w, d2.x));
        n = clip_plane(bary, tmp, n, __fsub_rn(v0.w, v0.x), __fsub_rn(d1.w, d1.x), __fsub_rn(d2.w, d2.x));
    }
    if ((v0.w < fabsf(v0.y)) | (v1.w < fabsf(v1.y)) | (v2.w < fabsf(v2.y)))
    {
        n = clip_plane(tmp, bary, n, __fadd_rn(v0.w, v0.y), __fadd_rn(d1.w, d1.y), __fadd_rn(d2.w, d2.y));
        n = clip_plane(bary, tmp, n, __fsub_rn(v0.w, v0.y), __fsub_rn(d1.w, d1.y), __fsub_rn(d2.w, d2.y));
    }
    if ((v0.w < fabsf(v0.z)) | (v1.w < fabsf(v1.z)) | (v2.w < fabsf(v2.z)))
    {
        n = clip_plane(tmp, bary, n, __fadd_rn(v0.w, v0.z), __fadd_rn(d1.w, d1.z), __fadd_rn(d2.w, d2.z));
        n = clip_plane(bary, tmp, n, __fsub_rn(v0.w, v0.z), __fsub_rn(d1.w, d1.z), __fsub_rn(d2.w, d2.z));
    }
    return n;
}

static __device__ __forceinline__ float4 bary_point(float4 v0, float4 d1, float4 d2, float b0, float b1)
{
    return make_float4(__fmaf_rn(d2.x, b1, __fmaf_rn(d1.x, b0, v0.x)), __fmaf_rn(d2.y, b1, __fmaf_rn(d1.y, b0, v0.y)),
                       __fmaf_rn(d2.z, b1, __fmaf_rn(d1.z, b0, v0.z)), __fmaf_rn(d2.w, b1, __fmaf_rn(d1.w, b0, v0.w)));
}

// Classify a triangle (impl/TriangleSetup.inl:271-352): 0 = culled, 1 = fully inside the frustum, 2 = needs clipping.
static __device__ __forceinline__ int classify(float4 v0, float4 v1, float4 v2)
{
    if ((v0.w < fabsf(v0.x)) | (v0.w < fabsf(v0.y)) | (v0.w < fabsf(v0.z)))
    {
        if (((v0.w < +v0.x) & (v1.w < +v1.x) & (v2.w < +v2.x)) |
            ((v0.w < -v0.x) & (v1.w < -v1.x) & (v2.w < -v2.x)) |
            ((v0.w < +v0.y) & (v1.w < +v1.y) & (v2.w < +v2.y)) |
            ((v0.w < -v0.y) & (v1.w < -v1.y) & (v2.w < -v2.y)) |
            ((v0.w < +v0.z) & (v1.w < +v1.z) & (v2.w < +v2.z)) |
            ((v0.w < -v0.z) & (v1.w < -v1.z) & (v2.w < -v2.z)))
            return 0;
    }
    if (v0.w >= fmaxf(fmaxf(fabsf(v0.x), fabsf(v0.y)), fabsf(v0.z)) &&
        v1.w >= fmaxf(fmaxf(fabsf(v1.x), fabsf(v1.y)), fabsf(v1.z)) &&
        v2.w >= fmaxf(fmaxf(fabsf(v2.x), fabsf(v2.y)), fabsf(v2.z)))
        return 1;
    return 2;
}

// Fetch the three clip-space vertices of triangle triIdx for image n; false if any index is bad.
static __device__ __forceinline__ bool fetch_triangle(const RastParams& p, int n, int triIdx, float4& v0, float4& v1, float4& v2)
{
    uint32_t i0 = (uint32_t)__ldg(p.tri + triIdx * 3 + 0), i1 = (uint32_t)__ldg(p.tri + triIdx * 3 + 1), i2 = (uint32_t)__ldg(p.tri + triIdx * 3 + 2);
    if (i0 >= (uint32_t)p.V || i1 >= (uint32_t)p.V || i2 >= (uint32_t)p.V) return false;
    const float4* vb = (const float4*)p.pos + (p.instance_mode ? (size_t)n * p.V : 0);
    v0 = __ldg(vb + i0); v1 = __ldg(vb + i1); v2 = __ldg(vb + i2);
    // viewport remap, mul + fma as in the reference build (impl/TriangleSetup.inl:262-267); identity for a single viewport
    v0.x = __fmaf_rn(v0.x, p.vxs, __fmul_rn(v0.w, p.vxo)); v0.y = __fmaf_rn(v0.y, p.vys, __fmul_rn(v0.w, p.vyo));
    v1.x = __fmaf_rn(v1.x, p.vxs, __fmul_rn(v1.w, p.vxo)); v1.y = __fmaf_rn(v1.y, p.vys, __fmul_rn(v1.w, p.vyo));
    v2.x = __fmaf_rn(v2.x, p.vxs, __fmul_rn(v2.w, p.vxo)); v2.y = __fmaf_rn(v2.y, p.vys, __fmul_rn(v2.w, p.vyo));
    return true;
}

// ------------------------------------------------------------------------------------------------
// Pixel shader: replaces RasterizeCudaFwdShaderKernel (csrc/common/rasterize.cu:15-114).  The arithmetic lives in
// pixel_math.cuh (pm::pixel_tri + pm::shade), shared with the backward kernel and executable on the host for the tests.

static __device__ __forceinline__ void shade_pixel(const RastParams& p, int n, int triIdx, int px, int py, float4& o, float4& db)
{
    const float4* vb = (const float4*)p.pos + (p.instance_mode ? (size_t)n * p.V : 0);
    const int32_t* vi = p.tri + triIdx * 3;
    pm::PixelTri t;
    pm::pixel_tri(t, __ldg(vb + __ldg(vi)), __ldg(vb + __ldg(vi + 1)), __ldg(vb + __ldg(vi + 2)),
                  p.xs * (float)px + p.xo, p.ys * (float)py + p.yo);
    pm::shade(t, p.xs, p.ys, triidx_to_float(triIdx + 1), o, db);
}

// ------------------------------------------------------------------------------------------------
// Edge functions in pixel steps.  For a directed edge a->b of the CCW triangle and pixel (X,Y) with sample
// s = (16X + 8 - 8W, 16Y + 8 - 8H):  c = (a-s) x (b-a) - excl,  covered iff c >= 0 for all three edges, where
// excl = 1 for edges with dy > 0 or (dy == 0 and dx <= 0)  (impl/Util.inl:304-309).  |c| <= 2^30 for W,H <= 2048.

struct Edge { int c, sx, sy; };   // value at a base pixel, step per +1 pixel in x / y

static __device__ __forceinline__ Edge make_edge(int ax, int ay, int bx, int by, int X, int Y, int W, int H)
{
    int dx = bx - ax, dy = by - ay;
    int ox = ax - (16 * X + 8 - 8 * W), oy = ay - (16 * Y + 8 - 8 * H);
    Edge e;
    e.c = ox * dy - oy * dx - ((dy > 0 || (dy == 0 && dx <= 0)) ? 1 : 0);
    e.sx = -16 * dy;
    e.sy = 16 * dx;
    return e;
}

// ================================================================================================
// BINS path
// ================================================================================================

#define BIN_LOG2   6
#define BIN_SIZE   64
#define BIN_LIST   256          // triangles set up in shared memory at a time (longer per-bin lists are processed in chunks)
#define BIN_IDX    2048         // records scanned per segment = capacity of the per-bin index list
#define BIN_THREADS 256
#define B200DR_BINS_MAX_TRIS 49152
#define BIN_MAXCHUNK 1024       // chunk boxes staged per TMA slab (32768 triangles); larger images re-stage
#define BIN_ROWS   32           // a viewport is at most 2048 = 32 * BIN_SIZE pixels wide and high: one mask word per bin row
#ifndef BIN_MASK
#define BIN_MASK 1              // bins no bounding box touches are cleared without scanning anything
#endif

struct __align__(16) BinTri     // 64 bytes, broadcast-read by every thread of the CTA
{
    int e0c, e0sx, e0sy, e1c;
    int e1sx, e1sy, e2c, e2sx;
    int e2sy; uint32_t zc, zx, zy;
    uint32_t id, box, pad0, pad1;   // box = x0 | x1<<8 | y0<<16 | y1<<24, bin-relative pixels
};

static __device__ __forceinline__ void mark_bins(uint32_t* rows, int x0, int x1, int y0, int y1)
{
    if (x0 > x1 || y0 > y1) return;
    const uint32_t bits = (0xFFFFFFFFu >> (31 - (x1 >> BIN_LOG2))) & (0xFFFFFFFFu << (x0 >> BIN_LOG2));
    for (int by = y0 >> BIN_LOG2; by <= (y1 >> BIN_LOG2); by++) atomicOr(rows + by, bits);
}

struct SetupEmitRecords
{
    const RastParams& p; int n; int slot0; int emitted; int id; uint2 first;    // first: bounding box that went to slot0
    __device__ __forceinline__ void operator()(const SubTri& s)
    {
        int slot = slot0;
        if (emitted > 0) slot = p.max_tris + atomicAdd(p.extra + n, 1);
        size_t base = (size_t)n * p.cap + slot;
        const uint2 box = make_uint2((uint32_t)s.bx0 | ((uint32_t)s.bx1 << 16), (uint32_t)s.by0 | ((uint32_t)s.by1 << 16));
        if (emitted == 0) first = box;
        else if (BIN_MASK) mark_bins(p.binmask + n * BIN_ROWS, s.bx0, s.bx1, s.by0, s.by1);    // extras are not covered by the chunk boxes
        emitted++;
        p.bbox[base] = box;
        p.rec[base * 2 + 0] = make_uint4(((uint32_t)s.x0 & 0xFFFFu) | ((uint32_t)s.y0 << 16), ((uint32_t)s.x1 & 0xFFFFu) | ((uint32_t)s.y1 << 16),
                                          ((uint32_t)s.x2 & 0xFFFFu) | ((uint32_t)s.y2 << 16), s.zx);
        p.rec[base * 2 + 1] = make_uint4(s.zy, s.zb, (uint32_t)id, 0u);
    }
};

// Run setup for one triangle and hand every visible sub-triangle to emit().
template <class Emit>
static __device__ __forceinline__ void process_triangle(const RastParams& p, int n, int triIdx, Emit& emit)
{
    float4 v0, v1, v2;
    if (!fetch_triangle(p, n, triIdx, v0, v1, v2)) return;
    int cls = classify(v0, v1, v2);
    if (cls == 0) return;
    SubTri s;
    if (cls == 1)
    {
        if (snap_and_setup(v0, v1, v2, p.vpW, p.vpH, s)) emit(s);
        return;
    }
    // Rare: clip to the frustum, fan-triangulate from the first polygon vertex (impl/TriangleSetup.inl:355-434).
    float bary[18];
    float4 d1 = make_float4(__fsub_rn(v1.x, v0.x), __fsub_rn(v1.y, v0.y), __fsub_rn(v1.z, v0.z), __fsub_rn(v1.w, v0.w));
    float4 d2 = make_float4(__fsub_rn(v2.x, v0.x), __fsub_rn(v2.y, v0.y), __fsub_rn(v2.z, v0.z), __fsub_rn(v2.w, v0.w));
    int nv = clip_frustum(bary, v0, v1, v2, d1, d2);
    float4 c0 = bary_point(v0, d1, d2, bary[0], bary[1]);
    float4 c1 = bary_point(v0, d1, d2, bary[2], bary[3]);
    for (int i = 2; i < nv; i++)
    {
        float4 c2 = bary_point(v0, d1, d2, bary[i * 2 + 0], bary[i * 2 + 1]);
        if (snap_and_setup(c0, c1, c2, p.vpW, p.vpH, s)) emit(s);
        c1 = c2;
    }
}

static __device__ __forceinline__ void image_range(const RastParams& p, int n, int& start, int& count)
{
    if (p.instance_mode) { start = 0; count = p.T; }
    else { start = __ldg(p.ranges + 2 * n); count = __ldg(p.ranges + 2 * n + 1); }
}

// Thread per (image, triangle).  Besides the records, every warp writes the union of its 32 first-slot bounding boxes: meshes
// keep neighbouring triangles at neighbouring indices, so a bin can discard 32 triangles with one test (raster_bins_kernel).
__global__ void __launch_bounds__(128) raster_setup_kernel(const RastParams p)
{
    int t = blockIdx.x * 128 + threadIdx.x;
    int n = blockIdx.y;
    int start, count;
    image_range(p, n, start, count);
    SetupEmitRecords emit = { p, n, t, 0, 0, make_uint2(0x0000FFFFu, 0x0000FFFFu) };                    // empty box: lo > hi
    uint2 bb = emit.first;
    if (t < p.max_tris)
    {
        if (t < count)
        {
            int triIdx = start + t;
            if ((uint32_t)triIdx < (uint32_t)p.T)
            {
                emit.id = triIdx + 1;
                process_triangle(p, n, triIdx, emit);
            }
        }
        if (emit.emitted == 0) p.bbox[(size_t)n * p.cap + t] = bb;
        else                   bb = emit.first;
    }
    // union over the warp (whole warps stay alive for the shuffles; lanes without a box contribute the empty box)
    int x0 = (int)(bb.x & 0xFFFFu), x1 = (int)(bb.x >> 16), y0 = (int)(bb.y & 0xFFFFu), y1 = (int)(bb.y >> 16);
    if (x0 > x1) { x0 = 0xFFFF; x1 = 0; y0 = 0xFFFF; y1 = 0; }
    #pragma unroll
    for (int d = 16; d > 0; d >>= 1)
    {
        x0 = min(x0, __shfl_xor_sync(0xffffffffu, x0, d)); x1 = max(x1, __shfl_xor_sync(0xffffffffu, x1, d));
        y0 = min(y0, __shfl_xor_sync(0xffffffffu, y0, d)); y1 = max(y1, __shfl_xor_sync(0xffffffffu, y1, d));
    }
    if ((threadIdx.x & 31) == 0 && (t >> 5) < p.nchunk)
    {
        p.cbox[(size_t)n * p.nchunk + (t >> 5)] = make_uint2((uint32_t)x0 | ((uint32_t)x1 << 16), (uint32_t)y0 | ((uint32_t)y1 << 16));
        if (BIN_MASK && x0 <= x1) mark_bins(p.binmask + n * BIN_ROWS, x0, x1, y0, y1);
    }
}

static __device__ __forceinline__ int s16lo(uint32_t v) { return (int)(short)(v & 0xFFFFu); }
static __device__ __forceinline__ int s16hi(uint32_t v) { return (int)v >> 16; }

#ifndef BIN_MINB
#define BIN_MINB 4          // 63 registers, 4 CTAs per SM: 0.55 ms on c3; unconstrained (95 registers, 2 CTAs) 0.71 ms
#endif
template <bool PEEL_IN, bool DEPTH_OUT>
__global__ void __launch_bounds__(BIN_THREADS, BIN_MINB) raster_bins_kernel(const RastParams p)
{
    __shared__ BinTri   s_tri[BIN_LIST];            // 16 KB
    __shared__ int      s_idx[BIN_IDX];             //  8 KB
    __shared__ __align__(16) uint32_t s_ids[BIN_SIZE * BIN_SIZE]; // 16 KB; during the scan: chunk boxes + hit chunks + per-warp lists
    __shared__ __align__(8) unsigned long long s_mbar;
    __shared__ int      s_count, s_nhit;
    uint2* const s_cb = (uint2*)s_ids;                                  // [BIN_MAXCHUNK] chunk boxes of this image (8 KB), one TMA bulk copy
    unsigned short* const s_hit = (unsigned short*)(s_ids + 2 * BIN_MAXCHUNK);            // [BIN_IDX / 32] chunks of the segment that touch the bin
    unsigned char* const s_wl = (unsigned char*)(s_ids + 2 * BIN_MAXCHUNK + BIN_IDX / 64); // [8 warps][BIN_LIST] per-warp triangle lists

    const int tid = threadIdx.x;
    const int n = blockIdx.z;
    const int binX0 = blockIdx.x << BIN_LOG2, binY0 = blockIdx.y << BIN_LOG2;
    const int binX1 = min(binX0 + BIN_SIZE, p.vpW) - 1, binY1 = min(binY0 + BIN_SIZE, p.vpH) - 1;   // viewport-local pixels

    // Thread -> 4x4 pixel block.  A warp covers 32x16 pixels (8x4 blocks).
    const int lane = tid & 31, warp = tid >> 5;
    const int wx = (warp & 1) << 5, wy = (warp >> 1) << 4;             // this warp's 32x16 region, bin-relative
    const int tx = wx + ((lane & 7) << 2);
    const int ty = wy + ((lane >> 3) << 2);

    if (BIN_MASK && !((__ldg(p.binmask + n * BIN_ROWS + blockIdx.y) >> blockIdx.x) & 1u))
    {
        // no bounding box touches this bin (CTA-uniform): cleared pixels, cleared depth, nothing scanned
        const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
        #pragma unroll 4
        for (int k = 0; k < (BIN_SIZE * BIN_SIZE) / BIN_THREADS; k++)
        {
            const int q = tid + k * BIN_THREADS;
            const int lx = q & (BIN_SIZE - 1), ly = q >> BIN_LOG2;
            if (binX0 + lx >= p.vpW || binY0 + ly >= p.vpH) continue;
            const size_t pi = ((size_t)n * p.H + p.offY + binY0 + ly) * p.W + p.offX + binX0 + lx;
            st_stream_f4((float4*)p.rast + pi, zero);
            st_stream_f4((float4*)p.rast_db + pi, zero);
            if (DEPTH_OUT) p.depth_out[pi] = DR_DEPTH_MAX;
        }
        return;
    }

    if (tid == 0) mbar_init(&s_mbar, 1);

    // best visibility key of each of the thread's 16 pixels: (depth << 32) | ~id.  "Nearer depth wins, equal depth: the later
    // triangle wins" (impl/FineRaster.inl:344-361, 152-172) is the unsigned 64-bit order of these keys; empty = (DEPTH_MAX, id 0).
    unsigned long long best[16];
    uint32_t peelD[16];
    #pragma unroll
    for (int i = 0; i < 16; i++) best[i] = ((unsigned long long)DR_DEPTH_MAX << 32) | 0xFFFFFFFFull;
    if (PEEL_IN)
    {
        #pragma unroll
        for (int j = 0; j < 4; j++)
        {
            int ly = binY0 + ty + j, lx = binX0 + tx;
            int Y = p.offY + ly, X = p.offX + lx;
            if (ly < p.vpH && lx + 3 < p.vpW && (p.W & 3) == 0 && (p.offX & 3) == 0)
            {
                uint4 v = *(const uint4*)(p.peel_in + ((size_t)n * p.H + Y) * p.W + X);
                peelD[j * 4 + 0] = v.x; peelD[j * 4 + 1] = v.y; peelD[j * 4 + 2] = v.z; peelD[j * 4 + 3] = v.w;
            }
            else
            {
                #pragma unroll
                for (int i = 0; i < 4; i++)
                    peelD[j * 4 + i] = (ly < p.vpH && lx + i < p.vpW) ? p.peel_in[((size_t)n * p.H + Y) * p.W + X + i] : 0xFFFFFFFFu;
            }
        }
    }

    const int total = p.max_tris + min(p.extra[n], p.cap - p.max_tris);
    const uint2* bboxes = p.bbox + (size_t)n * p.cap;
    const uint4* recs = p.rec + (size_t)n * p.cap * 2;
    const unsigned lt = (1u << lane) - 1u;

    bool any = false;
    uint32_t phase = 0;
    __syncthreads();                                            // barrier initialised
    // Records are scanned in segments of BIN_IDX records (the index list cannot overflow).  Two levels: a thread tests the
    // union box of a chunk of 32 consecutive slots (staged above); warps then expand the chunks that touch the bin, one
    // coalesced 256-byte read of the 32 boxes each, and append the hits to s_idx with one shared-memory atomic per chunk.
    // Clipped extras (slots >= max_tris, rare) have no chunk boxes and are tested one by one.
    for (int seg = 0; seg < total; seg += BIN_IDX)
    {
        const int seg_end = min(seg + BIN_IDX, total);
        const int c0 = seg >> 5, cb_base = c0 & ~(BIN_MAXCHUNK - 1);
        if (tid == 0)
        {
            s_count = 0; s_nhit = 0;
            // every BIN_MAXCHUNK chunks (= 32768 triangles; once per image for any scene the BINS path is picked for) the next
            // slab of chunk boxes is staged into shared memory by one TMA bulk copy; nchunk is even: a multiple of 16 bytes
            if (c0 == cb_base && c0 < p.nchunk)
                tma_load_1d(s_cb, p.cbox + (size_t)n * p.nchunk + c0, (uint32_t)min(BIN_MAXCHUNK, p.nchunk - c0) * 8u, &s_mbar);
        }
        __syncthreads();
        if (c0 == cb_base && c0 < p.nchunk) { mbar_wait(&s_mbar, phase); phase ^= 1u; }
        const int c_end = (min(seg_end, p.max_tris) + 31) >> 5;
        for (int c = c0 + tid; c < c_end; c += BIN_THREADS)
        {
            uint2 cb = s_cb[c - cb_base];
            int x0 = (int)(cb.x & 0xFFFFu), x1 = (int)(cb.x >> 16), y0 = (int)(cb.y & 0xFFFFu), y1 = (int)(cb.y >> 16);
            if (x0 <= binX1 && x1 >= binX0 && y0 <= binY1 && y1 >= binY0) s_hit[atomicAdd(&s_nhit, 1)] = (unsigned short)(c - c0);
        }
        for (int i = max(seg, p.max_tris) + tid; i < seg_end; i += BIN_THREADS)
        {
            uint2 bb = __ldg(bboxes + i);
            int x0 = (int)(bb.x & 0xFFFFu), x1 = (int)(bb.x >> 16), y0 = (int)(bb.y & 0xFFFFu), y1 = (int)(bb.y >> 16);
            if (x0 <= binX1 && x1 >= binX0 && y0 <= binY1 && y1 >= binY0 && x0 <= x1) s_idx[atomicAdd(&s_count, 1)] = i;
        }
        __syncthreads();
        const int nhit = s_nhit;
        for (int h = warp; h < nhit; h += BIN_THREADS / 32)
        {
            const int i = ((c0 + (int)s_hit[h]) << 5) + lane;
            bool hit = false;
            if (i < p.max_tris)
            {
                uint2 bb = __ldg(bboxes + i);
                int x0 = (int)(bb.x & 0xFFFFu), x1 = (int)(bb.x >> 16), y0 = (int)(bb.y & 0xFFFFu), y1 = (int)(bb.y >> 16);
                hit = x0 <= binX1 && x1 >= binX0 && y0 <= binY1 && y1 >= binY0 && x0 <= x1;
            }
            const unsigned m = __ballot_sync(0xffffffffu, hit);
            if (m)
            {
                int base = 0;
                if (lane == 0) base = atomicAdd(&s_count, __popc(m));
                base = __shfl_sync(0xffffffffu, base, 0);
                if (hit) s_idx[base + __popc(m & lt)] = i;
            }
        }
        __syncthreads();
        const int nidx = s_count;
        __syncthreads();                                        // every thread has read the counts before the next segment resets them
        if (nidx > 0) any = true;

        // ---- rasterize this bin's triangles, BIN_LIST at a time
        for (int lb = 0; lb < nidx; lb += BIN_LIST)
        {
            const int cnt = min(BIN_LIST, nidx - lb);
            for (int j = tid; j < cnt; j += BIN_THREADS)        // bin-relative edge / depth setup, one triangle per thread
            {
                int i = s_idx[lb + j];
                uint2 bb = __ldg(bboxes + i);
                int x0 = (int)(bb.x & 0xFFFFu), x1 = (int)(bb.x >> 16), y0 = (int)(bb.y & 0xFFFFu), y1 = (int)(bb.y >> 16);
                uint4 ra = __ldg(recs + i * 2), rb = __ldg(recs + i * 2 + 1);
                int ax = s16lo(ra.x), ay = s16hi(ra.x), bx = s16lo(ra.y), by = s16hi(ra.y), cx = s16lo(ra.z), cy = s16hi(ra.z);
                Edge e0 = make_edge(ax, ay, bx, by, binX0, binY0, p.vpW, p.vpH);
                Edge e1 = make_edge(bx, by, cx, cy, binX0, binY0, p.vpW, p.vpH);
                Edge e2 = make_edge(cx, cy, ax, ay, binX0, binY0, p.vpW, p.vpH);
                BinTri t;
                t.e0c = e0.c; t.e0sx = e0.sx; t.e0sy = e0.sy;
                t.e1c = e1.c; t.e1sx = e1.sx; t.e1sy = e1.sy;
                t.e2c = e2.c; t.e2sx = e2.sx; t.e2sy = e2.sy;
                t.zx = ra.w; t.zy = rb.x;
                t.zc = rb.y + ra.w * (uint32_t)binX0 + rb.x * (uint32_t)binY0;
                t.id = rb.z;
                t.box = (uint32_t)(max(x0, binX0) - binX0) | ((uint32_t)(min(x1, binX1) - binX0) << 8) |
                        ((uint32_t)(max(y0, binY0) - binY0) << 16) | ((uint32_t)(min(y1, binY1) - binY0) << 24);
                t.pad0 = 0; t.pad1 = 0;
                s_tri[j] = t;
            }
            __syncthreads();
            // per-warp list: the triangles of this batch whose (bin-clipped) box touches the warp's 32x16 pixels.  Every lane then
            // walks a list ~4-8x shorter than the bin's, with warp-uniform trip count.
            unsigned char* wl = s_wl + warp * BIN_LIST;
            int wn = 0;
            for (int j0 = 0; j0 < cnt; j0 += 32)
            {
                const int k = j0 + lane;
                bool hit = false;
                if (k < cnt)
                {
                    uint32_t box = s_tri[k].box;
                    int x0 = (int)(box & 0xFF), x1 = (int)((box >> 8) & 0xFF), y0 = (int)((box >> 16) & 0xFF), y1 = (int)(box >> 24);
                    hit = !(wx > x1 || wx + 31 < x0 || wy > y1 || wy + 15 < y0);
                }
                const unsigned m = __ballot_sync(0xffffffffu, hit);
                if (hit) wl[wn + __popc(m & lt)] = (unsigned char)k;
                wn += __popc(m);
            }
            __syncwarp();
            for (int q_ = 0; q_ < wn; q_++)
            {
                const int k = wl[q_];
                uint32_t box = s_tri[k].box;
                int x0 = (int)(box & 0xFF), x1 = (int)((box >> 8) & 0xFF), y0 = (int)((box >> 16) & 0xFF), y1 = (int)(box >> 24);
                if (tx > x1 || tx + 3 < x0 || ty > y1 || ty + 3 < y0) continue;
                const int4* q = (const int4*)&s_tri[k];
                int4 q0 = q[0], q1 = q[1], q2 = q[2], q3 = q[3];
                int e0sx = q0.y, e0sy = q0.z, e1sx = q1.x, e1sy = q1.y, e2sx = q1.w, e2sy = q2.x;
                int e0 = q0.x + e0sx * tx + e0sy * ty;
                int e1 = q0.w + e1sx * tx + e1sy * ty;
                int e2 = q1.z + e2sx * tx + e2sy * ty;
                // trivial reject of the 4x4 block: the largest value of an edge function over the block is still negative
                if ((e0 + 3 * max(e0sx, 0) + 3 * max(e0sy, 0)) < 0) continue;
                if ((e1 + 3 * max(e1sx, 0) + 3 * max(e1sy, 0)) < 0) continue;
                if ((e2 + 3 * max(e2sx, 0) + 3 * max(e2sy, 0)) < 0) continue;
                uint32_t zx = (uint32_t)q2.z, zy = (uint32_t)q2.w;
                uint32_t zr = (uint32_t)q2.y + zx * (uint32_t)tx + zy * (uint32_t)ty;
                const uint32_t nid = 0xFFFFFFFFu - (uint32_t)q3.x;
                #pragma unroll
                for (int j = 0; j < 4; j++)
                {
                    int f0 = e0, f1 = e1, f2 = e2;
                    uint32_t z = zr;
                    #pragma unroll
                    for (int i = 0; i < 4; i++)
                    {
                        const int sidx = j * 4 + i;
                        const unsigned long long key = ((unsigned long long)z << 32) | nid;
                        bool pass = ((f0 | f1 | f2) >= 0) && (key < best[sidx]);
                        if (PEEL_IN) pass = pass && (z > peelD[sidx]);
                        if (pass) best[sidx] = key;
                        f0 += e0sx; f1 += e1sx; f2 += e2sx; z += zx;
                    }
                    e0 += e0sy; e1 += e1sy; e2 += e2sy; zr += zy;
                }
            }
            __syncthreads();                                    // s_tri is rebuilt by the next chunk
        }
    }

    // Note: a fragment whose depth equals DR_DEPTH_MAX with id >= 1 wins over the cleared pixel (equal depth passes),
    // fragments deeper than DR_DEPTH_MAX never pass: identical to the reference's clear-depth test.

    // ---- hand the winners to the shader phase through shared memory (transposes 4x4 blocks to row order)
    #pragma unroll
    for (int j = 0; j < 4; j++)
        *(uint4*)&s_ids[(ty + j) * BIN_SIZE + tx] = make_uint4(~(uint32_t)best[j * 4 + 0], ~(uint32_t)best[j * 4 + 1], ~(uint32_t)best[j * 4 + 2], ~(uint32_t)best[j * 4 + 3]);

    if (DEPTH_OUT)
    {
        #pragma unroll
        for (int j = 0; j < 4; j++)
        {
            int ly = binY0 + ty + j, lx = binX0 + tx;
            if (ly >= p.vpH) continue;
            uint32_t* dst = p.depth_out + ((size_t)n * p.H + p.offY + ly) * p.W + p.offX + lx;
            if (lx + 3 < p.vpW && (p.W & 3) == 0 && (p.offX & 3) == 0)
                *(uint4*)dst = make_uint4((uint32_t)(best[j * 4 + 0] >> 32), (uint32_t)(best[j * 4 + 1] >> 32), (uint32_t)(best[j * 4 + 2] >> 32), (uint32_t)(best[j * 4 + 3] >> 32));
            else
            {
                #pragma unroll
                for (int i = 0; i < 4; i++) if (lx + i < p.vpW) dst[i] = (uint32_t)(best[j * 4 + i] >> 32);
            }
        }
    }
    __syncthreads();

    // ---- pixel shader + store.  Warp = 32 consecutive pixels of one row: 512-byte contiguous float4 stores.
    float4* out = (float4*)p.rast + (size_t)n * p.H * p.W;
    float4* odb = (float4*)p.rast_db + (size_t)n * p.H * p.W;
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
    #pragma unroll 4
    for (int k = 0; k < (BIN_SIZE * BIN_SIZE) / BIN_THREADS; k++)
    {
        int q = tid + k * BIN_THREADS;
        int lx = q & (BIN_SIZE - 1), ly = q >> BIN_LOG2;
        if (binX0 + lx >= p.vpW || binY0 + ly >= p.vpH) continue;
        int X = p.offX + binX0 + lx, Y = p.offY + binY0 + ly;                   // output pixel
        float4 o = zero, db = zero;
        if (any)
        {
            int id = (int)s_ids[q] - 1;
            if (id >= 0) shade_pixel(p, n, id, X, Y, o, db);
        }
        size_t pi = (size_t)Y * p.W + X;
        st_stream_f4(out + pi, o);
        st_stream_f4(odb + pi, db);
    }
}

// ================================================================================================
// SCATTER path
// ================================================================================================

template <bool PEEL_IN>
static __device__ __forceinline__ void scatter_fragment(const RastParams& p, int n, int X, int Y, uint32_t depth, uint32_t id)
{
    if (depth > DR_DEPTH_MAX) return;
    size_t pi = ((size_t)n * p.H + p.offY + Y) * p.W + p.offX + X;    // X, Y are viewport-local
    if (PEEL_IN) { if (depth <= __ldg(p.peel_in + pi)) return; }
    unsigned long long key = ((unsigned long long)depth << 32) | (unsigned long long)(0xFFFFFFFFu - id);
    atomicMin(p.keys + pi, key);
}

#define SCATTER_SMALL_AREA 48   // sub-triangles whose bbox has at most this many pixels are walked by their own thread

template <bool PEEL_IN>
struct ScatterEmit
{
    const RastParams& p; int n; uint32_t id;
    // warp-cooperative: lanes that hold a large sub-triangle publish it, the whole warp walks its bbox in 8x4 blocks
    __device__ __forceinline__ void walk_small(const SubTri& s)
    {
        Edge e0 = make_edge(s.x0, s.y0, s.x1, s.y1, s.bx0, s.by0, p.vpW, p.vpH);
        Edge e1 = make_edge(s.x1, s.y1, s.x2, s.y2, s.bx0, s.by0, p.vpW, p.vpH);
        Edge e2 = make_edge(s.x2, s.y2, s.x0, s.y0, s.bx0, s.by0, p.vpW, p.vpH);
        uint32_t zr = s.zb + s.zx * (uint32_t)s.bx0 + s.zy * (uint32_t)s.by0;
        for (int Y = s.by0; Y <= s.by1; Y++)
        {
            int f0 = e0.c, f1 = e1.c, f2 = e2.c; uint32_t z = zr;
            for (int X = s.bx0; X <= s.bx1; X++)
            {
                if ((f0 | f1 | f2) >= 0) scatter_fragment<PEEL_IN>(p, n, X, Y, z, id);
                f0 += e0.sx; f1 += e1.sx; f2 += e2.sx; z += s.zx;
            }
            e0.c += e0.sy; e1.c += e1.sy; e2.c += e2.sy; zr += s.zy;
        }
    }
};

// All 32 lanes call this; lanes with valid==true carry a sub-triangle.  Small ones are walked per lane, large ones by
// the whole warp (8x4 pixel blocks), one after the other.
template <bool PEEL_IN>
static __device__ __forceinline__ void scatter_subtri_warp(const RastParams& p, int n, bool valid, const SubTri& s, uint32_t id)
{
    const int lane = threadIdx.x & 31;
    int area = valid ? (s.bx1 - s.bx0 + 1) * (s.by1 - s.by0 + 1) : 0;
    bool big = area > SCATTER_SMALL_AREA;
    if (valid && !big)
    {
        ScatterEmit<PEEL_IN> e = { p, n, id };
        e.walk_small(s);
    }
    unsigned m = __ballot_sync(0xffffffffu, big);
    while (m)
    {
        int src = __ffs(m) - 1;
        m &= m - 1;
        int x0 = __shfl_sync(0xffffffffu, s.x0, src), y0 = __shfl_sync(0xffffffffu, s.y0, src);
        int x1 = __shfl_sync(0xffffffffu, s.x1, src), y1 = __shfl_sync(0xffffffffu, s.y1, src);
        int x2 = __shfl_sync(0xffffffffu, s.x2, src), y2 = __shfl_sync(0xffffffffu, s.y2, src);
        uint32_t zx = __shfl_sync(0xffffffffu, s.zx, src), zy = __shfl_sync(0xffffffffu, s.zy, src), zb = __shfl_sync(0xffffffffu, s.zb, src);
        int bx0 = __shfl_sync(0xffffffffu, s.bx0, src), bx1 = __shfl_sync(0xffffffffu, s.bx1, src);
        int by0 = __shfl_sync(0xffffffffu, s.by0, src), by1 = __shfl_sync(0xffffffffu, s.by1, src);
        uint32_t tid_ = __shfl_sync(0xffffffffu, id, src);
        int lx = lane & 7, ly = lane >> 3;
        Edge e0 = make_edge(x0, y0, x1, y1, bx0 + lx, by0 + ly, p.vpW, p.vpH);
        Edge e1 = make_edge(x1, y1, x2, y2, bx0 + lx, by0 + ly, p.vpW, p.vpH);
        Edge e2 = make_edge(x2, y2, x0, y0, bx0 + lx, by0 + ly, p.vpW, p.vpH);
        uint32_t zr = zb + zx * (uint32_t)(bx0 + lx) + zy * (uint32_t)(by0 + ly);
        for (int Y = by0 + ly; Y <= by1; Y += 4)
        {
            int f0 = e0.c, f1 = e1.c, f2 = e2.c; uint32_t z = zr;
            for (int X = bx0 + lx; X <= bx1; X += 8)
            {
                if ((f0 | f1 | f2) >= 0) scatter_fragment<PEEL_IN>(p, n, X, Y, z, tid_);
                f0 += 8 * e0.sx; f1 += 8 * e1.sx; f2 += 8 * e2.sx; z += 8u * zx;
            }
            e0.c += 4 * e0.sy; e1.c += 4 * e1.sy; e2.c += 4 * e2.sy; zr += 4u * zy;
        }
    }
}

#ifndef SCATTER_MINB
#define SCATTER_MINB 8      // 64 registers (28 bytes spilled), 8 CTAs/SM: c4 rasterize fwd 0.696 -> 0.673 ms; 10 CTAs (48 registers) 0.692
#endif
template <bool PEEL_IN>
__global__ void __launch_bounds__(128, SCATTER_MINB) raster_scatter_kernel(const RastParams p)
{
    int t = blockIdx.x * 128 + threadIdx.x;     // whole warps stay alive: no early return before the warp collectives
    int n = blockIdx.y;
    int start, count;
    image_range(p, n, start, count);
    int triIdx = start + t;
    bool live = (t < p.max_tris) && (t < count) && ((uint32_t)triIdx < (uint32_t)p.T);

    float4 v0, v1, v2;
    int cls = 0;
    if (live && fetch_triangle(p, n, triIdx, v0, v1, v2)) cls = classify(v0, v1, v2);
    uint32_t id = (uint32_t)(triIdx + 1);

    // fast path: triangles fully inside the frustum
    {
        SubTri s;
        bool valid = (cls == 1) && snap_and_setup(v0, v1, v2, p.vpW, p.vpH, s);
        scatter_subtri_warp<PEEL_IN>(p, n, valid, s, id);
    }
    // rare path: clipped triangles, one fan triangle at a time (warp-uniform trip count)
    if (__any_sync(0xffffffffu, cls == 2))
    {
        float bary[18];
        int nv = 0;
        float4 d1, d2, c0, c1;
        if (cls == 2)
        {
            d1 = make_float4(__fsub_rn(v1.x, v0.x), __fsub_rn(v1.y, v0.y), __fsub_rn(v1.z, v0.z), __fsub_rn(v1.w, v0.w));
            d2 = make_float4(__fsub_rn(v2.x, v0.x), __fsub_rn(v2.y, v0.y), __fsub_rn(v2.z, v0.z), __fsub_rn(v2.w, v0.w));
            nv = clip_frustum(bary, v0, v1, v2, d1, d2);
            c0 = bary_point(v0, d1, d2, bary[0], bary[1]);
            c1 = bary_point(v0, d1, d2, bary[2], bary[3]);
        }
        int nmax = nv;
        #pragma unroll
        for (int d = 16; d > 0; d >>= 1) nmax = max(nmax, __shfl_xor_sync(0xffffffffu, nmax, d));
        for (int i = 2; i < nmax; i++)
        {
            SubTri s;
            bool valid = false;
            if (i < nv)
            {
                float4 c2 = bary_point(v0, d1, d2, bary[i * 2 + 0], bary[i * 2 + 1]);
                valid = snap_and_setup(c0, c1, c2, p.vpW, p.vpH, s);
                c1 = c2;
            }
            scatter_subtri_warp<PEEL_IN>(p, n, valid, s, id);
        }
    }
}

// Resolve pass of the SCATTER path: key -> (id, depth), pixel shader, stores.  The chain key -> triangle indices -> vertex positions
// is three dependent gathers per pixel (ncu, 1 M triangles: long-scoreboard stalls 14 per issue).  A thread can walk RESOLVE_ILP
// pixels at once, each stage issued for all of them before the next starts -- measured on c4 (r02 call 16, rasterize fwd in ms):
// 1 pixel 0.677, 2 pixels 0.696, 4 pixels 0.747: occupancy (32 registers, 80 % of the warp slots) hides the latency better than the
// wider threads do, so the default stays 1.  grid = (blocks, images): 32-bit pixel arithmetic.
#ifndef RESOLVE_ILP
#define RESOLVE_ILP 1
#endif
template <bool DEPTH_OUT>
__global__ void __launch_bounds__(256) raster_resolve_kernel(const RastParams p)
{
    const unsigned hw = (unsigned)p.H * (unsigned)p.W;
    const int n = blockIdx.y;
    const size_t img = (size_t)n * hw;
    const float4* vb = (const float4*)p.pos + (p.instance_mode ? (size_t)n * p.V : 0);
    for (unsigned r0 = blockIdx.x * (256u * RESOLVE_ILP) + threadIdx.x; r0 < hw; r0 += gridDim.x * (256u * RESOLVE_ILP))
    {
        unsigned long long key[RESOLVE_ILP];
        int id[RESOLVE_ILP], vi[RESOLVE_ILP][3];
        float4 v[RESOLVE_ILP][3];
        #pragma unroll
        for (int u = 0; u < RESOLVE_ILP; u++)
        {
            const unsigned r = r0 + u * 256u;
            key[u] = (r < hw) ? p.keys[img + r] : 0xFFFFFFFFFFFFFFFFull;
        }
        #pragma unroll
        for (int u = 0; u < RESOLVE_ILP; u++)
        {
            id[u] = (int)(0xFFFFFFFFu - (uint32_t)key[u]) - 1;
            vi[u][0] = vi[u][1] = vi[u][2] = 0;
            if (id[u] >= 0)
            {
                const int32_t* t = p.tri + (size_t)id[u] * 3;
                vi[u][0] = __ldg(t); vi[u][1] = __ldg(t + 1); vi[u][2] = __ldg(t + 2);
            }
        }
        #pragma unroll
        for (int u = 0; u < RESOLVE_ILP; u++)
            if (id[u] >= 0) { v[u][0] = __ldg(vb + vi[u][0]); v[u][1] = __ldg(vb + vi[u][1]); v[u][2] = __ldg(vb + vi[u][2]); }
        #pragma unroll
        for (int u = 0; u < RESOLVE_ILP; u++)
        {
            const unsigned r = r0 + u * 256u;
            if (r >= hw) continue;
            float4 o = make_float4(0.f, 0.f, 0.f, 0.f), db = o;
            if (id[u] >= 0)
            {
                const unsigned Y = r / (unsigned)p.W, X = r - Y * (unsigned)p.W;
                pm::PixelTri t;
                pm::pixel_tri(t, v[u][0], v[u][1], v[u][2], p.xs * (float)X + p.xo, p.ys * (float)Y + p.yo);
                pm::shade(t, p.xs, p.ys, triidx_to_float(id[u] + 1), o, db);
            }
            st_stream_f4((float4*)p.rast + img + r, o);
            st_stream_f4((float4*)p.rast_db + img + r, db);
            if (DEPTH_OUT) p.depth_out[img + r] = (id[u] >= 0) ? (uint32_t)(key[u] >> 32) : DR_DEPTH_MAX;
        }
    }
}

// ================================================================================================
// Backward (csrc/common/rasterize.cu:119-273)
// ================================================================================================

struct RastGradParams
{
    const float* pos; const int32_t* tri; const float* rast; const float* dy; const float* ddb;
    float* grad;
    int N, V, T, H, W, instance_mode;
    int use_smem;       // 1: accumulate this CTA's image in shared memory (V*4 floats), flush once
    int chunks;         // 256-pixel chunks per image
    float xs, xo, ys, yo;
};

// Gradient of one pixel w.r.t. the clip-space positions of its triangle: replaces the per-pixel part of RasterizeGradKernel[Db]
// (csrc/common/rasterize.cu:119-273); derivation and code in pixel_math.cuh (pm::raster_grad).  g = (x, y, w) per vertex.
template <bool ENABLE_DB>
static __device__ __forceinline__ void raster_grad_pixel(const RastGradParams& p, int n, int px, int py, int vi0, int vi1, int vi2,
                                                         float2 dy, float4 ddb, bool has_ddb, float (&g)[9])
{
    const float4* vb = (const float4*)p.pos + (p.instance_mode ? (size_t)n * p.V : 0);
    const float fx = p.xs * (float)px + p.xo, fy = p.ys * (float)py + p.yo;
    pm::PixelTri t;
    pm::pixel_tri(t, __ldg(vb + vi0), __ldg(vb + vi1), __ldg(vb + vi2), fx, fy);
    float gx[3], gy[3], gw[3];
    pm::raster_grad<ENABLE_DB>(t, fx, fy, p.xs, p.ys, dy.x, dy.y, ddb, has_ddb, gx, gy, gw);
    #pragma unroll
    for (int m = 0; m < 3; m++) { g[3 * m] = gx[m]; g[3 * m + 1] = gy[m]; g[3 * m + 2] = gw[m]; }
}

// Shared-memory gradient accumulators pay a compare-and-swap loop per fp32 add (ATOMS.CAST.SPIN; only 32-bit integer adds are
// native in shared memory): they win only while so few vertices exist that global reductions would collide (c2: 8 vertices).
// Above the threshold every triangle run issues one 128-bit global vector reduction per vertex instead (measured on c3,
// 2 930 vertices x 16 images: 0.724 -> 0.664 ms).
#ifndef RG_SMEM_MAX_V
#define RG_SMEM_MAX_V 1024
#endif
#ifndef RG_UNROLL
#define RG_UNROLL 2
#endif
#ifndef RG_MINB
#define RG_MINB 4
#endif

// Warp-autonomous: a warp owns groups of 32 consecutive pixels of image blockIdx.y, two groups in flight.
template <bool ENABLE_DB>
__global__ void __launch_bounds__(256, RG_MINB) raster_grad_kernel(const RastGradParams p)
{
    extern __shared__ float s_acc[];
    const int n = blockIdx.y;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int hw = p.H * p.W;
    const int groups = (hw + 31) >> 5;
    const int wstride = gridDim.x * 8;
    const size_t img = (size_t)n * hw;
    float* gbase = p.grad + (p.instance_mode ? (size_t)n * p.V * 4 : 0);
    float* acc = p.use_smem ? s_acc : gbase;
    if (p.use_smem)
    {
        for (int i = tid; i < p.V * 4; i += 256) s_acc[i] = 0.f;
        __syncthreads();
    }

    for (int g0 = blockIdx.x * 8 + warp; g0 < groups; g0 += RG_UNROLL * wstride)
    {
        float4 ro[RG_UNROLL], ddb[RG_UNROLL]; float2 dy[RG_UNROLL]; bool inb[RG_UNROLL];
        #pragma unroll
        for (int u = 0; u < RG_UNROLL; u++)
        {
            int r = (g0 + u * wstride) * 32 + lane;
            inb[u] = (g0 + u * wstride < groups) && (r < hw);
            ro[u] = make_float4(0.f, 0.f, 0.f, 0.f); ddb[u] = ro[u]; dy[u] = make_float2(0.f, 0.f);
            if (inb[u]) ro[u] = ld_stream_f4((const float4*)p.rast + img + r);
        }
        // incoming gradients are fetched only under geometry (empty pixels cost 16 instead of 40-48 bytes); the extra
        // dependency runs in parallel with the triangle -> vertex gathers
        #pragma unroll
        for (int u = 0; u < RG_UNROLL; u++)
        {
            int r = (g0 + u * wstride) * 32 + lane;
            if (inb[u] && ro[u].w != 0.f)
            {
                dy[u] = ld_stream_f2((const float2*)p.dy + (img + r) * 2);
                if (ENABLE_DB) ddb[u] = ld_stream_f4((const float4*)p.ddb + img + r);
            }
        }
        #pragma unroll
        for (int u = 0; u < RG_UNROLL; u++)
        {
            int g = g0 + u * wstride;
            if (g >= groups) continue;                      // warp-uniform
            int r = g * 32 + lane;
            int triIdx = inb[u] ? float_to_triidx(ro[u].w) - 1 : -1;
            if (triIdx < 0 || triIdx >= p.T) triIdx = -1;
            int grad_all_dy = __float_as_int(dy[u].x) | __float_as_int(dy[u].y);
            int grad_all_ddb = ENABLE_DB ? (__float_as_int(ddb[u].x) | __float_as_int(ddb[u].y) | __float_as_int(ddb[u].z) | __float_as_int(ddb[u].w)) : 0;
            if (((grad_all_dy | grad_all_ddb) << 1) == 0) triIdx = -1;     // all incoming gradients are +-0
            int vi0 = 0, vi1 = 0, vi2 = 0;
            if (triIdx >= 0)
            {
                vi0 = __ldg(p.tri + triIdx * 3 + 0); vi1 = __ldg(p.tri + triIdx * 3 + 1); vi2 = __ldg(p.tri + triIdx * 3 + 2);
                if (vi0 < 0 || vi0 >= p.V || vi1 < 0 || vi1 >= p.V || vi2 < 0 || vi2 >= p.V) triIdx = -1;
            }
            if (__all_sync(0xffffffffu, triIdx < 0)) continue;

            float gv[9];
            #pragma unroll
            for (int i = 0; i < 9; i++) gv[i] = 0.f;
            if (triIdx >= 0)
            {
                int py = r / p.W, px = r - py * p.W;
                raster_grad_pixel<ENABLE_DB>(p, n, px, py, vi0, vi1, vi2, dy[u], ddb[u], (grad_all_ddb << 1) != 0, gv);
            }
            // one atomic per (run of equal triangles, value) instead of one per pixel
            WarpRuns runs = warp_runs(triIdx);
            #pragma unroll
            for (int i = 0; i < 9; i++) gv[i] = warp_run_sum(gv[i], runs);
            if (runs.is_head && triIdx >= 0)
            {
                if (p.use_smem)     // shared copy: an fp32 add is a compare-and-swap loop (ATOMS.CAST.SPIN), one per component
                {
                    red_add_shared(acc + 4 * vi0 + 0, gv[0]); red_add_shared(acc + 4 * vi0 + 1, gv[1]); red_add_shared(acc + 4 * vi0 + 3, gv[2]);
                    red_add_shared(acc + 4 * vi1 + 0, gv[3]); red_add_shared(acc + 4 * vi1 + 1, gv[4]); red_add_shared(acc + 4 * vi1 + 3, gv[5]);
                    red_add_shared(acc + 4 * vi2 + 0, gv[6]); red_add_shared(acc + 4 * vi2 + 1, gv[7]); red_add_shared(acc + 4 * vi2 + 3, gv[8]);
                }
                else                // global: one 128-bit vector reduction per vertex (z receives +0)
                {
                    red_add4(acc + 4 * vi0, make_float4(gv[0], gv[1], 0.f, gv[2]));
                    red_add4(acc + 4 * vi1, make_float4(gv[3], gv[4], 0.f, gv[5]));
                    red_add4(acc + 4 * vi2, make_float4(gv[6], gv[7], 0.f, gv[8]));
                }
            }
        }
    }

    if (p.use_smem)
    {
        __syncthreads();
        for (int i = tid; i < p.V * 4; i += 256)
        {
            float v = s_acc[i];
            if (v != 0.f) red_add_global(gbase + i, v);
        }
    }
}

} // namespace

// ================================================================================================
// Host entry points
// ================================================================================================

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

extern "C" int b200dr_rasterize_pick_path(int N, int max_tris, int H, int W)
{
    // Measured sweep at 2048^2 x 16 images (scripts/tri_sweep.py, profiles/r02_tri_sweep.json; ms per launch, BINS / SCATTER):
    //   12 tris 0.31 / 13.9   1 k 0.40 / 2.6   5.9 k 0.55 / 1.4   12 k 0.68 / 1.4   30 k 1.06 / 1.46   100 k 2.5 / 1.9   1 M 17.7 / 1.27
    // BINS (CTA per 64x64 bin over chunked bounding boxes) costs grow with the triangles per bin, SCATTER (fused setup + 64-bit
    // atomicMin) gets cheaper per triangle as triangles shrink: they cross between 30 k and 100 k triangles per image.
    (void)N; (void)H; (void)W;
    return (max_tris <= B200DR_BINS_MAX_TRIS) ? B200DR_RAST_BINS : B200DR_RAST_SCATTER;
}

extern "C" size_t b200dr_rasterize_scratch_bytes(int N, int max_tris, int H, int W, int path)
{
    if (path == B200DR_RAST_AUTO) path = b200dr_rasterize_pick_path(N, max_tris, H, W);
    if (path == B200DR_RAST_BINS)
    {
        size_t cap = ((size_t)7 * max_tris + 1) & ~(size_t)1;
        size_t nchunk = ((((size_t)max_tris + 31) >> 5) + 1) & ~(size_t)1;   // even: every image's chunk-box array stays a 16-byte aligned TMA source
        return align_up((size_t)N * (1 + BIN_ROWS) * sizeof(int), 256) + align_up((size_t)N * nchunk * sizeof(uint2), 256) +
               align_up((size_t)N * cap * sizeof(uint2), 256) + (size_t)N * cap * 2 * sizeof(uint4);
    }
    return (size_t)N * H * W * sizeof(unsigned long long);
}

extern "C" int b200dr_rasterize_fwd(const float* pos, const int32_t* tri, const int32_t* ranges_dev,
                                    int N, int V, int T, int H, int W, int instance_mode, int max_tris,
                                    const uint32_t* peel_depth_in, uint32_t* depth_out,
                                    void* scratch, size_t scratch_bytes, int path,
                                    float* rast, float* rast_db, void* stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    B200DR_CHECK(N > 0 && V > 0 && T > 0, "rasterize_fwd(): pos/tri must be non-empty");
    B200DR_CHECK(H > 0 && W > 0, "rasterize_fwd(): resolution must be [>0, >0]");
    B200DR_CHECK(H <= 32768 && W <= 32768, "rasterize_fwd(): resolution too large");
    B200DR_CHECK(instance_mode || ranges_dev, "rasterize_fwd(): range mode requires ranges");
    B200DR_CHECK(max_tris > 0 && max_tris < (1 << 24), "rasterize_fwd(): triangle count per image must be in [1, 2^24)");
    B200DR_CHECK(B200DR_ALIGNED16(pos), "rasterize_fwd(): pos input tensor not aligned to float4");
    B200DR_CHECK(B200DR_ALIGNED16(rast), "rasterize_fwd(): out output tensor not aligned to float4");
    B200DR_CHECK(B200DR_ALIGNED16(rast_db), "rasterize_fwd(): out_db output tensor not aligned to float4");
    if (path == B200DR_RAST_AUTO) path = b200dr_rasterize_pick_path(N, max_tris, H, W);
    B200DR_CHECK(path == B200DR_RAST_BINS || path == B200DR_RAST_SCATTER, "rasterize_fwd(): invalid path");
    B200DR_CHECK(scratch && scratch_bytes >= b200dr_rasterize_scratch_bytes(N, max_tris, H, W, path), "rasterize_fwd(): scratch buffer too small");
    B200DR_CHECK(B200DR_ALIGNED16(scratch), "rasterize_fwd(): scratch buffer not 16-byte aligned");      // TMA bulk copies read the bbox arrays
    B200DR_CHECK(N <= 65535, "rasterize_fwd(): minibatch size above 65535 is not supported");
    B200DR_CHECK(!peel_depth_in || depth_out, "rasterize_fwd(): peeling needs a depth output buffer");

    RastParams p = {};
    p.pos = pos; p.tri = tri; p.ranges = ranges_dev;
    p.N = N; p.V = V; p.T = T; p.H = H; p.W = W; p.instance_mode = instance_mode; p.max_tris = max_tris;
    p.cap = (7 * max_tris + 1) & ~1;
    p.peel_in = peel_depth_in; p.depth_out = depth_out;
    p.rast = rast; p.rast_db = rast_db;
    p.xs = 2.f / (float)W; p.xo = 1.f / (float)W - 1.f;
    p.ys = 2.f / (float)H; p.yo = 1.f / (float)H - 1.f;

    if (path == B200DR_RAST_BINS)
    {
        char* s = (char*)scratch;
        p.nchunk = (((max_tris + 31) >> 5) + 1) & ~1;
        p.extra = (int*)s; p.binmask = (uint32_t*)s + N;       // cleared together, once per viewport
        s += align_up((size_t)N * (1 + BIN_ROWS) * sizeof(int), 256);
        p.cbox = (uint2*)s;                   s += align_up((size_t)N * p.nchunk * sizeof(uint2), 256);
        p.bbox = (uint2*)s;                   s += align_up((size_t)N * p.cap * sizeof(uint2), 256);
        p.rec = (uint4*)s;
    }
    else
    {
        p.keys = (unsigned long long*)scratch;
        B200DR_CUDA(cudaMemsetAsync(p.keys, 0xFF, (size_t)N * H * W * sizeof(unsigned long long), stream), "rasterize()");
    }

    // Viewports larger than 2048 px are rendered as a grid of <= 2048^2 sub-viewports (the snapped coordinates are S16):
    // same tiling and vertex remap as the reference (csrc/torch/torch_rasterize.cpp:98-124, impl/RasterImpl.cpp:295-298).
    const int MAXVP = 2048;
    int width8 = (W + 7) & ~7, height8 = (H + 7) & ~7;
    int tileCountX = (width8 + MAXVP - 1) / MAXVP, tileCountY = (height8 + MAXVP - 1) / MAXVP;
    int tileSizeX = (((width8 + tileCountX - 1) / tileCountX) + 7) & ~7;
    int tileSizeY = (((height8 + tileCountY - 1) / tileCountY) + 7) & ~7;
    for (int tileY = 0; tileY < tileCountY; tileY++)
    for (int tileX = 0; tileX < tileCountX; tileX++)
    {
        p.offX = tileX * tileSizeX; p.offY = tileY * tileSizeY;
        p.vpW = (W - p.offX) < tileSizeX ? (W - p.offX) : tileSizeX;
        p.vpH = (H - p.offY) < tileSizeY ? (H - p.offY) : tileSizeY;
        if (p.vpW <= 0 || p.vpH <= 0) continue;
        p.vxs = (float)W / (float)p.vpW; p.vys = (float)H / (float)p.vpH;
        p.vxo = (float)(W - p.vpW - 2 * p.offX) / (float)p.vpW;
        p.vyo = (float)(H - p.vpH - 2 * p.offY) / (float)p.vpH;
        if (path == B200DR_RAST_BINS)
        {
            B200DR_CUDA(cudaMemsetAsync(p.extra, 0, (size_t)N * (1 + BIN_ROWS) * sizeof(int), stream), "rasterize()");
            raster_setup_kernel<<<dim3((max_tris + 127) / 128, N), 128, 0, stream>>>(p);
            B200DR_CHECK_LAUNCH("rasterize_fwd(setup)");
            dim3 grid((p.vpW + BIN_SIZE - 1) / BIN_SIZE, (p.vpH + BIN_SIZE - 1) / BIN_SIZE, N);
            if (peel_depth_in)  raster_bins_kernel<true, true><<<grid, BIN_THREADS, 0, stream>>>(p);
            else if (depth_out) raster_bins_kernel<false, true><<<grid, BIN_THREADS, 0, stream>>>(p);
            else                raster_bins_kernel<false, false><<<grid, BIN_THREADS, 0, stream>>>(p);
            B200DR_CHECK_LAUNCH("rasterize_fwd(bins)");
        }
        else
        {
            dim3 grid((max_tris + 127) / 128, N);
            if (peel_depth_in) raster_scatter_kernel<true><<<grid, 128, 0, stream>>>(p);
            else               raster_scatter_kernel<false><<<grid, 128, 0, stream>>>(p);
            B200DR_CHECK_LAUNCH("rasterize_fwd(scatter)");
        }
    }
    if (path == B200DR_RAST_SCATTER)
    {
        size_t per_image = (size_t)H * W;
        int blocks = (int)((per_image + 256 * RESOLVE_ILP - 1) / (256 * RESOLVE_ILP));
        int maxb = (b200dr_num_sms() * 16 + N - 1) / N;
        if (blocks > maxb) blocks = maxb;
        if (blocks < 1) blocks = 1;
        if (depth_out) raster_resolve_kernel<true><<<dim3(blocks, N), 256, 0, stream>>>(p);
        else           raster_resolve_kernel<false><<<dim3(blocks, N), 256, 0, stream>>>(p);
        B200DR_CHECK_LAUNCH("rasterize_fwd(resolve)");
    }
    return 0;
}

extern "C" int b200dr_rasterize_grad(const float* pos, const int32_t* tri, const float* rast, const float* dy, const float* ddb,
                                     int N, int V, int T, int H, int W, int instance_mode,
                                     float* grad_pos, void* stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    B200DR_CHECK(N > 0 && H > 0 && W > 0, "rasterize_grad(): resolution must be [>0, >0, >0]");
    B200DR_CHECK(V > 0 && T > 0, "rasterize_grad(): pos/tri must be non-empty");
    B200DR_CHECK(B200DR_ALIGNED16(pos), "rasterize_grad(): pos input tensor not aligned to float4");
    B200DR_CHECK(B200DR_ALIGNED16(rast), "rasterize_grad(): out input tensor not aligned to float4");
    B200DR_CHECK((((uintptr_t)dy) & 7) == 0, "rasterize_grad(): dy input tensor not aligned to float2");
    B200DR_CHECK(B200DR_ALIGNED16(ddb), "rasterize_grad(): ddb input tensor not aligned to float4");
    B200DR_CHECK(N <= 65535, "rasterize_grad(): minibatch size above 65535 is not supported");
    B200DR_CHECK((size_t)H * W < (1u << 30), "rasterize_grad(): image too large");

    RastGradParams p = {};
    p.pos = pos; p.tri = tri; p.rast = rast; p.dy = dy; p.ddb = ddb; p.grad = grad_pos;
    p.N = N; p.V = V; p.T = T; p.H = H; p.W = W; p.instance_mode = instance_mode;
    p.xs = 2.f / (float)W; p.xo = 1.f / (float)W - 1.f;
    p.ys = 2.f / (float)H; p.yo = 1.f / (float)H - 1.f;
    size_t hw = (size_t)H * W;
    p.chunks = (int)((hw + 255) / 256);
    size_t smem = (size_t)V * 4 * sizeof(float);
    p.use_smem = (smem <= 64 * 1024 && V <= RG_SMEM_MAX_V) ? 1 : 0;
    // Persistent-ish grid: enough CTAs to fill the machine, few enough that the shared-memory flush stays negligible.
    int target = b200dr_num_sms() * RG_MINB;
    int gx = p.use_smem ? max(target / N, 1) : 16384;       // shared accumulators: one resident wave, rounded down
    int max_gx = (p.chunks + RG_UNROLL - 1) / RG_UNROLL;        // 256-pixel chunks = 8 warp groups per CTA pass
    if (gx > max_gx) gx = max_gx;
    if (gx < 1) gx = 1;
    dim3 grid(gx, N);
    size_t dyn = p.use_smem ? smem : 0;
    if (ddb)
    {
        if (dyn > 48 * 1024) B200DR_CUDA(cudaFuncSetAttribute(raster_grad_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024), "rasterize()");
        raster_grad_kernel<true><<<grid, 256, dyn, stream>>>(p);
    }
    else
    {
        if (dyn > 48 * 1024) B200DR_CUDA(cudaFuncSetAttribute(raster_grad_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024), "rasterize()");
        raster_grad_kernel<false><<<grid, 256, dyn, stream>>>(p);
    }
    B200DR_CHECK_LAUNCH("rasterize_grad()");
    return 0;
}
