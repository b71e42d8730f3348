// texture.cu -- B200-native replacement of the reference's texture op, 2-D modes:
//   texture_construct_mip  (csrc/torch/torch_texture.cpp:98-169,  csrc/common/texture_kernel.cu:644-704)
//   texture_fwd[_mip]      (csrc/torch/torch_texture.cpp:174-416, csrc/common/texture_kernel.cu:709-800)
//   texture_grad_*         (csrc/torch/torch_texture.cpp:421-716, csrc/common/texture_kernel.cu:843-1154)
//   mip layout             (csrc/common/texture.cpp:62-102)
//
// One kernel per direction handles every 2-D filter/boundary mode (mode switches are warp-uniform), instead of the
// reference's 48-entry kernel table.  Per-pixel tensors are staged through shared memory for full-line accesses.
// Texels are read from a channel-padded pyramid (Cp = C rounded up to 4 floats per texel, level 0 included) so that a
// bilinear tap is one 128-bit load per corner instead of C scalar gathers; the front end pads level 0 per call
// (28 bytes per texel, noise next to the per-pixel traffic) and builds / caches the padded mip levels.
// Backward: texel gradients are accumulated into a channel-padded gradient pyramid with 128-bit vector reductions
// (REDG.E.ADD.F32x4: one atomic per texel instead of one per texel channel), then a resolve pass pulls the mip levels
// down to the base level and writes the gradient in the caller's channel layout.
#include "common.cuh"
#include "pixel_math.cuh"
#include <stdlib.h>

namespace {

#define TEX_THREADS DR_THREADS
#define TEX_MAX_MIP B200DR_TEX_MAX_MIP_LEVEL

struct TexParams
{
    const float* tex[TEX_MAX_MIP + 1];      // level pointers, [0] = base texture
    float*       gpad[TEX_MAX_MIP + 1];     // backward: channel-padded gradient pyramid (Cp floats per texel)
    const float* uv; const float* uv_da; const float* bias; const float* dy;
    float* out; float* g_uv; float* g_uv_da; float* g_bias;
    int N, H, W, texD, texH, texW, C, Cp;
    int filter, boundary, mip_max;
    int staged;
    int chunks;                             // 256-pixel chunks over the flat N*H*W pixel array
    int tile2d, tilesX, tilesPerImg;        // TEXG_TILE2D experiment: a warp's 32 pixels form an 8x4 tile instead of a row segment
};

static __device__ __forceinline__ int lvl_w(const TexParams& p, int l) { int v = p.texW >> l; return v > 1 ? v : 1; }
static __device__ __forceinline__ int lvl_h(const TexParams& p, int l) { int v = p.texH >> l; return v > 1 ? v : 1; }

// Filter / boundary mode and optional-input flags as seen by the device helpers.  The generic kernels fill it from the
// launch parameters; the specialised fast kernels from template constants, so every mode switch below folds away.
struct TexMode { int filter, boundary; bool has_da, has_bias; };

static __device__ __forceinline__ TexMode tex_mode(const TexParams& p)
{
    TexMode m; m.filter = p.filter; m.boundary = p.boundary; m.has_da = (p.uv_da != 0); m.has_bias = (p.bias != 0);
    return m;
}
template <int F, int B>
static __device__ __forceinline__ TexMode tex_mode_static(const TexParams& p)
{
    TexMode m = tex_mode(p);
    if (F >= 0) m.filter = F;
    if (B >= 0) m.boundary = B;
    return m;
}

// texture_kernel.cu:321-366 (2-D)
static __device__ __forceinline__ int index_nearest(const TexParams& p, const TexMode& m, float u, float v, int tz)
{
    int w = p.texW, h = p.texH;
    if (m.boundary == B200DR_BOUNDARY_WRAP) { u = u - (float)__float2int_rd(u); v = v - (float)__float2int_rd(v); }
    u = u * (float)w; v = v * (float)h;
    int iu = __float2int_rd(u), iv = __float2int_rd(v);
    if (m.boundary == B200DR_BOUNDARY_ZERO && (iu < 0 || iu >= w || iv < 0 || iv >= h)) return -1;
    iu = min(max(iu, 0), w - 1); iv = min(max(iv, 0), h - 1);
    return iu + w * (iv + tz * h);
}

// texture_kernel.cu:368-472 (2-D): four texel indices (-1 = outside in zero mode) and the bilinear weights.
static __device__ __forceinline__ float2 index_linear(const TexParams& p, const TexMode& m, float u, float v, int tz, int level, int4& tc)
{
    int w = lvl_w(p, level), h = lvl_h(p, level);
    bool clampU = false, clampV = false;
    if (m.boundary == B200DR_BOUNDARY_WRAP) { u = u - (float)__float2int_rd(u); v = v - (float)__float2int_rd(v); }
    u = u * (float)w - 0.5f; v = v * (float)h - 0.5f;
    if (m.boundary == B200DR_BOUNDARY_CLAMP)
    {
        u = fminf(fmaxf(u, 0.f), w - 1.f); v = fminf(fmaxf(v, 0.f), h - 1.f);
        clampU = (u == 0.f || u == w - 1.f); clampV = (v == 0.f || v == h - 1.f);
    }
    int iu0 = __float2int_rd(u), iv0 = __float2int_rd(v);
    int iu1 = iu0 + (clampU ? 0 : 1), iv1 = iv0 + (clampV ? 0 : 1);
    u -= (float)iu0; v -= (float)iv0;
    if (m.boundary == B200DR_BOUNDARY_WRAP)
    {
        if (iu0 < 0) iu0 += w;
        if (iv0 < 0) iv0 += h;
        if (iu1 >= w) iu1 -= w;
        if (iv1 >= h) iv1 -= h;
    }
    int iu0z = iu0 + tz * w * h, iu1z = iu1 + tz * w * h;
    tc.x = iu0z + w * iv0; tc.y = iu1z + w * iv0; tc.z = iu0z + w * iv1; tc.w = iu1z + w * iv1;
    if (m.boundary == B200DR_BOUNDARY_ZERO)
    {
        bool u0o = (iu0 < 0 || iu0 >= w), u1o = (iu1 < 0 || iu1 >= w), v0o = (iv0 < 0 || iv0 >= h), v1o = (iv1 < 0 || iv1 >= h);
        if (u0o || v0o) tc.x = -1;
        if (u1o || v0o) tc.y = -1;
        if (u0o || v1o) tc.z = -1;
        if (u1o || v1o) tc.w = -1;
    }
    return make_float2(u, v);
}

// texture_kernel.cu:477-585 (2-D).  WANT_DW: also the derivative of the (unclamped) level w.r.t. uv_da.
template <bool WANT_DW>
static __device__ __forceinline__ void mip_level_from(const TexParams& p, const TexMode& m, float4 q, float bias, int& level0, int& level1, float& flevel, float4& dw);

template <bool WANT_DW>
static __device__ __forceinline__ void mip_level(const TexParams& p, const TexMode& m, size_t pidx, int& level0, int& level1, float& flevel, float4& dw)
{
    level0 = 0; level1 = 0; flevel = 0.f;
    if (m.filter == B200DR_FILTER_NEAREST || m.filter == B200DR_FILTER_LINEAR) return;
    float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
    if (m.has_da) q = ld_stream_f4((const float4*)p.uv_da + pidx);
    float bias = m.has_bias ? ld_stream_f(p.bias + pidx) : 0.f;
    mip_level_from<WANT_DW>(p, m, q, bias, level0, level1, flevel, dw);
}

// same, with the pixel's uv_da / bias values already in registers
template <bool WANT_DW>
static __device__ __forceinline__ void mip_level_from(const TexParams& p, const TexMode& m, float4 q, float bias, int& level0, int& level1, float& flevel, float4& dw)
{
    level0 = 0; level1 = 0; flevel = 0.f;
    if (m.filter == B200DR_FILTER_NEAREST || m.filter == B200DR_FILTER_LINEAR) return;
    float fl = 0.f;
    float uscl = 0.f, vscl = 0.f, dsdx = 0.f, dsdy = 0.f, dtdx = 0.f, dtdy = 0.f, A = 0.f, B = 0.f, Cc = 0.f, l2a = 0.f, l2b = 0.f, l2n = 0.f;
    if (m.has_da)
    {
        uscl = (float)p.texW; vscl = (float)p.texH;
        dsdx = q.x * uscl; dsdy = q.y * uscl; dtdx = q.z * vscl; dtdy = q.w * vscl;
        A = dsdx * dsdx + dtdx * dtdx;
        B = dsdy * dsdy + dtdy * dtdy;
        Cc = dsdx * dsdy + dtdx * dtdy;
        l2b = 0.5f * (A + B);
        l2n = 0.25f * (A - B) * (A - B) + Cc * Cc;
        l2a = sqrtf(l2n);
        fl = .5f * __log2f(l2b + l2a);
    }
    if (m.has_bias) fl += bias;
    fl = fminf(fmaxf(fl, 0.f), (float)p.mip_max);
    level0 = __float2int_rd(fl);
    if (m.filter == B200DR_FILTER_LINEAR_MIPMAP_LINEAR && fl > 0.f)
    {
        level1 = min(level0 + 1, p.mip_max);
        fl -= (float)level0;
    }
    flevel = fl;
    // d(level)/d(uv_da) only ever multiplies the difference between the two levels' samples, which exists only when the
    // pixel blends two levels (flevel > 0): skipped for magnified / single-level pixels.
    if (WANT_DW && m.has_da && m.filter == B200DR_FILTER_LINEAR_MIPMAP_LINEAR && flevel > 0.f)
    {
        float d = __fdividef(0.72134752f, l2n + l2a * l2b);      // 0.5 / ln 2; gradient-only quantity: approximate division (2 ulp)
        float AB = d * .5f * (A - B), Cw = d * Cc, l2aw = d * l2a;
        float4 r = make_float4(uscl * (dsdx * (l2aw + AB) + dsdy * Cw), uscl * (dsdy * (l2aw - AB) + dsdx * Cw),
                               vscl * (dtdx * (l2aw + AB) + dtdy * Cw), vscl * (dtdy * (l2aw - AB) + dtdx * Cw));
        bool fin = isfinite(r.x) && isfinite(r.y) && isfinite(r.z) && isfinite(r.w);
        dw = fin ? r : make_float4(0.f, 0.f, 0.f, 0.f);
    }
}

// one 128-bit load: channels [k, k+4) of texel tc in a Cp-padded level (zeros for tc < 0)
static __device__ __forceinline__ float4 texel4(const float* t, int tc, int Cp, int k)
{
    return tc >= 0 ? __ldg((const float4*)(t + (size_t)tc * Cp + k)) : make_float4(0.f, 0.f, 0.f, 0.f);
}
static __device__ __forceinline__ float lerpf(float a, float b, float c) { return a + c * (b - a); }
static __device__ __forceinline__ float bilerp(float a, float b, float c, float d, float2 e) { return lerpf(lerpf(a, b, e.x), lerpf(c, d, e.x), e.y); }
static __device__ __forceinline__ float4 bilerp4(float4 a, float4 b, float4 c, float4 d, float2 e)
{
    return make_float4(bilerp(a.x, b.x, c.x, d.x, e), bilerp(a.y, b.y, c.y, d.y, e), bilerp(a.z, b.z, c.z, d.z, e), bilerp(a.w, b.w, c.w, d.w, e));
}
static __device__ __forceinline__ void put4(float* o, int k, int C, float4 v)
{
    o[k] = v.x;
    if (k + 1 < C) o[k + 1] = v.y;
    if (k + 2 < C) o[k + 2] = v.z;
    if (k + 3 < C) o[k + 3] = v.w;
}

// ------------------------------------------------------------------------------------------------
// Mip construction (texture_kernel.cu:644-704): level L from level L-1.

__global__ void __launch_bounds__(256) mip_build_kernel(const float* __restrict__ in, float* __restrict__ out, int wi, int hi, int wo, int ho, int D, int C)
{
    // C is the padded channel count (multiple of 4): one thread averages one 4-channel group of one output texel
    const int G = C >> 2;
    size_t total = (size_t)wo * ho * D * G;
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (size_t)gridDim.x * 256)
    {
        int g = (int)(i % G); size_t t = i / G;
        int x = (int)(t % wo); t /= wo;
        int y = (int)(t % ho); int z = (int)(t / ho);
        size_t i0 = (size_t)C * ((((size_t)x + (size_t)wi * y) << 1) + (size_t)z * wi * hi) + 4 * g;
        float4 v;
        if (wi == 1 || hi == 1)
        {
            size_t i1 = (hi == 1) ? i0 + C : i0 + (size_t)C * wi;
            float4 a = *(const float4*)(in + i0), b = *(const float4*)(in + i1);
            v = make_float4(.5f * (a.x + b.x), .5f * (a.y + b.y), .5f * (a.z + b.z), .5f * (a.w + b.w));
        }
        else
        {
            size_t i1 = i0 + (size_t)C * wi;
            float4 a = *(const float4*)(in + i0), b = *(const float4*)(in + i0 + C), c = *(const float4*)(in + i1), d = *(const float4*)(in + i1 + C);
            v = make_float4(.25f * (a.x + b.x + c.x + d.x), .25f * (a.y + b.y + c.y + d.y), .25f * (a.z + b.z + c.z + d.z), .25f * (a.w + b.w + c.w + d.w));
        }
        *(float4*)(out + (size_t)C * ((size_t)x + (size_t)wo * (y + (size_t)ho * z)) + 4 * g) = v;
    }
}

// dst[t][0..Cp) = src[t][0..C) padded with zeros
__global__ void __launch_bounds__(256) texture_pad_kernel(const float* __restrict__ src, float* __restrict__ dst, size_t ntexels, int C, int Cp)
{
    const int G = Cp >> 2;
    size_t total = ntexels * G;
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (size_t)gridDim.x * 256)
    {
        size_t t = i / G; int k = (int)(i % G) * 4;
        const float* q = src + t * C + k;
        float4 v;
        v.x = __ldg(q);
        v.y = (k + 1 < C) ? __ldg(q + 1) : 0.f;
        v.z = (k + 2 < C) ? __ldg(q + 2) : 0.f;
        v.w = (k + 3 < C) ? __ldg(q + 3) : 0.f;
        *(float4*)(dst + t * Cp + k) = v;
    }
}

// ------------------------------------------------------------------------------------------------
// Forward

__global__ void __launch_bounds__(TEX_THREADS) texture_fwd_kernel(const TexParams p)
{
    const TexMode m = tex_mode(p);
    extern __shared__ float s_out[];
    const int tid = threadIdx.x;
    const int C = p.C;
    const size_t hw = (size_t)p.H * p.W;
    const size_t total = hw * p.N;

    for (int c = blockIdx.x; c < p.chunks; c += gridDim.x)
    {
        size_t pbase = (size_t)c * TEX_THREADS;
        int npx = (int)min((size_t)TEX_THREADS, total - pbase);
        size_t pidx = pbase + tid;
        if (tid < npx)
        {
            int pz = (int)(pidx / hw);
            int tz = (p.texD == 1) ? 0 : pz;
            float2 uv = ld_stream_f2((const float2*)p.uv + pidx);
            float* o = p.staged ? (s_out + tid * C) : (p.out + pidx * C);
            if (m.filter == B200DR_FILTER_NEAREST)
            {
                int tc = index_nearest(p, m, uv.x, uv.y, tz);
                for (int k = 0; k < p.Cp; k += 4) put4(o, k, C, texel4(p.tex[0], tc, p.Cp, k));
            }
            else
            {
                int level0, level1; float flevel; float4 dw;
                mip_level<false>(p, m, pidx, level0, level1, flevel, dw);
                int4 tc0;
                float2 uv0 = index_linear(p, m, uv.x, uv.y, tz, level0, tc0);
                const float* t0 = p.tex[level0];
                const int Cp = p.Cp;
                if (m.filter != B200DR_FILTER_LINEAR_MIPMAP_LINEAR || !(flevel > 0.f))
                {
                    for (int k = 0; k < Cp; k += 4)
                        put4(o, k, C, bilerp4(texel4(t0, tc0.x, Cp, k), texel4(t0, tc0.y, Cp, k), texel4(t0, tc0.z, Cp, k), texel4(t0, tc0.w, Cp, k), uv0));
                }
                else
                {
                    int4 tc1;
                    float2 uv1 = index_linear(p, m, uv.x, uv.y, tz, level1, tc1);
                    const float* t1 = p.tex[level1];
                    for (int k = 0; k < Cp; k += 4)
                    {
                        float4 a = bilerp4(texel4(t0, tc0.x, Cp, k), texel4(t0, tc0.y, Cp, k), texel4(t0, tc0.z, Cp, k), texel4(t0, tc0.w, Cp, k), uv0);
                        float4 b = bilerp4(texel4(t1, tc1.x, Cp, k), texel4(t1, tc1.y, Cp, k), texel4(t1, tc1.z, Cp, k), texel4(t1, tc1.w, Cp, k), uv1);
                        put4(o, k, C, make_float4(lerpf(a.x, b.x, flevel), lerpf(a.y, b.y, flevel), lerpf(a.z, b.z, flevel), lerpf(a.w, b.w, flevel)));
                    }
                }
            }
        }
        if (p.staged)
        {
            __syncthreads();
            coop_store(p.out + pbase * C, s_out, npx * C, tid);
            __syncthreads();
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Backward

// Warp-cooperative texel-gradient accumulation.  Every lane of the warp calls this (inactive lanes pass tc = -1).
// reduce == false: every lane issues its own 128-bit vector reduction (REDG.E.ADD.F32x4); measured 150-400 G texel
//   updates/s on B200 as long as addresses are spread (scripts/atomics_bench.cu).
// reduce == true : lanes that hit the same texel of the same level in consecutive pixels (strongly magnified textures,
//   constant uv such as the background) are summed with shuffles first and only the run head issues the reduction;
//   without this, 32 pixels on one texel are 32 serialised same-address atomics (0.85 clk each, chip-wide).
// hot: per-CTA one-entry cache per bilinear tap for texels that a whole warp hits at once (constant-uv regions): such
//   warps add into shared memory and the CTA flushes once, instead of one same-address global reduction per warp.
#define TEX_HOT_SLOTS 4
#define TEX_HOT_MAXC  16
struct HotCache { int key[TEX_HOT_SLOTS]; float val[TEX_HOT_SLOTS][TEX_HOT_MAXC]; };

static __device__ __forceinline__ void accum_texel_warp(float* g, int level, int tc, int C, int Cp, float w, const float* dy, bool reduce,
                                                        HotCache* hot = 0, int slot = -1)
{
    float* q = g + (size_t)(tc >= 0 ? tc : 0) * Cp;
    if (!reduce)
    {
        if (tc < 0) return;
        for (int k = 0; k < Cp; k += 4)
            red_add4(q + k, make_float4(w * dy[k], (k + 1 < C) ? w * dy[k + 1] : 0.f, (k + 2 < C) ? w * dy[k + 2] : 0.f, (k + 3 < C) ? w * dy[k + 3] : 0.f));
        return;
    }
    int key = (tc >= 0) ? (tc ^ (level << 27)) : -1;
    WarpRuns runs = warp_runs(key);
    for (int k = 0; k < Cp; k += 4)
    {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (tc >= 0)
        {
            v.x = w * dy[k];
            v.y = (k + 1 < C) ? w * dy[k + 1] : 0.f;
            v.z = (k + 2 < C) ? w * dy[k + 2] : 0.f;
            v.w = (k + 3 < C) ? w * dy[k + 3] : 0.f;
        }
        v.x = warp_run_sum(v.x, runs);
        if (k + 1 < C) v.y = warp_run_sum(v.y, runs);
        if (k + 2 < C) v.z = warp_run_sum(v.z, runs);
        if (k + 3 < C) v.w = warp_run_sum(v.w, runs);
        if (runs.is_head && tc >= 0)
        {
            bool cached = false;
            if (hot && runs.heads == 1u)                    // the whole warp sits on this texel
            {
                int old = (k == 0) ? atomicCAS(&hot->key[slot], -1, key) : hot->key[slot];
                if (old == -1 || old == key)
                {
                    atomicAdd(&hot->val[slot][k + 0], v.x); atomicAdd(&hot->val[slot][k + 1], v.y);
                    atomicAdd(&hot->val[slot][k + 2], v.z); atomicAdd(&hot->val[slot][k + 3], v.w);
                    cached = true;
                }
            }
            if (!cached) red_add4(q + k, v);
        }
    }
}

__global__ void __launch_bounds__(TEX_THREADS) texture_grad_kernel(const TexParams p)
{
    const TexMode m = tex_mode(p);
    extern __shared__ float s_dy[];
    __shared__ HotCache s_hot;
    const int tid = threadIdx.x;
    const int C = p.C, Cp = p.Cp;
    const size_t hw = (size_t)p.H * p.W;
    const size_t total = hw * p.N;
    const bool trilinear = (m.filter == B200DR_FILTER_LINEAR_MIPMAP_LINEAR);
    HotCache* hot = (Cp <= TEX_HOT_MAXC) ? &s_hot : 0;
    if (tid < TEX_HOT_SLOTS) s_hot.key[tid] = -1;
    if (tid < TEX_HOT_SLOTS * TEX_HOT_MAXC) s_hot.val[tid / TEX_HOT_MAXC][tid % TEX_HOT_MAXC] = 0.f;
    __syncthreads();

    for (int c = blockIdx.x; c < p.chunks; c += gridDim.x)
    {
        size_t pbase = (size_t)c * TEX_THREADS;
        int npx = (int)min((size_t)TEX_THREADS, total - pbase);
        size_t pidx = pbase + tid;
        if (p.staged)
        {
            coop_load(s_dy, p.dy + pbase * C, npx * C, tid);
            __syncthreads();
        }
        const bool inb = tid < npx;
        const float* pDy = p.staged ? (s_dy + tid * C) : (p.dy + (inb ? pidx : pbase) * C);

        // ---- per-pixel setup (lanes without work keep tc = -1 and flow through the warp-cooperative part)
        bool act = false;
        if (inb)
        {
            unsigned dmax = 0u;
            for (int i = 0; i < C; i++) dmax |= __float_as_uint(pDy[i]);
            act = (__uint_as_float(dmax) != 0.f);       // zero incoming gradient: only the zero stores below (texture_kernel.cu:922-971)
        }
        int level0 = 0, level1 = 0; float flevel = 0.f;
        int4 tc0 = make_int4(-1, -1, -1, -1), tc1 = make_int4(-1, -1, -1, -1);
        float2 uv0 = make_float2(0.f, 0.f), uv1 = make_float2(0.f, 0.f);
        float4 dw = make_float4(0.f, 0.f, 0.f, 0.f);
        bool two = false;
        float f0 = 1.f;
        if (act)
        {
            int pz = (int)(pidx / hw);
            int tz = (p.texD == 1) ? 0 : pz;
            float2 uv = ld_stream_f2((const float2*)p.uv + pidx);
            if (m.filter == B200DR_FILTER_NEAREST)
            {
                tc0.x = index_nearest(p, m, uv.x, uv.y, tz);
            }
            else
            {
                mip_level<true>(p, m, pidx, level0, level1, flevel, dw);
                uv0 = index_linear(p, m, uv.x, uv.y, tz, level0, tc0);
                two = trilinear && (flevel > 0.f);
                f0 = trilinear ? (1.f - flevel) : 1.f;
                if (two) uv1 = index_linear(p, m, uv.x, uv.y, tz, level1, tc1);
            }
        }

        // ---- texel gradients (warp-uniform control flow)
        if (__any_sync(0xffffffffu, act))
        {
            // few distinct texels in this warp's 32 pixels -> pre-reduce runs; otherwise scatter directly
            const bool reduce = __popc(warp_runs(act ? (tc0.x ^ (level0 << 27)) : -1).heads) <= 10;
            if (m.filter == B200DR_FILTER_NEAREST)
            {
                accum_texel_warp(p.gpad[0], 0, tc0.x, C, Cp, 1.f, pDy, reduce, hot, 0);
            }
            else
            {
                float uv011 = uv0.x * uv0.y, uv010 = uv0.x - uv011, uv001 = uv0.y - uv011, uv000 = 1.f - uv0.x - uv001;
                float* g0 = p.gpad[level0];
                accum_texel_warp(g0, level0, tc0.x, C, Cp, uv000 * f0, pDy, reduce, hot, 0);
                accum_texel_warp(g0, level0, tc0.y, C, Cp, uv010 * f0, pDy, reduce, hot, 1);
                accum_texel_warp(g0, level0, tc0.z, C, Cp, uv001 * f0, pDy, reduce, hot, 2);
                accum_texel_warp(g0, level0, tc0.w, C, Cp, uv011 * f0, pDy, reduce, hot, 3);
                if (__any_sync(0xffffffffu, two))
                {
                    float uv111 = uv1.x * uv1.y, uv110 = uv1.x - uv111, uv101 = uv1.y - uv111, uv100 = 1.f - uv1.x - uv101;
                    float* g1 = p.gpad[level1];
                    accum_texel_warp(g1, level1, two ? tc1.x : -1, C, Cp, uv100 * flevel, pDy, reduce);
                    accum_texel_warp(g1, level1, two ? tc1.y : -1, C, Cp, uv110 * flevel, pDy, reduce);
                    accum_texel_warp(g1, level1, two ? tc1.z : -1, C, Cp, uv101 * flevel, pDy, reduce);
                    accum_texel_warp(g1, level1, two ? tc1.w : -1, C, Cp, uv111 * flevel, pDy, reduce);
                }
            }
        }

        // ---- uv and mip-level gradients
        if (inb)
        {
            float gu = 0.f, gv = 0.f, df = 0.f;
            if (act && m.filter != B200DR_FILTER_NEAREST)
            {
                const float* t0 = p.tex[level0];
                const float* t1 = p.tex[level1];
                float sclu0 = (float)lvl_w(p, level0), sclv0 = (float)lvl_h(p, level0);
                float sclu1 = (float)lvl_w(p, level1), sclv1 = (float)lvl_h(p, level1);
                for (int k = 0; k < Cp; k += 4)
                {
                    float4 A00 = texel4(t0, tc0.x, Cp, k), A10 = texel4(t0, tc0.y, Cp, k), A01 = texel4(t0, tc0.z, Cp, k), A11 = texel4(t0, tc0.w, Cp, k);
                    float4 B00 = make_float4(0.f, 0.f, 0.f, 0.f), B10 = B00, B01 = B00, B11 = B00;
                    if (two) { B00 = texel4(t1, tc1.x, Cp, k); B10 = texel4(t1, tc1.y, Cp, k); B01 = texel4(t1, tc1.z, Cp, k); B11 = texel4(t1, tc1.w, Cp, k); }
                    #pragma unroll
                    for (int j = 0; j < 4; j++)
                    {
                        if (k + j >= C) break;
                        float dy = pDy[k + j];
                        float dy0 = f0 * dy;
                        float a00 = (&A00.x)[j], a10 = (&A10.x)[j], a01 = (&A01.x)[j], a11 = (&A11.x)[j];
                        float ad = (a11 + a00 - a10 - a01);
                        gu += dy0 * ((a10 - a00) + uv0.y * ad) * sclu0;
                        gv += dy0 * ((a01 - a00) + uv0.x * ad) * sclv0;
                        if (two)
                        {
                            float dy1 = flevel * dy;
                            float b00 = (&B00.x)[j], b10 = (&B10.x)[j], b01 = (&B01.x)[j], b11 = (&B11.x)[j];
                            float bd = (b11 + b00 - b10 - b01);
                            gu += dy1 * ((b10 - b00) + uv1.y * bd) * sclu1;
                            gv += dy1 * ((b01 - b00) + uv1.x * bd) * sclv1;
                            float a = bilerp(a00, a10, a01, a11, uv0);
                            float b = bilerp(b00, b10, b01, b11, uv1);
                            df += (b - a) * dy;
                        }
                    }
                }
            }
            if (p.g_uv) st_stream_f2((float2*)p.g_uv + pidx, make_float2(gu, gv));
            if (p.g_bias) st_stream_f(p.g_bias + pidx, df);
            if (p.g_uv_da) st_stream_f4((float4*)p.g_uv_da + pidx, make_float4(dw.x * df, dw.y * df, dw.z * df, dw.w * df));
        }
        if (p.staged) __syncthreads();
    }

    // flush the per-CTA hot-texel cache
    __syncthreads();
    if (hot && tid < TEX_HOT_SLOTS * (TEX_HOT_MAXC / 4))
    {
        int slot = tid / (TEX_HOT_MAXC / 4), k = (tid % (TEX_HOT_MAXC / 4)) * 4;
        int key = s_hot.key[slot];
        if (key != -1 && k < Cp)
        {
            int level = (int)((unsigned)key >> 27), tc = key & ((1 << 27) - 1);
            red_add4(p.gpad[level] + (size_t)tc * Cp + k, make_float4(s_hot.val[slot][k], s_hot.val[slot][k + 1], s_hot.val[slot][k + 2], s_hot.val[slot][k + 3]));
        }
    }
}

// ================================================================================================
// Fast paths: C in {1,2,3,4} (one 128-bit word per texel).  Warp-autonomous like the interpolate kernels: a warp owns
// groups of 32 consecutive pixels, no block barriers in the loop, one vector access per lane for every per-pixel tensor.
// ================================================================================================

template <int C>
static __device__ __forceinline__ void unpack4(float4 v, float (&o)[C])
{
    o[0] = v.x;
    if (C > 1) o[1 % C] = v.y;
    if (C > 2) o[2 % C] = v.z;
    if (C > 3) o[3 % C] = v.w;
}

// Result cache for "uniform" pixel groups.  Rendered images contain large regions whose 32-pixel groups carry bitwise
// identical texture coordinates (the background: uv = uv_da = 0).  Every such group has the same mip level, texel quad,
// weights and fetched texels, so a warp remembers the inputs of the last such group it fully evaluated (key: uv, uv_da,
// bias bits and the texture slice) together with the results that depend only on them, and later groups that match the key
// lane for lane skip level selection, indexing and texel fetches.  The slot lives in shared memory (one per warp); `valid`
// is a warp-uniform register.  Same arithmetic, same results; non-matching groups pay ~12 instructions for the probe.
struct UniKey { unsigned k[8]; };
#ifdef TEX_COUNT_HITS
__device__ unsigned long long g_tex_counters[8];      // [0] fwd hit [1] fwd miss [2] fwd fill [3] grad hit [4] grad miss [5] grad fill [6] grad uniform
#define TEX_COUNT(i) do { if (lane == 0) atomicAdd(&g_tex_counters[i], 1ull); } while (0)
#else
#define TEX_COUNT(i) do { } while (0)
#endif

static __device__ __forceinline__ bool uni_match(const UniKey* slot, float2 uv, float4 da, float bias, int tz)
{
    uint4 a = *(const uint4*)&slot->k[0], b = *(const uint4*)&slot->k[4];
    unsigned diff = (__float_as_uint(uv.x) ^ a.x) | (__float_as_uint(uv.y) ^ a.y) | (__float_as_uint(da.x) ^ a.z) | (__float_as_uint(da.y) ^ a.w)
                  | (__float_as_uint(da.z) ^ b.x) | (__float_as_uint(da.w) ^ b.y) | (__float_as_uint(bias) ^ b.z) | ((unsigned)tz ^ b.w);
    return diff == 0u;
}

// all 32 lanes carry the same inputs as lane 0?  (cheap one-value prefilter first: almost every non-uniform group stops there)
static __device__ __forceinline__ bool uni_identical(float2 uv, float4 da, float bias, int tz)
{
    const unsigned full = 0xffffffffu;
    unsigned x0 = __shfl_sync(full, __float_as_uint(uv.x), 0);
    if (!__all_sync(full, x0 == __float_as_uint(uv.x))) return false;
    unsigned d = (__shfl_sync(full, __float_as_uint(uv.y), 0) ^ __float_as_uint(uv.y));
    d |= (__shfl_sync(full, __float_as_uint(da.x), 0) ^ __float_as_uint(da.x));
    d |= (__shfl_sync(full, __float_as_uint(da.y), 0) ^ __float_as_uint(da.y));
    d |= (__shfl_sync(full, __float_as_uint(da.z), 0) ^ __float_as_uint(da.z));
    d |= (__shfl_sync(full, __float_as_uint(da.w), 0) ^ __float_as_uint(da.w));
    d |= (__shfl_sync(full, __float_as_uint(bias), 0) ^ __float_as_uint(bias));
    d |= ((unsigned)__shfl_sync(full, tz, 0) ^ (unsigned)tz);
    return __all_sync(full, d == 0u);
}

static __device__ __forceinline__ void uni_store_key(UniKey* slot, float2 uv, float4 da, float bias, int tz)
{
    *(uint4*)&slot->k[0] = make_uint4(__float_as_uint(uv.x), __float_as_uint(uv.y), __float_as_uint(da.x), __float_as_uint(da.y));
    *(uint4*)&slot->k[4] = make_uint4(__float_as_uint(da.z), __float_as_uint(da.w), __float_as_uint(bias), (unsigned)tz);
}

// Pixel of lane `lane` in group g.  Default: 32 consecutive pixels of the flat image array.  With TEXG_TILE2D (round-2
// experiment, off by default, needs W % 8 == 0 and H % 4 == 0): an 8x4 pixel tile, so that neighbouring lanes share texel
// quads in both directions (fewer distinct sectors per tap instruction, longer merge runs).
#ifndef TEXG_TILE2D
#define TEXG_TILE2D 1
#endif
template <bool TILE>
static __device__ __forceinline__ size_t group_pixel(const TexParams& p, size_t g, int lane)
{
#if TEXG_TILE2D
    if (TILE && p.tile2d)
    {
        unsigned t = (unsigned)g, n = t / (unsigned)p.tilesPerImg, r = t - n * (unsigned)p.tilesPerImg;
        unsigned ty = r / (unsigned)p.tilesX, tx = r - ty * (unsigned)p.tilesX;
        return ((size_t)n * p.H + ty * 4 + (lane >> 3)) * p.W + tx * 8 + (lane & 7);
    }
#endif
    return g * 32 + lane;
}

#ifndef TEXF_UNROLL
#define TEXF_UNROLL 2
#endif
#ifndef TEXF_MINB
#define TEXF_MINB 4
#endif

// one group's [32][C] output rows: warp-cooperative 128-bit stores (row segment) or per-lane stores (8x4 tile)
template <int C>
static __device__ __forceinline__ void tex_store_out(const TexParams& p, size_t g, size_t pidx, const float (&o)[C], int lane, bool inb, bool full, float* stage)
{
    (void)pidx;
    warp_store_aos<C>(p.out + g * 32 * C, o, lane, inb, full, stage);
}

template <int C, int F, int B>
__global__ void __launch_bounds__(TEX_THREADS, TEXF_MINB) texture_fwd_fast(const TexParams p)
{
    const TexMode m = tex_mode_static<F, B>(p);
    constexpr int SK = (C == 3) ? 3 : 1;
    __shared__ __align__(16) float s_stage[TEX_THREADS / 32][32 * SK];
    __shared__ __align__(16) UniKey s_key[TEX_THREADS / 32];
    __shared__ __align__(16) float4 s_val[TEX_THREADS / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t hw = (size_t)p.H * p.W;
    const size_t total = hw * p.N;
    const size_t groups = (total + 31) >> 5;
    const size_t wstride = (size_t)gridDim.x * (TEX_THREADS / 32);
    const bool mip = (m.filter == B200DR_FILTER_LINEAR_MIPMAP_NEAREST || m.filter == B200DR_FILTER_LINEAR_MIPMAP_LINEAR);
    bool cached = false;                                // warp-uniform: s_key/s_val[warp] hold a uniform group's inputs/result

    for (size_t g0 = (size_t)blockIdx.x * (TEX_THREADS / 32) + warp; g0 < groups; g0 += TEXF_UNROLL * wstride)
    {
        // issue the streaming loads of all groups in flight first
        float2 uv[TEXF_UNROLL]; float4 da[TEXF_UNROLL]; float bias[TEXF_UNROLL]; bool inb[TEXF_UNROLL];
        #pragma unroll
        for (int u = 0; u < TEXF_UNROLL; u++)
        {
            size_t pidx = group_pixel<false>(p, g0 + u * wstride, lane);
            inb[u] = (g0 + u * wstride < groups) && (pidx < total);
            uv[u] = make_float2(0.f, 0.f); da[u] = make_float4(0.f, 0.f, 0.f, 0.f); bias[u] = 0.f;
            if (inb[u])
            {
                uv[u] = ld_stream_f2((const float2*)p.uv + pidx);
                if (mip && m.has_da) da[u] = ld_stream_f4((const float4*)p.uv_da + pidx);
                if (mip && m.has_bias) bias[u] = ld_stream_f(p.bias + pidx);
            }
        }
        #pragma unroll
        for (int u = 0; u < TEXF_UNROLL; u++)
        {
            size_t g = g0 + u * wstride;
            if (g >= groups) continue;                  // warp-uniform
            size_t pidx = group_pixel<false>(p, g, lane);
            const bool full = (g * 32 + 32 <= total);
            const int tz = (p.texD == 1) ? 0 : (int)(pidx / hw);
            float o[C];
            if (cached && full && __all_sync(0xffffffffu, uni_match(&s_key[warp], uv[u], da[u], bias[u], tz)))
            {
                TEX_COUNT(0);
                unpack4<C>(s_val[warp], o);
                tex_store_out<C>(p, g, pidx, o, lane, true, true, s_stage[warp]);
                continue;
            }
            TEX_COUNT(1);
            float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
            if (inb[u])
            {
                if (m.filter == B200DR_FILTER_NEAREST) r = texel4(p.tex[0], index_nearest(p, m, uv[u].x, uv[u].y, tz), 4, 0);
                else
                {
                    int level0, level1; float flevel; float4 dw;
                    mip_level_from<false>(p, m, da[u], bias[u], level0, level1, flevel, dw);
                    int4 tc0;
                    float2 uv0 = index_linear(p, m, uv[u].x, uv[u].y, tz, level0, tc0);
                    const float* t0 = p.tex[level0];
                    r = bilerp4(texel4(t0, tc0.x, 4, 0), texel4(t0, tc0.y, 4, 0), texel4(t0, tc0.z, 4, 0), texel4(t0, tc0.w, 4, 0), uv0);
                    if (m.filter == B200DR_FILTER_LINEAR_MIPMAP_LINEAR && flevel > 0.f)
                    {
                        int4 tc1;
                        float2 uv1 = index_linear(p, m, uv[u].x, uv[u].y, tz, level1, tc1);
                        const float* t1 = p.tex[level1];
                        float4 b = bilerp4(texel4(t1, tc1.x, 4, 0), texel4(t1, tc1.y, 4, 0), texel4(t1, tc1.z, 4, 0), texel4(t1, tc1.w, 4, 0), uv1);
                        r = make_float4(lerpf(r.x, b.x, flevel), lerpf(r.y, b.y, flevel), lerpf(r.z, b.z, flevel), lerpf(r.w, b.w, flevel));
                    }
                }
            }
            unpack4<C>(r, o);
            tex_store_out<C>(p, g, pidx, o, lane, inb[u], full, s_stage[warp]);
            if (full && uni_identical(uv[u], da[u], bias[u], tz))
            {
                TEX_COUNT(2);
                if (lane == 0) { uni_store_key(&s_key[warp], uv[u], da[u], bias[u], tz); s_val[warp] = r; }
                __syncwarp();
                cached = true;
            }
        }
    }
}

// sum v over the whole warp
static __device__ __forceinline__ float warp_sum(float v)
{
    #pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    return v;
}

// One bilinear tap of the texel-gradient scatter: a 128-bit vector reduction per lane.
template <int C>
static __device__ __forceinline__ void scatter_tap(float* g, int tc, float w, const float (&dy)[C])
{
    if (tc < 0) return;
    red_add4(g + (size_t)tc * 4, make_float4(w * dy[0], (C > 1) ? w * dy[1 % C] : 0.f, (C > 2) ? w * dy[2 % C] : 0.f, (C > 3) ? w * dy[3 % C] : 0.f));
}

// Pending contribution of "uniform" pixel groups (all 32 pixels on the same texel quad with the same weights: constant uv,
// typically the background).  A warp keeps the summed incoming gradient of consecutive uniform groups in registers and
// issues the four tap reductions only when the quad changes or the kernel ends: instead of one same-address atomic per
// pixel (serialised chip-wide at ~0.85 clk each) the texel sees a handful per warp.
template <int C>
struct Pending
{
    int level; int4 tc; float fu, fv;   // identity: level, the four tap indices and the fractional coordinates (level < 0: empty)
    float4 w;                           // tap weights (a function of fu, fv)
    float sum[C];                       // PER-LANE partial sums of the incoming gradient; reduced across the warp at flush time
};

// all lanes call (pd.level is warp-uniform)
template <int C>
static __device__ __forceinline__ void pending_flush(Pending<C>& pd, const TexParams& p, int lane)
{
    if (pd.level >= 0)
    {
        float tot[C];
        #pragma unroll
        for (int i = 0; i < C; i++) tot[i] = warp_sum(pd.sum[i]);
        if (lane < 4)
        {
            int tc = (lane == 0) ? pd.tc.x : (lane == 1) ? pd.tc.y : (lane == 2) ? pd.tc.z : pd.tc.w;
            float w = (lane == 0) ? pd.w.x : (lane == 1) ? pd.w.y : (lane == 2) ? pd.w.z : pd.w.w;
            scatter_tap<C>(p.gpad[pd.level], tc, w, tot);
        }
    }
    pd.level = -1;                                      // empty
}

#ifndef TEXG_THREADS
#define TEXG_THREADS 128
#endif
#ifndef TEXG_UNROLL
#define TEXG_UNROLL  1
#endif
#ifndef TEXG_MINB
#define TEXG_MINB 5
#endif

// Streaming inputs of one 32-pixel group, loaded up front (two groups are in flight per warp).
template <int C>
struct TexGradIn { AosRaw<C> dy; float2 uv; float4 da; float bias; bool inb, full; size_t g, pidx; };

template <int C>
static __device__ __forceinline__ void texgrad_issue(const TexParams& p, const TexMode& m, size_t g, size_t groups, size_t total, int lane, TexGradIn<C>& in)
{
    in.g = g;
    size_t pidx = group_pixel<true>(p, g, lane);
    in.pidx = pidx;
    in.inb = (g < groups) && (pidx < total);
    in.full = (g < groups) && (g * 32 + 32 <= total);
    in.uv = make_float2(0.f, 0.f); in.da = make_float4(0.f, 0.f, 0.f, 0.f); in.bias = 0.f;
    if (in.inb)
    {
        in.uv = ld_stream_f2((const float2*)p.uv + pidx);
        if (m.has_da) in.da = ld_stream_f4((const float4*)p.uv_da + pidx);
        if (m.has_bias) in.bias = ld_stream_f(p.bias + pidx);
    }
#if TEXG_TILE2D
    if (p.tile2d)
    {
        in.dy.q[0] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (in.inb)
        {
            const float* d = p.dy + pidx * C;
            in.dy.q[0].x = ld_stream_f(d);
            if (C > 1) in.dy.q[0].y = ld_stream_f(d + 1 % C);
            if (C > 2) in.dy.q[0].z = ld_stream_f(d + 2 % C);
            if (C > 3) in.dy.q[0].w = ld_stream_f(d + 3 % C);
        }
        return;
    }
#endif
    warp_load_aos_issue<C>(p.dy + g * 32 * C, in.dy, lane, in.inb, in.full);
}

// Per-warp slot of the uniform-group cache in the backward pass: key + d(out)/d(u,v) per channel (gu = sum_i dy_i * ku_i).
struct UniGradSlot { UniKey key; float4 ku, kv; };

template <int C>
static __device__ __forceinline__ float dot4(float4 k, const float (&dy)[C])
{
    float r = dy[0] * k.x;
    if (C > 1) r += dy[1 % C] * k.y;
    if (C > 2) r += dy[2 % C] * k.z;
    if (C > 3) r += dy[3 % C] * k.w;
    return r;
}

// Per-pixel results of one group: gradient w.r.t. uv, the level (bias) and uv_da.  The stand-alone texture backward stores them
// (STORE); the fused chain backward keeps them in registers and feeds them straight into the interpolate / rasterize math.
struct TexPixelGrad { float gu, gv, df; float4 gda; };

template <int C, int B, bool STORE>
static __device__ __forceinline__ TexPixelGrad texgrad_group(const TexParams& p, const TexMode& m, const TexGradIn<C>& in, Pending<C>& pd, bool& cached,
                                                             UniGradSlot* slot, int lane, float* sw, size_t hw, bool trilinear)
{
    TexPixelGrad res; res.gu = 0.f; res.gv = 0.f; res.df = 0.f; res.gda = make_float4(0.f, 0.f, 0.f, 0.f);
    const size_t pidx = in.pidx;
    const bool inb = in.inb;
    float dy[C];
#if TEXG_TILE2D
    if (p.tile2d) unpack4<C>(in.dy.q[0], dy);
    else
#endif
    warp_load_aos_finish<C>(p.dy + in.g * 32 * C, in.dy, dy, lane, inb, in.full, sw);
    const float2 uv = in.uv;
    const int tz = (p.texD == 1) ? 0 : (int)(pidx / hw);

    // Uniform group seen before (see UniKey): level, quad, weights and texels are those of the cached group, which was a
    // single-level (flevel == 0) group, so g_bias = g_uv_da = 0; the texel gradient joins the pending sums of that quad.
    if (cached && in.full && __all_sync(0xffffffffu, uni_match(&slot->key, uv, in.da, in.bias, tz)))
    {
        TEX_COUNT(3);
        #pragma unroll
        for (int i = 0; i < C; i++) pd.sum[i] += dy[i];
        res.gu = dot4<C>(slot->ku, dy); res.gv = dot4<C>(slot->kv, dy);
        if (STORE)
        {
            if (p.g_uv) st_stream_f2((float2*)p.g_uv + pidx, make_float2(res.gu, res.gv));
            if (p.g_bias) st_stream_f(p.g_bias + pidx, 0.f);
            if (p.g_uv_da) st_stream_f4((float4*)p.g_uv_da + pidx, make_float4(0.f, 0.f, 0.f, 0.f));
        }
        return res;
    }

    TEX_COUNT(4);
    unsigned dmax = 0u;
    #pragma unroll
    for (int i = 0; i < C; i++) dmax |= __float_as_uint(dy[i]);
    bool act = inb && (__uint_as_float(dmax) != 0.f);

    float gu = 0.f, gv = 0.f, df = 0.f;
    float4 dw = make_float4(0.f, 0.f, 0.f, 0.f);
    if (__any_sync(0xffffffffu, act))
    {
        int level0 = 0, level1 = 0; float flevel = 0.f;
        int4 tc0 = make_int4(-1, -1, -1, -1), tc1 = tc0;
        float2 uv0 = make_float2(0.f, 0.f), uv1 = uv0;
        bool two = false;
        float f0 = 1.f;
        if (act)
        {
            if (m.filter == B200DR_FILTER_NEAREST) tc0.x = index_nearest(p, m, uv.x, uv.y, tz);
            else
            {
                mip_level_from<true>(p, m, in.da, in.bias, level0, level1, flevel, dw);
                uv0 = index_linear(p, m, uv.x, uv.y, tz, level0, tc0);
                two = trilinear && (flevel > 0.f);
                f0 = trilinear ? (1.f - flevel) : 1.f;
                if (two) uv1 = index_linear(p, m, uv.x, uv.y, tz, level1, tc1);
            }
        }
        float uv011 = uv0.x * uv0.y, uv010 = uv0.x - uv011, uv001 = uv0.y - uv011, uv000 = 1.f - uv0.x - uv001;
        float4 w0 = (m.filter == B200DR_FILTER_NEAREST) ? make_float4(1.f, 0.f, 0.f, 0.f)
                                                        : make_float4(uv000 * f0, uv010 * f0, uv001 * f0, uv011 * f0);

        // Whole warp on one texel quad with identical weights?  Identity = level, ALL four tap indices (in zero-boundary mode the
        // first tap can be -1 for every lane while the others differ) and both fractional coordinates (mirrored pairs (u,v),
        // (1-v,1-u) share two of the four weights).  One-value prefilter first: almost every non-uniform group stops there.
        const unsigned full = 0xffffffffu;
        int key0 = act ? (tc0.x ^ (level0 << 27)) : -1 - lane;
        const int key_l0 = __shfl_sync(full, key0, 0);                  // collectives are evaluated by every lane: never inside a short-circuit
        bool uniform = __all_sync(full, act && !two && key0 == key_l0);
        if (uniform)
        {
            unsigned d = (unsigned)(__shfl_sync(full, tc0.y, 0) ^ tc0.y) | (unsigned)(__shfl_sync(full, tc0.z, 0) ^ tc0.z) | (unsigned)(__shfl_sync(full, tc0.w, 0) ^ tc0.w);
            d |= (unsigned)(__shfl_sync(full, __float_as_int(uv0.x), 0) ^ __float_as_int(uv0.x)) | (unsigned)(__shfl_sync(full, __float_as_int(uv0.y), 0) ^ __float_as_int(uv0.y));
            d |= (unsigned)(__shfl_sync(full, level0, 0) ^ level0);
            uniform = __all_sync(full, d == 0u);
        }
        bool fill = false;
        if (uniform)
        {
            TEX_COUNT(6);
            const bool same_quad = pd.level == level0 && pd.tc.x == tc0.x && pd.tc.y == tc0.y && pd.tc.z == tc0.z && pd.tc.w == tc0.w &&
                                   __float_as_int(pd.fu) == __float_as_int(uv0.x) && __float_as_int(pd.fv) == __float_as_int(uv0.y);
            if (!same_quad)                             // warp-uniform: every lane holds the same quad
            {
                pending_flush<C>(pd, p, lane);
                pd.level = level0; pd.tc = tc0; pd.fu = uv0.x; pd.fv = uv0.y; pd.w = w0;
                #pragma unroll
                for (int i = 0; i < C; i++) pd.sum[i] = 0.f;
                cached = false;                         // the cache is only valid while its quad is the pending one
            }
            #pragma unroll
            for (int i = 0; i < C; i++) pd.sum[i] += dy[i];
            fill = !cached && in.full && (m.filter != B200DR_FILTER_NEAREST) && uni_identical(uv, in.da, in.bias, tz);
        }
        else
        {
            // Neighbouring pixels of a magnified texture sit on the same texel quad: sum runs of up to 4 equal quads
            // with two shuffle steps, so a quad receives one vector reduction per run instead of one per pixel
            // (same-address reductions inside one instruction serialise).  Skipped when the warp has few duplicates.
            const bool pinned = (B == B200DR_BOUNDARY_WRAP || B == B200DR_BOUNDARY_CLAMP) ? true : (tc0.x >= 0 && tc0.w >= 0);
            WarpRuns runs = warp_runs_max4((act && !two && pinned) ? (tc0.x ^ (level0 << 27)) : -1 - lane, tc0.w);    // first and diagonal tap pin the whole quad
            const bool merge = (__popc(runs.heads) <= 24);
            float* g0 = p.gpad[level0];
            const int ntap = (m.filter == B200DR_FILTER_NEAREST) ? 1 : 4;
            for (int t = 0; t < ntap; t++)
            {
                int tc = (t == 0) ? tc0.x : (t == 1) ? tc0.y : (t == 2) ? tc0.z : tc0.w;
                float w = (t == 0) ? w0.x : (t == 1) ? w0.y : (t == 2) ? w0.z : w0.w;
                float v[C];
                #pragma unroll
                for (int i = 0; i < C; i++) v[i] = act ? w * dy[i] : 0.f;
                if (merge)
                {
                    #pragma unroll
                    for (int i = 0; i < C; i++) v[i] = warp_run_sum_max4(v[i], runs);
                    if (runs.is_head) scatter_tap<C>(g0, tc, 1.f, v);
                }
                else scatter_tap<C>(g0, tc, 1.f, v);
            }
            if (two)
            {
                float uv111 = uv1.x * uv1.y, uv110 = uv1.x - uv111, uv101 = uv1.y - uv111, uv100 = 1.f - uv1.x - uv101;
                float* g1 = p.gpad[level1];
                scatter_tap<C>(g1, tc1.x, uv100 * flevel, dy);
                scatter_tap<C>(g1, tc1.y, uv110 * flevel, dy);
                scatter_tap<C>(g1, tc1.z, uv101 * flevel, dy);
                scatter_tap<C>(g1, tc1.w, uv111 * flevel, dy);
            }
        }

        if (act && m.filter != B200DR_FILTER_NEAREST)
        {
            const float* t0 = p.tex[level0];
            float sclu0 = (float)lvl_w(p, level0), sclv0 = (float)lvl_h(p, level0);
            float a00[C], a10[C], a01[C], a11[C];
            unpack4<C>(texel4(t0, tc0.x, 4, 0), a00); unpack4<C>(texel4(t0, tc0.y, 4, 0), a10);
            unpack4<C>(texel4(t0, tc0.z, 4, 0), a01); unpack4<C>(texel4(t0, tc0.w, 4, 0), a11);
            float ku[4] = { 0.f, 0.f, 0.f, 0.f }, kv[4] = { 0.f, 0.f, 0.f, 0.f };
            #pragma unroll
            for (int i = 0; i < C; i++)
            {
                float ad = (a11[i] + a00[i] - a10[i] - a01[i]);
                ku[i] = f0 * ((a10[i] - a00[i]) + uv0.y * ad) * sclu0;
                kv[i] = f0 * ((a01[i] - a00[i]) + uv0.x * ad) * sclv0;
                gu += dy[i] * ku[i];
                gv += dy[i] * kv[i];
            }
            if (fill)                                   // warp-uniform; every lane holds the same ku / kv
            {
                TEX_COUNT(5);
                if (lane == 0)
                {
                    uni_store_key(&slot->key, uv, in.da, in.bias, tz);
                    slot->ku = make_float4(ku[0], ku[1], ku[2], ku[3]); slot->kv = make_float4(kv[0], kv[1], kv[2], kv[3]);
                }
                __syncwarp();
                cached = true;
            }
            if (two)
            {
                const float* t1 = p.tex[level1];
                float sclu1 = (float)lvl_w(p, level1), sclv1 = (float)lvl_h(p, level1);
                float b00[C], b10[C], b01[C], b11[C];
                unpack4<C>(texel4(t1, tc1.x, 4, 0), b00); unpack4<C>(texel4(t1, tc1.y, 4, 0), b10);
                unpack4<C>(texel4(t1, tc1.z, 4, 0), b01); unpack4<C>(texel4(t1, tc1.w, 4, 0), b11);
                #pragma unroll
                for (int i = 0; i < C; i++)
                {
                    float dy1 = flevel * dy[i];
                    float bd = (b11[i] + b00[i] - b10[i] - b01[i]);
                    gu += dy1 * ((b10[i] - b00[i]) + uv1.y * bd) * sclu1;
                    gv += dy1 * ((b01[i] - b00[i]) + uv1.x * bd) * sclv1;
                    float a = bilerp(a00[i], a10[i], a01[i], a11[i], uv0);
                    float b = bilerp(b00[i], b10[i], b01[i], b11[i], uv1);
                    df += (b - a) * dy[i];
                }
            }
        }
    }
    res.gu = gu; res.gv = gv; res.df = df; res.gda = make_float4(dw.x * df, dw.y * df, dw.z * df, dw.w * df);
    if (STORE && inb)
    {
        if (p.g_uv) st_stream_f2((float2*)p.g_uv + pidx, make_float2(gu, gv));
        if (p.g_bias) st_stream_f(p.g_bias + pidx, df);
        if (p.g_uv_da) st_stream_f4((float4*)p.g_uv_da + pidx, res.gda);
    }
    return res;
}

template <int C, int F, int B>
__global__ void __launch_bounds__(TEXG_THREADS, TEXG_MINB) texture_grad_fast(const TexParams p)
{
    const TexMode m = tex_mode_static<F, B>(p);
    constexpr int SK = (C == 3) ? 3 : 1;
    __shared__ __align__(16) float s_stage[TEXG_THREADS / 32][32 * SK];
    __shared__ __align__(16) UniGradSlot s_slot[TEXG_THREADS / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const size_t hw = (size_t)p.H * p.W;
    const size_t total = hw * p.N;
    const size_t groups = (total + 31) >> 5;
    const size_t wstride = (size_t)gridDim.x * (TEXG_THREADS / 32);
    const bool trilinear = (m.filter == B200DR_FILTER_LINEAR_MIPMAP_LINEAR);
    Pending<C> pd;
    pd.level = -1; pd.tc = make_int4(-1, -1, -1, -1); pd.fu = 0.f; pd.fv = 0.f; pd.w = make_float4(0.f, 0.f, 0.f, 0.f);
    #pragma unroll
    for (int i = 0; i < C; i++) pd.sum[i] = 0.f;
    bool cached = false;

    for (size_t g0 = (size_t)blockIdx.x * (TEXG_THREADS / 32) + warp; g0 < groups; g0 += TEXG_UNROLL * wstride)
    {
        TexGradIn<C> in[TEXG_UNROLL];
        #pragma unroll
        for (int u = 0; u < TEXG_UNROLL; u++) texgrad_issue<C>(p, m, g0 + u * wstride, groups, total, lane, in[u]);
        #pragma unroll
        for (int u = 0; u < TEXG_UNROLL; u++)
            if (in[u].g < groups) texgrad_group<C, B, true>(p, m, in[u], pd, cached, &s_slot[warp], lane, s_stage[warp], hw, trilinear);   // warp-uniform
    }
    pending_flush<C>(pd, p, lane);
}

// ================================================================================================
// Fused chain backward: texture backward -> interpolate backward -> rasterize backward in one pass.
//
// The op boundaries of the reference API force a tensor round trip between consecutive backward ops: texture_grad writes
// g_uv / g_uv_da (24 B/px) which interpolate_grad reads back together with rast, rast_db, and writes g_rast / g_rast_db
// (32 B/px) which rasterize_grad reads back with rast again.  When the front end can prove that texture()'s uv / uv_da are
// the untouched outputs of interpolate(attr, rast, rast_db, diff_attrs='all') on the untouched outputs of rasterize(pos),
// it routes the texture gradient through this kernel instead: per pixel it reads rast, rast_db and dy only (44 B instead of
// 196), RECOMPUTES uv / uv_da from the vertex attributes (same expressions as interpolate_fwd_fast), and keeps the chain
//      dy -> (g_uv, g_uv_da) -> (g_attr; g_rast, g_rast_db) -> g_pos
// in registers.  Background pixels (no triangle) cost rast + dy and a register add (their texel gradient joins the pending
// sums of the quad at uv = 0).  RASTER = false stops after interpolate backward and writes g_rast / g_rast_db (range-mode
// meshes, or rast consumers the front end cannot see through).
// ================================================================================================

struct ChainParams
{
    const float* rast; const float* rast_db;            // [N,H,W,4]
    const float* attr; const int32_t* attr_tri;         // [Va,2] (shared by all images) or [N,Va,2]; [T,3]
    const float* pos; const int32_t* pos_tri;           // RASTER: [N,V,4]; [T,3]
    float* g_attr; float* g_pos;                        // nullable: gradient not required
    float* g_rast; float* g_rast_db;                    // !RASTER
    int Va, V, T, attr_per_image, grad_db;
    float xs, xo, ys, yo;
    unsigned* strip_counter;                            // STRIPS schedule: zeroed work counter (tail of the scratch buffer)
};

// uv = b0 a0 + b1 a1 + b2 a2 and its screen-space derivatives: the expressions of interpolate_fwd_fast (A = 2, all
// differentials), kept textually identical so that the recomputed values are the ones the forward pass stored.
static __device__ __forceinline__ void interp_uv(float4 r, float4 db, const float (&a0)[2], const float (&a1)[2], const float (&a2)[2], float2& uv, float4& da)
{
    float b0 = r.x, b1 = r.y, b2 = 1.f - r.x - r.y;
    float o[2], oda[4];
    #pragma unroll
    for (int i = 0; i < 2; i++)
    {
        o[i] = b0 * a0[i] + b1 * a1[i] + b2 * a2[i];
        float dsdu = a0[i] - a2[i], dsdv = a1[i] - a2[i];
        oda[2 * i] = db.x * dsdu + db.z * dsdv;
        oda[2 * i + 1] = db.y * dsdu + db.w * dsdv;
    }
    uv = make_float2(o[0], o[1]);
    da = make_float4(oda[0], oda[1], oda[2], oda[3]);
}

#ifndef TCG_MINB
#define TCG_MINB 5          // <= 102 registers: 20 warps per SM (measured 2.72 -> 2.39 ms against 4 CTAs at 118 registers; 3 CTAs: 2.82)
#endif
// Work schedule of the fused kernel.  0: grid-stride over the flat array of 32-pixel groups (a warp's consecutive groups lie
// gridDim * 128 pixels apart).  1: STRIPS -- a warp takes 32-pixel-wide, TCG_STRIP_ROWS-high strips from a global work counter
// and walks them top to bottom: vertically adjacent groups share their triangles and texel quads (index / vertex / texel
// gathers hit L1 instead of L2), the pixel coordinates need no division, and the counter balances background against
// geometry exactly instead of statistically.  Needs W % 32 == 0 and H % TCG_STRIP_ROWS == 0 (otherwise schedule 0).
#ifndef TCG_STRIPS
#define TCG_STRIPS 1
#endif
#ifndef TCG_STRIP_ROWS
#define TCG_STRIP_ROWS 8
#endif
#ifndef TCG_PF_DB
#define TCG_PF_DB 1         // L2 prefetch of the next group's rast_db line as well
#endif
#ifndef TCG_EARLY_DB
#define TCG_EARLY_DB 1      // request rast_db with the index gathers instead of after their validity test
#endif
#ifndef TCG_STRIP_WAVES
#define TCG_STRIP_WAVES 1   // CTAs launched per resident CTA slot (1 = persistent grid)
#endif

template <int C, int F, int B, bool RASTER, bool STRIPS>
__global__ void __launch_bounds__(TEXG_THREADS, TCG_MINB) texture_chain_grad_fast(const TexParams p, const ChainParams q)
{
    const TexMode m = tex_mode_static<F, B>(p);
    constexpr int SK = (C == 3) ? 3 : 1;
    __shared__ __align__(16) float s_stage[TEXG_THREADS / 32][32 * SK];
    __shared__ __align__(16) UniGradSlot s_slot[TEXG_THREADS / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const size_t hw = (size_t)p.H * p.W;
    const size_t total = hw * p.N;
    const size_t groups = (total + 31) >> 5;
    const size_t wstride = (size_t)gridDim.x * (TEXG_THREADS / 32);
    const bool trilinear = (m.filter == B200DR_FILTER_LINEAR_MIPMAP_LINEAR);
    const size_t NO_GROUP = ~(size_t)0;
    Pending<C> pd;
    pd.level = -1; pd.tc = make_int4(-1, -1, -1, -1); pd.fu = 0.f; pd.fv = 0.f; pd.w = make_float4(0.f, 0.f, 0.f, 0.f);
    #pragma unroll
    for (int i = 0; i < C; i++) pd.sum[i] = 0.f;
    bool cached = false;

    // One 32-pixel group: g = group index in the flat pixel array, n = image of lane 0 .. (n + 1 when the group straddles two
    // images, schedule 0 only), rl = this lane's pixel index inside its image, (px, py) = its coordinates (STRIPS only), gn = the
    // group this warp will process next (L2 prefetch) or NO_GROUP.
    auto process = [&](const size_t g, const int n, const unsigned rl, const unsigned px_, const unsigned py_, const size_t gn)
    {
        TexGradIn<C> in;
        in.g = g; in.pidx = g * 32 + lane;
        in.inb = STRIPS ? true : (in.pidx < total); in.full = STRIPS ? true : (g * 32 + 32 <= total);
        in.bias = 0.f;
        const size_t pidx = in.pidx;
        float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
        if (in.inb) r = ld_stream_f4((const float4*)q.rast + pidx);
        warp_load_aos_issue<C>(p.dy + g * 32 * C, in.dy, lane, in.inb, in.full);
        {   // this warp's next group: one prefetch.global.L2 per lane and tensor (LSU path; measured -5 % on this kernel, whose
            // iterations are long enough for the lines to arrive; the bulk / TMA form and the short-iteration kernels lose)
            if (STRIPS ? (gn != NO_GROUP) : (gn * 32 + 32 <= total))
            {
                asm volatile("prefetch.global.L2 [%0];" :: "l"((const float4*)q.rast + gn * 32 + lane) : "memory");
                if (lane * 16 < 32 * C * 4) asm volatile("prefetch.global.L2 [%0];" :: "l"((const char*)(p.dy + gn * 32 * C) + lane * 16) : "memory");
#if TCG_PF_DB
                // rast_db is read under geometry only, AFTER the index gathers have been checked: without this its DRAM latency is
                // paid in series with theirs (background groups waste the 512 bytes; DRAM runs at 15 % in this kernel)
                if (q.rast_db) asm volatile("prefetch.global.L2 [%0];" :: "l"((const float4*)q.rast_db + gn * 32 + lane) : "memory");
#endif
            }
        }

        // ---- triangle of the pixel, attribute rows, recomputed texture coordinates
        int triIdx = in.inb ? float_to_triidx(r.w) - 1 : -1;
        if (triIdx >= q.T) triIdx = -1;
        int ai0 = 0, ai1 = 0, ai2 = 0, vi0 = 0, vi1 = 0, vi2 = 0;
#if TCG_EARLY_DB
        // requested together with the index gathers (it only needs the coverage bit), not after their validity test
        float4 db_early = make_float4(0.f, 0.f, 0.f, 0.f);
        if (triIdx >= 0 && q.rast_db) db_early = ld_stream_f4((const float4*)q.rast_db + pidx);
#endif
        if (triIdx >= 0)
        {
            ai0 = __ldg(q.attr_tri + triIdx * 3); ai1 = __ldg(q.attr_tri + triIdx * 3 + 1); ai2 = __ldg(q.attr_tri + triIdx * 3 + 2);
            if (RASTER && q.g_pos)                      // position indices are fetched with the attribute indices: one round trip
            {
                vi0 = __ldg(q.pos_tri + triIdx * 3); vi1 = __ldg(q.pos_tri + triIdx * 3 + 1); vi2 = __ldg(q.pos_tri + triIdx * 3 + 2);
            }
            if (ai0 < 0 || ai0 >= q.Va || ai1 < 0 || ai1 >= q.Va || ai2 < 0 || ai2 >= q.Va) triIdx = -1;
        }
        const bool valid = triIdx >= 0;
        const int aoff = q.attr_per_image ? n * q.Va : 0;
        float4 db = make_float4(0.f, 0.f, 0.f, 0.f);
        float a0[2] = { 0.f, 0.f }, a1[2] = { 0.f, 0.f }, a2[2] = { 0.f, 0.f };
        in.uv = make_float2(0.f, 0.f); in.da = make_float4(0.f, 0.f, 0.f, 0.f);
        if (valid)
        {
#if !TCG_EARLY_DB
            if (q.rast_db) db = ld_stream_f4((const float4*)q.rast_db + pidx);
#else
            db = db_early;
#endif
            float2 t;
            t = __ldg((const float2*)q.attr + aoff + ai0); a0[0] = t.x; a0[1] = t.y;
            t = __ldg((const float2*)q.attr + aoff + ai1); a1[0] = t.x; a1[1] = t.y;
            t = __ldg((const float2*)q.attr + aoff + ai2); a2[0] = t.x; a2[1] = t.y;
            interp_uv(r, db, a0, a1, a2, in.uv, in.da);
            if (!m.has_da) in.da = make_float4(0.f, 0.f, 0.f, 0.f);
        }

        // ---- texture backward of the group: texel gradients are scattered, (g_uv, g_uv_da) stay in registers
        const TexPixelGrad tg = texgrad_group<C, B, false>(p, m, in, pd, cached, &s_slot[warp], lane, s_stage[warp], hw, trilinear);

        if (!__any_sync(0xffffffffu, valid))            // background group: nothing flows further
        {
            if (!RASTER && q.g_rast && in.inb)
            {
                st_stream_f4((float4*)q.g_rast + pidx, make_float4(0.f, 0.f, 0.f, 0.f));
                if (q.g_rast_db) st_stream_f4((float4*)q.g_rast_db + pidx, make_float4(0.f, 0.f, 0.f, 0.f));
            }
            return;
        }

        // ---- interpolate backward (A = 2, all differentials): csrc/common/interpolate.cu:131-274 restated for registers
        const float y[2] = { valid ? tg.gu : 0.f, valid ? tg.gv : 0.f };
        const float yd[4] = { valid ? tg.gda.x : 0.f, valid ? tg.gda.y : 0.f, valid ? tg.gda.z : 0.f, valid ? tg.gda.w : 0.f };
        float2 gr = make_float2(0.f, 0.f);
        float4 gdb = make_float4(0.f, 0.f, 0.f, 0.f);
        float t0[2], t1[2], t2[2];
        {
            const float b0 = r.x, b1 = r.y, b2 = 1.f - r.x - r.y;
            #pragma unroll
            for (int i = 0; i < 2; i++)
            {
                const float dsdu = a0[i] - a2[i], dsdv = a1[i] - a2[i];
                gr.x += y[i] * dsdu; gr.y += y[i] * dsdv;
                const float dsdx = yd[2 * i], dsdy = yd[2 * i + 1];
                gdb.x += dsdu * dsdx; gdb.y += dsdu * dsdy; gdb.z += dsdv * dsdx; gdb.w += dsdv * dsdy;
                const float du = dsdx * db.x + dsdy * db.y, dv = dsdx * db.z + dsdy * db.w;
                t0[i] = b0 * y[i] + du; t1[i] = b1 * y[i] + dv; t2[i] = b2 * y[i] - du - dv;
            }
        }
        const WarpRuns runs = warp_runs2(valid ? triIdx : -1 - lane, n);     // (triangle, image): a group may straddle two images
        const bool head = runs.is_head && valid;
        if (q.g_attr)                                   // warp-uniform
        {
            #pragma unroll
            for (int i = 0; i < 2; i++) { t0[i] = warp_run_sum(t0[i], runs); t1[i] = warp_run_sum(t1[i], runs); t2[i] = warp_run_sum(t2[i], runs); }
            if (head)
            {
                float* ga = q.g_attr + (size_t)aoff * 2;
                red_add2(ga + 2 * ai0, make_float2(t0[0], t0[1])); red_add2(ga + 2 * ai1, make_float2(t1[0], t1[1])); red_add2(ga + 2 * ai2, make_float2(t2[0], t2[1]));
            }
        }
        if (!RASTER)
        {
            if (q.g_rast && in.inb)
            {
                st_stream_f4((float4*)q.g_rast + pidx, valid ? make_float4(gr.x, gr.y, 0.f, 0.f) : make_float4(0.f, 0.f, 0.f, 0.f));
                if (q.g_rast_db) st_stream_f4((float4*)q.g_rast_db + pidx, valid ? gdb : make_float4(0.f, 0.f, 0.f, 0.f));
            }
            return;
        }

        // ---- rasterize backward: pm::raster_grad on (g_rast.xy, g_rast_db), one 128-bit reduction per vertex and triangle run
        if (q.g_pos)                                    // warp-uniform
        {
            float gv9[9];
            #pragma unroll
            for (int i = 0; i < 9; i++) gv9[i] = 0.f;
            bool pv = valid;
            if (pv && (vi0 < 0 || vi0 >= q.V || vi1 < 0 || vi1 >= q.V || vi2 < 0 || vi2 >= q.V)) pv = false;
            // the reference skips pixels whose incoming gradients are all +-0 (csrc/common/rasterize.cu:143-148)
            const unsigned nz_dy = (__float_as_uint(gr.x) | __float_as_uint(gr.y)) << 1;
            const unsigned nz_db = q.grad_db ? ((__float_as_uint(gdb.x) | __float_as_uint(gdb.y) | __float_as_uint(gdb.z) | __float_as_uint(gdb.w)) << 1) : 0u;
            if ((nz_dy | nz_db) == 0u) pv = false;
            if (pv)
            {
                const float4* vb = (const float4*)q.pos + (size_t)n * q.V;
                const unsigned py = STRIPS ? py_ : rl / (unsigned)p.W, px = STRIPS ? px_ : rl - py * (unsigned)p.W;
                const float fx = q.xs * (float)px + q.xo, fy = q.ys * (float)py + q.yo;
                pm::PixelTri t;
                pm::pixel_tri(t, __ldg(vb + vi0), __ldg(vb + vi1), __ldg(vb + vi2), fx, fy);
                float gx[3], gy[3], gw[3];
                pm::raster_grad<true>(t, fx, fy, q.xs, q.ys, gr.x, gr.y, gdb, nz_db != 0u, gx, gy, gw);
                #pragma unroll
                for (int k = 0; k < 3; k++) { gv9[3 * k] = gx[k]; gv9[3 * k + 1] = gy[k]; gv9[3 * k + 2] = gw[k]; }
            }
            // same runs as above (key = triangle id and image; attribute and position index buffers may differ, the triangle does
            // not): lanes that were skipped contribute zeros, the head writes when the triangle's position indices are valid
            #pragma unroll
            for (int i = 0; i < 9; i++) gv9[i] = warp_run_sum(gv9[i], runs);
            if (head && !(vi0 < 0 || vi0 >= q.V || vi1 < 0 || vi1 >= q.V || vi2 < 0 || vi2 >= q.V))
            {
                float* gp = q.g_pos + (size_t)n * q.V * 4;
                red_add4(gp + 4 * vi0, make_float4(gv9[0], gv9[1], 0.f, gv9[2]));
                red_add4(gp + 4 * vi1, make_float4(gv9[3], gv9[4], 0.f, gv9[5]));
                red_add4(gp + 4 * vi2, make_float4(gv9[6], gv9[7], 0.f, gv9[8]));
            }
        }
    };

    if (STRIPS)
    {
        // strip s = (image, strip row, strip column); the warp holds its current strip and the one after it (whose first
        // group is the prefetch target of the current strip's last row)
        const unsigned gpr = (unsigned)p.W >> 5, spi = gpr * ((unsigned)p.H / TCG_STRIP_ROWS), nstrips = spi * (unsigned)p.N;
        auto grab = [&]() { unsigned s = 0; if (lane == 0) s = atomicAdd(q.strip_counter, 1u); return __shfl_sync(0xffffffffu, s, 0); };
        unsigned cur = grab();
        while (cur < nstrips)                           // warp-uniform
        {
            const unsigned nxt = grab();
            const unsigned n = cur / spi, rem = cur - n * spi, sy = rem / gpr, sx = rem - sy * gpr;
            unsigned g_nxt = 0xffffffffu;                // group indices fit 32 bits on this path (checked by the host code)
            if (nxt < nstrips)
            {
                const unsigned n2 = nxt / spi, rem2 = nxt - n2 * spi, sy2 = rem2 / gpr, sx2 = rem2 - sy2 * gpr;
                g_nxt = (n2 * (unsigned)p.H + sy2 * TCG_STRIP_ROWS) * gpr + sx2;
            }
            const unsigned y0 = sy * TCG_STRIP_ROWS, px = sx * 32 + (unsigned)lane;
            const unsigned g0 = (n * (unsigned)p.H + y0) * gpr + sx;
            #pragma unroll 1
            for (unsigned row = 0; row < TCG_STRIP_ROWS; row++)
            {
                const unsigned g = g0 + row * gpr, gn = (row + 1 < TCG_STRIP_ROWS) ? g + gpr : g_nxt;
                process((size_t)g, (int)n, (y0 + row) * (unsigned)p.W + px, px, y0 + row, gn == 0xffffffffu ? NO_GROUP : (size_t)gn);
            }
            cur = nxt;
        }
    }
    else
    {
        // image / in-image offset of the warp's current group, advanced incrementally (no 64-bit division per pixel); hw < 2^30
        size_t g = (size_t)blockIdx.x * (TEXG_THREADS / 32) + warp;
        int n0 = (int)((g * 32) / hw);
        unsigned r0 = (unsigned)(g * 32 - (size_t)n0 * hw);
        const unsigned hw32 = (unsigned)hw;
        for (; g < groups; g += wstride)
        {
            unsigned rl = r0 + (unsigned)lane;          // pixel index inside image n
            int n = n0;
            if (rl >= hw32) { rl -= hw32; n++; }
            {   // advance the warp's (image, offset) to its next group
                size_t adv = (size_t)r0 + wstride * 32;
                while (adv >= hw) { adv -= hw; n0++; }
                r0 = (unsigned)adv;
            }
            process(g, n, rl, 0u, 0u, g + wstride);
        }
    }
    pending_flush<C>(pd, p, lane);
}

// ================================================================================================
// Cube maps (csrc/common/texture_kernel.cu:31-317, boundary_mode 'cube'): tex [D,6,w,w,C], uv = 3-D directions,
// uv_da = d(direction)/d(X,Y) as three (dX,dY) pairs.  Separate, simple kernels (one pixel per thread, C <= 4 through the
// padded pyramid); cube-face slices are addressed as depth 6*tz + face.  Face geometry is one table of (major, s, t axes
// and signs) and the seam / corner handling is derived geometrically: the texel one step outside a face is projected onto
// the neighbouring face in exact integer arithmetic (the reference uses two lookup tables, texture_kernel.cu:31-92).
// ================================================================================================

struct CubeFace { int M, S, T; float ss, st; };

static __device__ __forceinline__ CubeFace cube_face(int idx)
{
    CubeFace f;
    f.M = idx >> 1;
    f.S = (f.M == 0) ? 2 : 0;
    f.T = (f.M == 1) ? 2 : 1;
    f.ss = (idx == 0 || idx == 5) ? -1.f : 1.f;
    f.st = (idx == 2) ? 1.f : -1.f;
    return f;
}
static __device__ __forceinline__ float comp3(float3 v, int a) { return a == 0 ? v.x : (a == 1 ? v.y : v.z); }
static __device__ __forceinline__ void set3(float3& v, int a, float x) { if (a == 0) v.x = x; else if (a == 1) v.y = x; else v.z = x; }

static __device__ __forceinline__ int cube_face_of(float3 v)
{
    float ax = fabsf(v.x), ay = fabsf(v.y), az = fabsf(v.z);
    int idx; float c;
    if (az > fmaxf(ax, ay)) { idx = 4; c = v.z; }
    else if (ay > ax)       { idx = 2; c = v.y; }
    else                    { idx = 0; c = v.x; }
    if (c < 0.f) idx += 1;
    return idx;
}

// direction -> face (or -1 for an invalid direction) and (s,t) in [0,1]
static __device__ __forceinline__ int cube_st(float3 v, float& s, float& t)
{
    int idx = cube_face_of(v);
    CubeFace f = cube_face(idx);
    float m = __frcp_rz(fabsf(comp3(v, f.M))) * .5f;
    float x = __fadd_rn(__fmul_rn(comp3(v, f.S), f.ss * m), .5f);
    float y = __fadd_rn(__fmul_rn(comp3(v, f.T), f.st * m), .5f);
    if (!isfinite(x) || !isfinite(y)) return -1;
    s = fminf(fmaxf(x, 0.f), 1.f);
    t = fminf(fmaxf(y, 0.f), 1.f);
    return idx;
}

// texel (ix,iy) of face f, possibly one step outside it: linear index x + w*(y + w*face) on the proper face, -1 at a corner
static __device__ __forceinline__ int cube_texel(int f, int ix, int iy, int w)
{
    bool xo = (ix < 0 || ix >= w), yo = (iy < 0 || iy >= w);
    if (!xo && !yo) return ix + w * (iy + w * f);
    if (xo && yo) return -1;
    int a = 2 * ix + 1 - w, b = 2 * iy + 1 - w;
    int dx, dy, dz;
    switch (f) {
    case 0: dx =  w; dy = -b; dz = -a; break;
    case 1: dx = -w; dy = -b; dz =  a; break;
    case 2: dx =  a; dy =  w; dz =  b; break;
    case 3: dx =  a; dy = -w; dz = -b; break;
    case 4: dx =  a; dy = -b; dz =  w; break;
    default: dx = -a; dy = -b; dz = -w; break;
    }
    int M = (abs(dx) > w) ? 0 : (abs(dy) > w) ? 1 : 2;
    int dM = (M == 0) ? dx : (M == 1) ? dy : dz;
    int f2 = 2 * M + (dM < 0 ? 1 : 0);
    int a2, b2;
    switch (f2) {
    case 0: a2 = -dz; b2 = -dy; break;
    case 1: a2 =  dz; b2 = -dy; break;
    case 2: a2 =  dx; b2 =  dz; break;
    case 3: a2 =  dx; b2 = -dz; break;
    case 4: a2 =  dx; b2 = -dy; break;
    default: a2 = -dx; b2 = -dy; break;
    }
    int x2 = (abs(a2) >= w) ? (a2 > 0 ? w - 1 : 0) : ((a2 + w - 1) >> 1);
    int y2 = (abs(b2) >= w) ? (b2 > 0 ? w - 1 : 0) : ((b2 + w - 1) >> 1);
    return x2 + w * (y2 + w * f2);
}

// J^T (gu, gv): gradient w.r.t. the direction from the gradient w.r.t. (s,t)
static __device__ __forceinline__ float3 cube_grad(float3 v, float gu, float gv)
{
    CubeFace f = cube_face(cube_face_of(v));
    float c = comp3(v, f.M);
    float m = __frcp_rz(fabsf(c));
    float c0 = gu * comp3(v, f.S) * f.ss, c1 = gv * comp3(v, f.T) * f.st;
    float gl = -(c0 + c1) * m * (c < 0.f ? -1.f : 1.f);
    float3 r;
    set3(r, f.S, gu * f.ss); set3(r, f.T, gv * f.st); set3(r, f.M, gl);
    float h = m * .5f;
    r.x *= h; r.y *= h; r.z *= h;
    if (!isfinite(r.x) || !isfinite(r.y) || !isfinite(r.z)) return make_float3(0.f, 0.f, 0.f);
    return r;
}

// (ds/dX, ds/dY, dt/dX, dt/dY) from d(direction)/d(X,Y)
static __device__ __forceinline__ float4 cube_grad_st(float3 v, float3 dX, float3 dY)
{
    CubeFace f = cube_face(cube_face_of(v));
    float c = comp3(v, f.M);
    float m = __frcp_rz(fabsf(c));
    float dm = m * .5f;
    float sg = (c < 0.f) ? -1.f : 1.f;
    float ku = -f.ss * comp3(v, f.S) * m * dm * sg;
    float kv = -f.st * comp3(v, f.T) * m * dm * sg;
    float4 r = make_float4(f.ss * dm * comp3(dX, f.S) + ku * comp3(dX, f.M), f.ss * dm * comp3(dY, f.S) + ku * comp3(dY, f.M),
                           f.st * dm * comp3(dX, f.T) + kv * comp3(dX, f.M), f.st * dm * comp3(dY, f.T) + kv * comp3(dY, f.M));
    if (!isfinite(r.x) || !isfinite(r.y) || !isfinite(r.z) || !isfinite(r.w)) return make_float4(0.f, 0.f, 0.f, 0.f);
    return r;
}

// d(ds/dX, ds/dY, dt/dX, dt/dY)/d(direction) contracted with dfdw
static __device__ __forceinline__ float3 cube_grad2_dot(float3 v, float3 dX, float3 dY, float4 w)
{
    CubeFace f = cube_face(cube_face_of(v));
    float c = comp3(v, f.M);
    float ic = 1.f / c;
    float h = .5f * fabsf(ic);
    float vS = comp3(v, f.S), vT = comp3(v, f.T);
    float oS = 0.f, oT = 0.f, oM = 0.f;
    {
        float SX = comp3(dX, f.S), TX = comp3(dX, f.T), cX = comp3(dX, f.M);
        oS += w.x * (-f.ss * h * cX * ic); oT += w.z * (-f.st * h * cX * ic);
        oM += w.x * (f.ss * h * ic * (-SX + 2.f * vS * cX * ic)) + w.z * (f.st * h * ic * (-TX + 2.f * vT * cX * ic));
    }
    {
        float SX = comp3(dY, f.S), TX = comp3(dY, f.T), cX = comp3(dY, f.M);
        oS += w.y * (-f.ss * h * cX * ic); oT += w.w * (-f.st * h * cX * ic);
        oM += w.y * (f.ss * h * ic * (-SX + 2.f * vS * cX * ic)) + w.w * (f.st * h * ic * (-TX + 2.f * vT * cX * ic));
    }
    float3 r;
    set3(r, f.S, oS); set3(r, f.T, oT); set3(r, f.M, oM);
    return r;
}

// bilinear footprint on a cube level: 4 texel indices (neighbouring faces across seams, -1 missing), weights, corner flag
static __device__ __forceinline__ float2 cube_index_linear(const TexParams& p, float3 v, int tz, int level, int4& tc, bool& corner)
{
    int w = lvl_w(p, level);
    float s, t;
    int face = cube_st(v, s, t);
    if (face < 0) { tc = make_int4(-1, -1, -1, -1); corner = true; return make_float2(0.f, 0.f); }
    float u = s * (float)w - 0.5f, vv = t * (float)w - 0.5f;
    int iu0 = __float2int_rd(u), iv0 = __float2int_rd(vv);
    u -= (float)iu0; vv -= (float)iv0;
    int base = 6 * tz * w * w;
    int q0 = cube_texel(face, iu0, iv0, w), q1 = cube_texel(face, iu0 + 1, iv0, w), q2 = cube_texel(face, iu0, iv0 + 1, w), q3 = cube_texel(face, iu0 + 1, iv0 + 1, w);
    // Parity with the reference (texture_kernel.cu:87-91, 430-435): the missing corner texel is texel (0,0) of "face -1"
    // (index -w*w) before the minibatch offset is added, so only tz == 0 sees a negative index (and the corner rule); for
    // tz > 0 the tap lands on texel (0,0) of the previous item's -z face.
    const int miss = -w * w;
    tc = make_int4((q0 < 0 ? miss : q0) + base, (q1 < 0 ? miss : q1) + base, (q2 < 0 ? miss : q2) + base, (q3 < 0 ? miss : q3) + base);
    corner = (tc.x | tc.y | tc.z | tc.w) < 0;
    return make_float2(u, vv);
}

// four taps with the corner rule: the missing texel is the average of the other three (texture_kernel.cu:590-612)
static __device__ __forceinline__ void cube_fetch(const float* t, int4 tc, bool corner, float4& a, float4& b, float4& c, float4& d, int Cp = 4, int k = 0)
{
    a = texel4(t, tc.x, Cp, k); b = texel4(t, tc.y, Cp, k); c = texel4(t, tc.z, Cp, k); d = texel4(t, tc.w, Cp, k);
    if (corner)
    {
        float4 avg = make_float4((a.x + b.x + c.x + d.x) * 0.33333333f, (a.y + b.y + c.y + d.y) * 0.33333333f,
                                 (a.z + b.z + c.z + d.z) * 0.33333333f, (a.w + b.w + c.w + d.w) * 0.33333333f);
        if (tc.x < 0) a = avg;
        if (tc.y < 0) b = avg;
        if (tc.z < 0) c = avg;
        if (tc.w < 0) d = avg;
    }
}

// scatter weights with the corner rule: the missing tap's weight is shared by the other three (texture_kernel.cu:614-639)
static __device__ __forceinline__ float4 cube_weights(float4 w, int4 tc, bool corner)
{
    if (!corner) return w;
    float wm = (tc.x < 0 ? w.x : 0.f) + (tc.y < 0 ? w.y : 0.f) + (tc.z < 0 ? w.z : 0.f) + (tc.w < 0 ? w.w : 0.f);
    wm *= 0.33333333f;
    return make_float4(w.x + wm, w.y + wm, w.z + wm, w.w + wm);
}

static __device__ __forceinline__ int cube_nearest(const TexParams& p, float3 v, int tz)
{
    float s, t;
    int face = cube_st(v, s, t);
    if (face < 0) return -1;
    int w = p.texW;
    int iu = min(max(__float2int_rd(s * (float)w), 0), w - 1), iv = min(max(__float2int_rd(t * (float)w), 0), w - 1);
    return iu + w * (iv + (6 * tz + face) * w);
}

static __device__ __forceinline__ void cube_load_dirs(const TexParams& p, size_t pidx, float3& v, float3& dX, float3& dY)
{
    const float* q = p.uv + pidx * 3;
    v = make_float3(ld_stream_f(q), ld_stream_f(q + 1), ld_stream_f(q + 2));
    dX = dY = make_float3(0.f, 0.f, 0.f);
    if (p.uv_da)
    {
        const float2* d = (const float2*)p.uv_da + pidx * 3;
        float2 d0 = ld_stream_f2(d), d1 = ld_stream_f2(d + 1), d2 = ld_stream_f2(d + 2);
        dX = make_float3(d0.x, d1.x, d2.x); dY = make_float3(d0.y, d1.y, d2.y);
    }
}

template <int C>
__global__ void __launch_bounds__(256) texture_cube_fwd_kernel(const TexParams p)
{
    const TexMode m = tex_mode(p);
    const size_t hw = (size_t)p.H * p.W, total = hw * p.N;
    for (size_t pidx = (size_t)blockIdx.x * 256 + threadIdx.x; pidx < total; pidx += (size_t)gridDim.x * 256)
    {
        int tz = (p.texD == 1) ? 0 : (int)(pidx / hw);
        float3 v, dX, dY;
        cube_load_dirs(p, pidx, v, dX, dY);
        float o[C];
        if (m.filter == B200DR_FILTER_NEAREST) unpack4<C>(texel4(p.tex[0], cube_nearest(p, v, tz), 4, 0), o);
        else
        {
            int level0, level1; float flevel; float4 dw;
            float bias = p.bias ? ld_stream_f(p.bias + pidx) : 0.f;
            mip_level_from<false>(p, m, cube_grad_st(v, dX, dY), bias, level0, level1, flevel, dw);
            int4 tc0; bool cor0;
            float2 uv0 = cube_index_linear(p, v, tz, level0, tc0, cor0);
            float4 a, b, c, d;
            cube_fetch(p.tex[level0], tc0, cor0, a, b, c, d);
            float4 r = bilerp4(a, b, c, d, uv0);
            if (m.filter == B200DR_FILTER_LINEAR_MIPMAP_LINEAR && flevel > 0.f)
            {
                int4 tc1; bool cor1;
                float2 uv1 = cube_index_linear(p, v, tz, level1, tc1, cor1);
                cube_fetch(p.tex[level1], tc1, cor1, a, b, c, d);
                float4 r1 = bilerp4(a, b, c, d, uv1);
                r = make_float4(lerpf(r.x, r1.x, flevel), lerpf(r.y, r1.y, flevel), lerpf(r.z, r1.z, flevel), lerpf(r.w, r1.w, flevel));
            }
            unpack4<C>(r, o);
        }
        #pragma unroll
        for (int i = 0; i < C; i++) p.out[pidx * C + i] = o[i];
    }
}

// One bilinear tap of the cube gradient scatter, warp-cooperative (every lane calls; inactive lanes pass tc = -1).  Rendered
// images contain regions whose pixels land on the same texel (a constant background direction; strongly magnified maps):
// `reduce` sums runs of equal (level, texel) across consecutive lanes first, so a texel receives one vector reduction per run
// instead of one same-address atomic per pixel (those serialise chip-wide).
template <int C>
static __device__ __forceinline__ void cube_scatter_warp(float* g, int level, int tc, float w, const float (&dy)[C], bool reduce)
{
    if (!reduce) { scatter_tap<C>(g, tc, w, dy); return; }
    const WarpRuns runs = warp_runs(tc >= 0 ? (tc ^ (level << 27)) : -1 - (int)(threadIdx.x & 31));
    float v[C];
    #pragma unroll
    for (int i = 0; i < C; i++) v[i] = warp_run_sum(tc >= 0 ? w * dy[i] : 0.f, runs);
    if (runs.is_head) scatter_tap<C>(g, tc, 1.f, v);
}

template <int C>
__global__ void __launch_bounds__(256) texture_cube_grad_kernel(const TexParams p)
{
    const TexMode m = tex_mode(p);
    const size_t hw = (size_t)p.H * p.W, total = hw * p.N;
    const bool trilinear = (m.filter == B200DR_FILTER_LINEAR_MIPMAP_LINEAR);
    const int lane = threadIdx.x & 31;
    // whole warps iterate together (warp collectives in the scatter): the loop runs over 32-pixel groups
    for (size_t base = ((size_t)blockIdx.x * 256 + threadIdx.x) - lane; base < total; base += (size_t)gridDim.x * 256)
    {
        const size_t pidx = base + lane;
        const bool inb = pidx < total;
        float dy[C];
        unsigned dmax = 0u;
        #pragma unroll
        for (int i = 0; i < C; i++) { dy[i] = inb ? p.dy[pidx * C + i] : 0.f; dmax |= __float_as_uint(dy[i]); }
        const bool act = inb && (__uint_as_float(dmax) != 0.f);
        float3 g_v = make_float3(0.f, 0.f, 0.f), g0 = g_v, g1 = g_v;
        float df = 0.f;
        // per-lane footprint; inactive lanes keep tc = -1 and flow through the warp-cooperative scatter
        int level0 = 0, level1 = 0; float flevel = 0.f; float4 dw = make_float4(0.f, 0.f, 0.f, 0.f);
        int4 tc0 = make_int4(-1, -1, -1, -1), tc1 = tc0; bool cor0 = false, cor1 = false;
        float2 uv0 = make_float2(0.f, 0.f), uv1 = uv0;
        float4 w0 = make_float4(0.f, 0.f, 0.f, 0.f), w1 = w0;
        float3 v = make_float3(0.f, 0.f, 1.f), dX = make_float3(0.f, 0.f, 0.f), dY = dX, dfdv = dX;
        bool two = false; float f0 = 1.f;
        if (act)
        {
            int tz = (p.texD == 1) ? 0 : (int)(pidx / hw);
            cube_load_dirs(p, pidx, v, dX, dY);
            if (m.filter == B200DR_FILTER_NEAREST) { tc0.x = cube_nearest(p, v, tz); w0.x = 1.f; }
            else
            {
                float bias = p.bias ? ld_stream_f(p.bias + pidx) : 0.f;
                mip_level_from<true>(p, m, cube_grad_st(v, dX, dY), bias, level0, level1, flevel, dw);
                if (trilinear && p.uv_da)
                {
                    dfdv = cube_grad2_dot(v, dX, dY, dw);
                    if (!isfinite(dfdv.x) || !isfinite(dfdv.y) || !isfinite(dfdv.z)) { dfdv = make_float3(0.f, 0.f, 0.f); dw = make_float4(0.f, 0.f, 0.f, 0.f); }
                }
                two = trilinear && (flevel > 0.f);
                f0 = trilinear ? (1.f - flevel) : 1.f;
                uv0 = cube_index_linear(p, v, tz, level0, tc0, cor0);
                float uv011 = uv0.x * uv0.y, uv010 = uv0.x - uv011, uv001 = uv0.y - uv011, uv000 = 1.f - uv0.x - uv001;
                w0 = cube_weights(make_float4(uv000 * f0, uv010 * f0, uv001 * f0, uv011 * f0), tc0, cor0);
                if (two)
                {
                    uv1 = cube_index_linear(p, v, tz, level1, tc1, cor1);
                    float uv111 = uv1.x * uv1.y, uv110 = uv1.x - uv111, uv101 = uv1.y - uv111, uv100 = 1.f - uv1.x - uv101;
                    w1 = cube_weights(make_float4(uv100 * flevel, uv110 * flevel, uv101 * flevel, uv111 * flevel), tc1, cor1);
                }
                else tc1 = make_int4(-1, -1, -1, -1);
            }
        }
        // ---- texel gradients, warp-uniform control flow
        if (__any_sync(0xffffffffu, act))
        {
            const bool reduce = __popc(warp_runs(act ? (tc0.x ^ (level0 << 27)) : -1 - lane).heads) <= 12;   // few distinct texels in this group
            float* gp0 = p.gpad[level0];        // level differs per lane: the key carries it, the pointer is the lane's own
            cube_scatter_warp<C>(gp0, level0, tc0.x, w0.x, dy, reduce);
            if (m.filter != B200DR_FILTER_NEAREST)
            {
                cube_scatter_warp<C>(gp0, level0, tc0.y, w0.y, dy, reduce);
                cube_scatter_warp<C>(gp0, level0, tc0.z, w0.z, dy, reduce);
                cube_scatter_warp<C>(gp0, level0, tc0.w, w0.w, dy, reduce);
                if (__any_sync(0xffffffffu, two))
                {
                    float* gp1 = p.gpad[level1];
                    cube_scatter_warp<C>(gp1, level1, tc1.x, w1.x, dy, reduce);
                    cube_scatter_warp<C>(gp1, level1, tc1.y, w1.y, dy, reduce);
                    cube_scatter_warp<C>(gp1, level1, tc1.z, w1.z, dy, reduce);
                    cube_scatter_warp<C>(gp1, level1, tc1.w, w1.w, dy, reduce);
                }
            }
        }
        // ---- direction / level gradients
        if (act && m.filter != B200DR_FILTER_NEAREST)
        {
            float gu = 0.f, gv = 0.f;
            float4 A00, A10, A01, A11;
            cube_fetch(p.tex[level0], tc0, cor0, A00, A10, A01, A11);
            float a00[C], a10[C], a01[C], a11[C];
            unpack4<C>(A00, a00); unpack4<C>(A10, a10); unpack4<C>(A01, a01); unpack4<C>(A11, a11);
            float sclu0 = (float)lvl_w(p, level0);
            #pragma unroll
            for (int i = 0; i < C; i++)
            {
                float dy0 = f0 * dy[i];
                float ad = (a11[i] + a00[i] - a10[i] - a01[i]);
                gu += dy0 * ((a10[i] - a00[i]) + uv0.y * ad) * sclu0;
                gv += dy0 * ((a01[i] - a00[i]) + uv0.x * ad) * sclu0;
            }
            if (two)
            {
                float4 B00, B10, B01, B11;
                cube_fetch(p.tex[level1], tc1, cor1, B00, B10, B01, B11);
                float b00[C], b10[C], b01[C], b11[C];
                unpack4<C>(B00, b00); unpack4<C>(B10, b10); unpack4<C>(B01, b01); unpack4<C>(B11, b11);
                float sclu1 = (float)lvl_w(p, level1);
                #pragma unroll
                for (int i = 0; i < C; i++)
                {
                    float dy1 = flevel * dy[i];
                    float bd = (b11[i] + b00[i] - b10[i] - b01[i]);
                    gu += dy1 * ((b10[i] - b00[i]) + uv1.y * bd) * sclu1;
                    gv += dy1 * ((b01[i] - b00[i]) + uv1.x * bd) * sclu1;
                    df += (bilerp(b00[i], b10[i], b01[i], b11[i], uv1) - bilerp(a00[i], a10[i], a01[i], a11[i], uv0)) * dy[i];
                }
            }
            g_v = cube_grad(v, gu, gv);
            if (trilinear) { g_v.x += dfdv.x * df; g_v.y += dfdv.y * df; g_v.z += dfdv.z * df; }
            if (trilinear && p.uv_da)
            {
                g0 = cube_grad(v, dw.x * df, dw.z * df);
                g1 = cube_grad(v, dw.y * df, dw.w * df);
            }
        }
        if (inb)
        {
            if (p.g_uv) { p.g_uv[pidx * 3] = g_v.x; p.g_uv[pidx * 3 + 1] = g_v.y; p.g_uv[pidx * 3 + 2] = g_v.z; }
            if (p.g_bias) p.g_bias[pidx] = df;
            if (p.g_uv_da)
            {
                float2* q = (float2*)p.g_uv_da + pidx * 3;
                q[0] = make_float2(g0.x, g1.x); q[1] = make_float2(g0.y, g1.y); q[2] = make_float2(g0.z, g1.z);
            }
        }
    }
}

// ---- cube maps with any channel count: same addressing, the texel is walked in 128-bit channel groups of the padded
// pyramid (the reference serves every channel count through vector widths 1/2/4, csrc/torch/torch_texture.cpp:344-401).
__global__ void __launch_bounds__(256) texture_cube_fwd_generic(const TexParams p)
{
    const TexMode m = tex_mode(p);
    const size_t hw = (size_t)p.H * p.W, total = hw * p.N;
    const int C = p.C, Cp = p.Cp;
    for (size_t pidx = (size_t)blockIdx.x * 256 + threadIdx.x; pidx < total; pidx += (size_t)gridDim.x * 256)
    {
        int tz = (p.texD == 1) ? 0 : (int)(pidx / hw);
        float3 v, dX, dY;
        cube_load_dirs(p, pidx, v, dX, dY);
        float* o = p.out + pidx * C;
        if (m.filter == B200DR_FILTER_NEAREST)
        {
            int tc = cube_nearest(p, v, tz);
            for (int k = 0; k < Cp; k += 4) put4(o, k, C, texel4(p.tex[0], tc, Cp, k));
            continue;
        }
        int level0, level1; float flevel; float4 dw;
        float bias = p.bias ? ld_stream_f(p.bias + pidx) : 0.f;
        mip_level_from<false>(p, m, cube_grad_st(v, dX, dY), bias, level0, level1, flevel, dw);
        int4 tc0, tc1 = make_int4(-1, -1, -1, -1); bool cor0, cor1 = false;
        float2 uv0 = cube_index_linear(p, v, tz, level0, tc0, cor0), uv1 = make_float2(0.f, 0.f);
        const bool two = (m.filter == B200DR_FILTER_LINEAR_MIPMAP_LINEAR && flevel > 0.f);
        if (two) uv1 = cube_index_linear(p, v, tz, level1, tc1, cor1);
        for (int k = 0; k < Cp; k += 4)
        {
            float4 a, b, c, d;
            cube_fetch(p.tex[level0], tc0, cor0, a, b, c, d, Cp, k);
            float4 r = bilerp4(a, b, c, d, uv0);
            if (two)
            {
                cube_fetch(p.tex[level1], tc1, cor1, a, b, c, d, Cp, k);
                float4 r1 = bilerp4(a, b, c, d, uv1);
                r = make_float4(lerpf(r.x, r1.x, flevel), lerpf(r.y, r1.y, flevel), lerpf(r.z, r1.z, flevel), lerpf(r.w, r1.w, flevel));
            }
            put4(o, k, C, r);
        }
    }
}

static __device__ __forceinline__ void scatter_group(float* g, int tc, int Cp, int k, float w, float4 dy)
{
    if (tc < 0) return;
    red_add4(g + (size_t)tc * Cp + k, make_float4(w * dy.x, w * dy.y, w * dy.z, w * dy.w));
}

__global__ void __launch_bounds__(256) texture_cube_grad_generic(const TexParams p)
{
    const TexMode m = tex_mode(p);
    const size_t hw = (size_t)p.H * p.W, total = hw * p.N;
    const bool trilinear = (m.filter == B200DR_FILTER_LINEAR_MIPMAP_LINEAR);
    const int C = p.C, Cp = p.Cp;
    for (size_t pidx = (size_t)blockIdx.x * 256 + threadIdx.x; pidx < total; pidx += (size_t)gridDim.x * 256)
    {
        const float* pdy = p.dy + pidx * C;
        unsigned dmax = 0u;
        for (int i = 0; i < C; i++) dmax |= __float_as_uint(pdy[i]);
        float3 g_v = make_float3(0.f, 0.f, 0.f), g0 = g_v, g1 = g_v;
        float df = 0.f;
        if (__uint_as_float(dmax) != 0.f)
        {
            int tz = (p.texD == 1) ? 0 : (int)(pidx / hw);
            float3 v, dX, dY;
            cube_load_dirs(p, pidx, v, dX, dY);
            if (m.filter == B200DR_FILTER_NEAREST)
            {
                int tc = cube_nearest(p, v, tz);
                for (int k = 0; k < Cp; k += 4)
                    scatter_group(p.gpad[0], tc, Cp, k, 1.f, make_float4(pdy[k], (k + 1 < C) ? pdy[k + 1] : 0.f, (k + 2 < C) ? pdy[k + 2] : 0.f, (k + 3 < C) ? pdy[k + 3] : 0.f));
            }
            else
            {
                int level0, level1; float flevel; float4 dw = make_float4(0.f, 0.f, 0.f, 0.f);
                float bias = p.bias ? ld_stream_f(p.bias + pidx) : 0.f;
                mip_level_from<true>(p, m, cube_grad_st(v, dX, dY), bias, level0, level1, flevel, dw);
                float3 dfdv = make_float3(0.f, 0.f, 0.f);
                if (trilinear && p.uv_da)
                {
                    dfdv = cube_grad2_dot(v, dX, dY, dw);
                    if (!isfinite(dfdv.x) || !isfinite(dfdv.y) || !isfinite(dfdv.z)) { dfdv = make_float3(0.f, 0.f, 0.f); dw = make_float4(0.f, 0.f, 0.f, 0.f); }
                }
                const bool two = trilinear && (flevel > 0.f);
                const float f0 = trilinear ? (1.f - flevel) : 1.f;
                float gu = 0.f, gv = 0.f;
                int4 tc0, tc1 = make_int4(-1, -1, -1, -1); bool cor0, cor1 = false;
                float2 uv0 = cube_index_linear(p, v, tz, level0, tc0, cor0), uv1 = make_float2(0.f, 0.f);
                if (two) uv1 = cube_index_linear(p, v, tz, level1, tc1, cor1);
                float uv011 = uv0.x * uv0.y, uv010 = uv0.x - uv011, uv001 = uv0.y - uv011, uv000 = 1.f - uv0.x - uv001;
                const float4 w0 = cube_weights(make_float4(uv000 * f0, uv010 * f0, uv001 * f0, uv011 * f0), tc0, cor0);
                float uv111 = uv1.x * uv1.y, uv110 = uv1.x - uv111, uv101 = uv1.y - uv111, uv100 = 1.f - uv1.x - uv101;
                const float4 w1 = cube_weights(make_float4(uv100 * flevel, uv110 * flevel, uv101 * flevel, uv111 * flevel), tc1, cor1);
                const float sclu0 = (float)lvl_w(p, level0), sclu1 = (float)lvl_w(p, level1);
                for (int k = 0; k < Cp; k += 4)
                {
                    const float4 dy4 = make_float4(pdy[k], (k + 1 < C) ? pdy[k + 1] : 0.f, (k + 2 < C) ? pdy[k + 2] : 0.f, (k + 3 < C) ? pdy[k + 3] : 0.f);
                    float* gp0 = p.gpad[level0];
                    scatter_group(gp0, tc0.x, Cp, k, w0.x, dy4); scatter_group(gp0, tc0.y, Cp, k, w0.y, dy4);
                    scatter_group(gp0, tc0.z, Cp, k, w0.z, dy4); scatter_group(gp0, tc0.w, Cp, k, w0.w, dy4);
                    float4 A00, A10, A01, A11, B00, B10, B01, B11;
                    cube_fetch(p.tex[level0], tc0, cor0, A00, A10, A01, A11, Cp, k);
                    if (two)
                    {
                        float* gp1 = p.gpad[level1];
                        scatter_group(gp1, tc1.x, Cp, k, w1.x, dy4); scatter_group(gp1, tc1.y, Cp, k, w1.y, dy4);
                        scatter_group(gp1, tc1.z, Cp, k, w1.z, dy4); scatter_group(gp1, tc1.w, Cp, k, w1.w, dy4);
                        cube_fetch(p.tex[level1], tc1, cor1, B00, B10, B01, B11, Cp, k);
                    }
                    #pragma unroll
                    for (int j = 0; j < 4; j++)
                    {
                        if (k + j >= C) break;
                        const float dyc = (&dy4.x)[j];
                        const float a00 = (&A00.x)[j], a10 = (&A10.x)[j], a01 = (&A01.x)[j], a11 = (&A11.x)[j];
                        const float ad = (a11 + a00 - a10 - a01), dy0 = f0 * dyc;
                        gu += dy0 * ((a10 - a00) + uv0.y * ad) * sclu0;
                        gv += dy0 * ((a01 - a00) + uv0.x * ad) * sclu0;
                        if (two)
                        {
                            const float b00 = (&B00.x)[j], b10 = (&B10.x)[j], b01 = (&B01.x)[j], b11 = (&B11.x)[j];
                            const float bd = (b11 + b00 - b10 - b01), dy1 = flevel * dyc;
                            gu += dy1 * ((b10 - b00) + uv1.y * bd) * sclu1;
                            gv += dy1 * ((b01 - b00) + uv1.x * bd) * sclu1;
                            df += (bilerp(b00, b10, b01, b11, uv1) - bilerp(a00, a10, a01, a11, uv0)) * dyc;
                        }
                    }
                }
                g_v = cube_grad(v, gu, gv);
                if (trilinear) { g_v.x += dfdv.x * df; g_v.y += dfdv.y * df; g_v.z += dfdv.z * df; }
                if (trilinear && p.uv_da)
                {
                    g0 = cube_grad(v, dw.x * df, dw.z * df);
                    g1 = cube_grad(v, dw.y * df, dw.w * df);
                }
            }
        }
        if (p.g_uv) { p.g_uv[pidx * 3] = g_v.x; p.g_uv[pidx * 3 + 1] = g_v.y; p.g_uv[pidx * 3 + 2] = g_v.z; }
        if (p.g_bias) p.g_bias[pidx] = df;
        if (p.g_uv_da)
        {
            float2* q = (float2*)p.g_uv_da + pidx * 3;
            q[0] = make_float2(g0.x, g1.x); q[1] = make_float2(g0.y, g1.y); q[2] = make_float2(g0.z, g1.z);
        }
    }
}

// Resolve: g_out[texel][c] = pad[0][texel][c] (+ sum over levels of weight * pad[L][parent texel][c] when pulling mips),
// written in the caller's C-channel layout.  Pull weights as in MipGradKernel (texture_kernel.cu:843-900).
struct ResolveParams
{
    const float* gpad[TEX_MAX_MIP + 1];
    float* g_out;
    int texW, texH, texD, C, Cp, mip_max, pull;     // texW/texH: size of the level being resolved
    int level;                                      // index into gpad of the level being resolved
    int baseW, baseH;                               // base texture size (for the pull chain)
};

__global__ void __launch_bounds__(256) texture_grad_resolve_kernel(const ResolveParams p)
{
    size_t total = (size_t)p.texW * p.texH * p.texD * p.C;
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (size_t)gridDim.x * 256)
    {
        int c = (int)(i % p.C); size_t t = i / p.C;
        int x = (int)(t % p.texW); size_t t2 = t / p.texW;
        int y = (int)(t2 % p.texH); int z = (int)(t2 / p.texH);
        float acc = p.gpad[p.level][t * p.Cp + c];
        if (p.pull)
        {
            float w = 1.f;
            int xx = x, yy = y, pw = p.baseW, ph = p.baseH;
            for (int L = 1; L <= p.mip_max; L++)
            {
                if (pw > 1) w *= .5f;
                if (ph > 1) w *= .5f;
                pw = (p.baseW >> L) > 1 ? (p.baseW >> L) : 1;
                ph = (p.baseH >> L) > 1 ? (p.baseH >> L) : 1;
                xx >>= 1; yy >>= 1;
                acc += p.gpad[L][((size_t)xx + (size_t)pw * (yy + (size_t)ph * z)) * p.Cp + c] * w;
            }
        }
        p.g_out[i] = acc;
    }
}

static int tex_fill(TexParams& p, const char* fn, const float* tex, const float* const* mip_ptrs, int max_level,
                    const float* uv, const float* uv_da, const float* bias,
                    int N, int H, int W, int texD, int texH, int texW, int C, int filter, int boundary)
{
    B200DR_CHECK(filter >= 0 && filter < 4, "%s: filter_mode unsupported", fn);
    B200DR_CHECK(boundary >= 0 && boundary < 4, "%s: boundary_mode unsupported", fn);
    const bool cube = (boundary == B200DR_BOUNDARY_CUBE);
    B200DR_CHECK(!cube || texW == texH, "%s: texture shape must be square in cube map mode", fn);
    B200DR_CHECK(N > 0 && H > 0 && W > 0, "%s: uv must have shape [>0, >0, >0, %d]", fn, cube ? 3 : 2);
    B200DR_CHECK(texD > 0 && texH > 0 && texW > 0 && C > 0, "%s: tex must have shape[>0, >0, >0, >0]", fn);
    B200DR_CHECK(texD == 1 || texD == N, "%s: minibatch size mismatch between inputs tex, uv", fn);
    B200DR_CHECK(texW <= (1 << TEX_MAX_MIP) && texH <= (1 << TEX_MAX_MIP), "%s: texture size too large", fn);
    B200DR_CHECK(max_level >= 0 && max_level <= TEX_MAX_MIP, "%s: invalid mip level count", fn);
    B200DR_CHECK((long long)texD * (cube ? 6 : 1) * texH * texW < (1ll << 27), "%s: more than 2^27 texels per level are not supported", fn);
    bool mip = (filter == B200DR_FILTER_LINEAR_MIPMAP_NEAREST || filter == B200DR_FILTER_LINEAR_MIPMAP_LINEAR);
    B200DR_CHECK(!mip || uv_da || bias, "%s: mipmapping filter mode requires uv_da and/or mip_level_bias input", fn);
    B200DR_CHECK(cube || (((uintptr_t)uv) & 7) == 0, "%s: uv input tensor not aligned to float2", fn);
    B200DR_CHECK(cube ? ((((uintptr_t)uv_da) & 7) == 0) : B200DR_ALIGNED16(uv_da), "%s: uv_da input tensor not aligned to %s", fn, cube ? "float2" : "float4");
    B200DR_CHECK(B200DR_ALIGNED16(tex), "%s: padded texture level 0 not aligned to float4", fn);
    p.tex[0] = tex;
    p.mip_max = mip ? max_level : 0;
    for (int i = 1; i <= p.mip_max; i++)
    {
        B200DR_CHECK(mip_ptrs && mip_ptrs[i], "%s: mipmapping filter mode requires mip wrapper or mip stack input", fn);
        B200DR_CHECK(B200DR_ALIGNED16(mip_ptrs[i]), "%s: padded mip level not aligned to float4", fn);
        p.tex[i] = mip_ptrs[i];
    }
    p.uv = uv; p.uv_da = mip ? uv_da : 0; p.bias = mip ? bias : 0;
    p.N = N; p.H = H; p.W = W; p.texD = texD; p.texH = texH; p.texW = texW; p.C = C; p.Cp = (C + 3) & ~3;
    p.filter = filter; p.boundary = boundary;
    size_t total = (size_t)N * H * W;
    B200DR_CHECK(total < (1ull << 31) * 128, "%s: too many pixels", fn);
    p.chunks = (int)((total + TEX_THREADS - 1) / TEX_THREADS);
    p.tile2d = (TEXG_TILE2D && (W % 8) == 0 && (H % 4) == 0) ? 1 : 0;
    p.tilesX = W / 8; p.tilesPerImg = (W / 8) * (H / 4);
    return 0;
}

static long long mip_info(int texW, int texH, int texD, int C, int limit, int* offsets, int* max_level)
{
    if (limit == 0) { *max_level = 0; offsets[0] = 0; return 0; }
    int w = texW, h = texH, level = 0; long long total = 0;
    offsets[0] = 0;
    while ((w | h) > 1)
    {
        level++;
        if ((w > 1 && (w & 1)) || (h > 1 && (h & 1))) return -1;
        if (w > 1) w >>= 1;
        if (h > 1) h >>= 1;
        offsets[level] = (int)total;
        total += (long long)w * h * texD * C;
        if (limit >= 0 && level == limit) break;
        if (level == TEX_MAX_MIP) break;
    }
    *max_level = level;
    return total;
}

// Fast-kernel dispatch: the common mode pairs (bilinear / trilinear x wrap / clamp) get kernels with the mode switches
// folded at compile time; every other combination runs the same code with run-time (warp-uniform) switches.
template <int F, int B>
static void launch_fast_fwd_c(const TexParams& p, unsigned blocks, cudaStream_t stream)
{
    switch (p.C) {
    case 1: texture_fwd_fast<1, F, B><<<blocks, TEX_THREADS, 0, stream>>>(p); break;
    case 2: texture_fwd_fast<2, F, B><<<blocks, TEX_THREADS, 0, stream>>>(p); break;
    case 3: texture_fwd_fast<3, F, B><<<blocks, TEX_THREADS, 0, stream>>>(p); break;
    case 4: texture_fwd_fast<4, F, B><<<blocks, TEX_THREADS, 0, stream>>>(p); break;
    }
}
template <int F, int B>
static void launch_fast_grad_c(const TexParams& p, unsigned blocks, cudaStream_t stream)
{
    switch (p.C) {
    case 1: texture_grad_fast<1, F, B><<<blocks, TEXG_THREADS, 0, stream>>>(p); break;
    case 2: texture_grad_fast<2, F, B><<<blocks, TEXG_THREADS, 0, stream>>>(p); break;
    case 3: texture_grad_fast<3, F, B><<<blocks, TEXG_THREADS, 0, stream>>>(p); break;
    case 4: texture_grad_fast<4, F, B><<<blocks, TEXG_THREADS, 0, stream>>>(p); break;
    }
}
#define TEX_MODE_DISPATCH(FN)                                                                                                    \
    do {                                                                                                                         \
        const int f = p.filter, b = p.boundary;                                                                                  \
        if      (f == B200DR_FILTER_LINEAR_MIPMAP_LINEAR && b == B200DR_BOUNDARY_WRAP)  FN<B200DR_FILTER_LINEAR_MIPMAP_LINEAR, B200DR_BOUNDARY_WRAP>(p, blocks, stream);  \
        else if (f == B200DR_FILTER_LINEAR_MIPMAP_LINEAR && b == B200DR_BOUNDARY_CLAMP) FN<B200DR_FILTER_LINEAR_MIPMAP_LINEAR, B200DR_BOUNDARY_CLAMP>(p, blocks, stream); \
        else if (f == B200DR_FILTER_LINEAR && b == B200DR_BOUNDARY_WRAP)                FN<B200DR_FILTER_LINEAR, B200DR_BOUNDARY_WRAP>(p, blocks, stream);                \
        else if (f == B200DR_FILTER_LINEAR && b == B200DR_BOUNDARY_CLAMP)               FN<B200DR_FILTER_LINEAR, B200DR_BOUNDARY_CLAMP>(p, blocks, stream);               \
        else                                                                            FN<-1, -1>(p, blocks, stream);                                                    \
    } while (0)
static void launch_fast_fwd(const TexParams& p, unsigned blocks, cudaStream_t stream) { TEX_MODE_DISPATCH(launch_fast_fwd_c); }
static void launch_fast_grad(const TexParams& p, unsigned blocks, cudaStream_t stream) { TEX_MODE_DISPATCH(launch_fast_grad_c); }

} // namespace

extern "C" long long b200dr_texture_mip_info(int texW, int texH, int texD, int C, int max_mip_level, int* offsets, int* max_level)
{
    return mip_info(texW, texH, texD, C, max_mip_level, offsets, max_level);
}

extern "C" int b200dr_texture_construct_mip(const float* tex, float* mip, int texW, int texH, int texD, int C,
                                            const int* offsets, int max_level, void* stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    B200DR_CHECK(texD > 0 && texH > 0 && texW > 0 && C > 0, "texture_construct_mip(): tex must have shape[>0, >0, >0, >0]");
    B200DR_CHECK((C & 3) == 0 && B200DR_ALIGNED16(tex) && B200DR_ALIGNED16(mip), "texture_construct_mip(): expects channel-padded (multiple of 4), 16-byte aligned levels");
    B200DR_CHECK(max_level >= 0 && max_level <= TEX_MAX_MIP, "texture_construct_mip(): invalid max_mip_level");
    for (int L = 1; L <= max_level; L++)
    {
        const float* in = (L == 1) ? tex : mip + offsets[L - 1];
        float* out = mip + offsets[L];
        int wi = (texW >> (L - 1)) > 1 ? (texW >> (L - 1)) : 1, hi = (texH >> (L - 1)) > 1 ? (texH >> (L - 1)) : 1;
        int wo = (texW >> L) > 1 ? (texW >> L) : 1, ho = (texH >> L) > 1 ? (texH >> L) : 1;
        size_t total = (size_t)wo * ho * texD * (C >> 2);
        size_t blocks = (total + 255) / 256;
        size_t cap = (size_t)b200dr_num_sms() * 16;
        if (blocks > cap) blocks = cap;
        if (blocks < 1) blocks = 1;
        mip_build_kernel<<<(unsigned)blocks, 256, 0, stream>>>(in, out, wi, hi, wo, ho, texD, C);
    }
    b200dr_count_launches(max_level - 1);   // max_level kernels, the check below counts one
    B200DR_CHECK_LAUNCH("texture_construct_mip()");
    return 0;
}

extern "C" int b200dr_texture_pad(const float* src, float* dst, long long ntexels, int C, void* stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    B200DR_CHECK(ntexels > 0 && C > 0, "texture_pad(): empty texture");
    B200DR_CHECK(B200DR_ALIGNED16(dst), "texture_pad(): destination not aligned to float4");
    int Cp = (C + 3) & ~3;
    size_t total = (size_t)ntexels * (Cp >> 2);
    size_t blocks = (total + 255) / 256;
    size_t cap = (size_t)b200dr_num_sms() * 16;
    if (blocks > cap) blocks = cap;
    texture_pad_kernel<<<(unsigned)blocks, 256, 0, stream>>>(src, dst, (size_t)ntexels, C, Cp);
    B200DR_CHECK_LAUNCH("texture_pad()");
    return 0;
}

extern "C" int b200dr_texture_fwd(const float* tex, const float* const* mip_ptrs, int max_level,
                                  const float* uv, const float* uv_da, const float* mip_level_bias,
                                  int N, int H, int W, int texD, int texH, int texW, int C,
                                  int filter_mode, int boundary_mode, float* out, void* stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    TexParams p = {};
    if (tex_fill(p, "texture_fwd()", tex, mip_ptrs, max_level, uv, uv_da, mip_level_bias, N, H, W, texD, texH, texW, C, filter_mode, boundary_mode)) return 1;
    p.out = out;
    p.tile2d = 0;                                       // 8x4 pixel tiles only pay in the backward pass (measured: fwd 0.59 -> 0.78 ms)
    if (boundary_mode == B200DR_BOUNDARY_CUBE)
    {
        size_t blocks = ((size_t)N * H * W + 255) / 256;
        size_t bcap = (size_t)b200dr_num_sms() * 16;
        if (blocks > bcap) blocks = bcap;
        switch (C) {
        case 1: texture_cube_fwd_kernel<1><<<(unsigned)blocks, 256, 0, stream>>>(p); break;
        case 2: texture_cube_fwd_kernel<2><<<(unsigned)blocks, 256, 0, stream>>>(p); break;
        case 3: texture_cube_fwd_kernel<3><<<(unsigned)blocks, 256, 0, stream>>>(p); break;
        case 4: texture_cube_fwd_kernel<4><<<(unsigned)blocks, 256, 0, stream>>>(p); break;
        default: texture_cube_fwd_generic<<<(unsigned)blocks, 256, 0, stream>>>(p); break;
        }
        B200DR_CHECK_LAUNCH("texture_fwd()");
        return 0;
    }
    if (C <= 4 && B200DR_ALIGNED16(out))
    {
        size_t groups = ((size_t)N * H * W + 31) / 32;
        size_t blocks = (groups + (TEX_THREADS / 32) - 1) / (TEX_THREADS / 32);
        size_t bcap = (size_t)b200dr_num_sms() * 12;
        if (blocks > bcap) blocks = b200dr_balanced_blocks(bcap, TEX_THREADS / 32, W);
        launch_fast_fwd(p, (unsigned)blocks, stream);
        B200DR_CHECK_LAUNCH("texture_fwd()");
        return 0;
    }
    size_t smem = (size_t)TEX_THREADS * C * sizeof(float);
    p.staged = smem <= 96 * 1024;
    if (!p.staged) smem = 0;
    if (smem > 48 * 1024) B200DR_CUDA(cudaFuncSetAttribute(texture_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024), "texture()");
    int grid = p.chunks;
    int cap = b200dr_num_sms() * 16;
    if (grid > cap) grid = cap;
    texture_fwd_kernel<<<grid, TEX_THREADS, smem, stream>>>(p);
    B200DR_CHECK_LAUNCH("texture_fwd()");
    return 0;
}

// Bytes of the channel-padded gradient pyramid used by b200dr_texture_grad (levels 0..max_level), followed by a
// TEXG_TAIL_BYTES tail of kernel bookkeeping (the fused chain backward's work counter); all of it is zeroed by the op.
#define TEXG_TAIL_BYTES 256
extern "C" size_t b200dr_texture_grad_scratch_bytes(int texD, int texH, int texW, int C, int max_level)
{
    size_t Cp = (size_t)((C + 3) & ~3), total = 0;
    for (int L = 0; L <= max_level; L++)
    {
        size_t w = (texW >> L) > 1 ? (texW >> L) : 1, h = (texH >> L) > 1 ? (texH >> L) : 1;
        total += w * h * texD * Cp * sizeof(float);
    }
    return total + TEXG_TAIL_BYTES;
}

extern "C" int b200dr_texture_grad(const float* tex, const float* const* mip_ptrs, int max_level,
                                   const float* uv, const float* uv_da, const float* mip_level_bias, const float* dy,
                                   int N, int H, int W, int texD, int texH, int texW, int C,
                                   int filter_mode, int boundary_mode,
                                   float* g_tex, float* const* g_mip_ptrs, int pull_mips,
                                   float* g_uv, float* g_uv_da, float* g_bias,
                                   void* scratch, size_t scratch_bytes, void* stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    TexParams p = {};
    if (tex_fill(p, "texture_grad()", tex, mip_ptrs, max_level, uv, uv_da, mip_level_bias, N, H, W, texD, texH, texW, C, filter_mode, boundary_mode)) return 1;
    const bool cube = (boundary_mode == B200DR_BOUNDARY_CUBE);
    const int depth = cube ? 6 * texD : texD;           // cube faces are separate 2-D slices of every level
    B200DR_CHECK(cube || (((uintptr_t)g_uv) & 7) == 0, "texture_grad(): grad_uv output tensor not aligned to float2");
    B200DR_CHECK(cube ? ((((uintptr_t)g_uv_da) & 7) == 0) : B200DR_ALIGNED16(g_uv_da), "texture_grad(): grad_uv_da output tensor misaligned");
    size_t need = b200dr_texture_grad_scratch_bytes(depth, texH, texW, C, p.mip_max);
    B200DR_CHECK(scratch && scratch_bytes >= need && B200DR_ALIGNED16(scratch), "texture_grad(): scratch buffer too small or misaligned");
    p.dy = dy; p.g_uv = (filter_mode != B200DR_FILTER_NEAREST) ? g_uv : 0;
    p.g_uv_da = (filter_mode == B200DR_FILTER_LINEAR_MIPMAP_LINEAR && uv_da) ? g_uv_da : 0;
    p.g_bias = (filter_mode == B200DR_FILTER_LINEAR_MIPMAP_LINEAR && mip_level_bias) ? g_bias : 0;

    // padded gradient pyramid
    B200DR_CUDA(cudaMemsetAsync(scratch, 0, need, stream), "texture()");
    ResolveParams r = {};
    {
        char* s = (char*)scratch;
        for (int L = 0; L <= p.mip_max; L++)
        {
            size_t w = (texW >> L) > 1 ? (texW >> L) : 1, h = (texH >> L) > 1 ? (texH >> L) : 1;
            p.gpad[L] = (float*)s; r.gpad[L] = (const float*)s;
            s += w * h * depth * p.Cp * sizeof(float);
        }
    }

    if (cube)
    {
        size_t blocks = ((size_t)N * H * W + 255) / 256;
        size_t bcap = (size_t)b200dr_num_sms() * 16;
        if (blocks > bcap) blocks = bcap;
        switch (C) {
        case 1: texture_cube_grad_kernel<1><<<(unsigned)blocks, 256, 0, stream>>>(p); break;
        case 2: texture_cube_grad_kernel<2><<<(unsigned)blocks, 256, 0, stream>>>(p); break;
        case 3: texture_cube_grad_kernel<3><<<(unsigned)blocks, 256, 0, stream>>>(p); break;
        case 4: texture_cube_grad_kernel<4><<<(unsigned)blocks, 256, 0, stream>>>(p); break;
        default: texture_cube_grad_generic<<<(unsigned)blocks, 256, 0, stream>>>(p); break;
        }
    }
    else if (C <= 4 && B200DR_ALIGNED16(dy))
    {
        size_t groups = ((size_t)N * H * W + 31) / 32;
        size_t per_cta = (TEXG_THREADS / 32) * TEXG_UNROLL;
        size_t blocks = (groups + per_cta - 1) / per_cta;
        size_t bcap = (size_t)b200dr_num_sms() * 20;
        if (blocks > bcap) blocks = b200dr_balanced_blocks(bcap, TEXG_THREADS / 32, p.tile2d ? W * 4 : W);   // 8x4 tiles: W/8 groups per tile row
        launch_fast_grad(p, (unsigned)blocks, stream);
    }
    else
    {
        size_t smem = (size_t)TEX_THREADS * C * sizeof(float);
        p.staged = smem <= 96 * 1024;
        if (!p.staged) smem = 0;
        if (smem > 48 * 1024) B200DR_CUDA(cudaFuncSetAttribute(texture_grad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024), "texture()");
        int grid = p.chunks;
        int cap = b200dr_num_sms() * 16;
        if (grid > cap) grid = cap;
        texture_grad_kernel<<<grid, TEX_THREADS, smem, stream>>>(p);
    }
    B200DR_CHECK_LAUNCH("texture_grad()");

    // resolve into the caller's layout
    r.C = C; r.Cp = p.Cp; r.texD = depth; r.mip_max = p.mip_max; r.baseW = texW; r.baseH = texH;
    int nlev = (pull_mips || p.mip_max == 0) ? 1 : p.mip_max + 1;
    for (int L = 0; L < nlev; L++)
    {
        r.level = L;
        r.texW = (texW >> L) > 1 ? (texW >> L) : 1; r.texH = (texH >> L) > 1 ? (texH >> L) : 1;
        r.pull = (L == 0 && pull_mips && p.mip_max > 0) ? 1 : 0;
        r.g_out = (L == 0) ? g_tex : (g_mip_ptrs ? g_mip_ptrs[L] : 0);
        if (!r.g_out) continue;
        size_t total = (size_t)r.texW * r.texH * depth * C;
        size_t blocks = (total + 255) / 256;
        size_t bcap = (size_t)b200dr_num_sms() * 16;
        if (blocks > bcap) blocks = bcap;
        texture_grad_resolve_kernel<<<(unsigned)blocks, 256, 0, stream>>>(r);
        b200dr_count_launches(1);
    }
    b200dr_count_launches(-1);              // the check below counts one
    B200DR_CHECK_LAUNCH("texture_grad(resolve)");
    return 0;
}

template <int F, int B, bool RASTER, bool STRIPS>
static void launch_chain_grad_c(const TexParams& p, const ChainParams& q, unsigned blocks, cudaStream_t stream)
{
    switch (p.C) {
    case 1: texture_chain_grad_fast<1, F, B, RASTER, STRIPS><<<blocks, TEXG_THREADS, 0, stream>>>(p, q); break;
    case 2: texture_chain_grad_fast<2, F, B, RASTER, STRIPS><<<blocks, TEXG_THREADS, 0, stream>>>(p, q); break;
    case 3: texture_chain_grad_fast<3, F, B, RASTER, STRIPS><<<blocks, TEXG_THREADS, 0, stream>>>(p, q); break;
    case 4: texture_chain_grad_fast<4, F, B, RASTER, STRIPS><<<blocks, TEXG_THREADS, 0, stream>>>(p, q); break;
    }
}
template <bool RASTER, bool STRIPS>
static void launch_chain_grad(const TexParams& p, const ChainParams& q, unsigned blocks, cudaStream_t stream)
{
    const int f = p.filter, b = p.boundary;
    if      (f == B200DR_FILTER_LINEAR_MIPMAP_LINEAR && b == B200DR_BOUNDARY_WRAP)  launch_chain_grad_c<B200DR_FILTER_LINEAR_MIPMAP_LINEAR, B200DR_BOUNDARY_WRAP, RASTER, STRIPS>(p, q, blocks, stream);
    else if (f == B200DR_FILTER_LINEAR_MIPMAP_LINEAR && b == B200DR_BOUNDARY_CLAMP) launch_chain_grad_c<B200DR_FILTER_LINEAR_MIPMAP_LINEAR, B200DR_BOUNDARY_CLAMP, RASTER, STRIPS>(p, q, blocks, stream);
    else                                                                            launch_chain_grad_c<-1, -1, RASTER, STRIPS>(p, q, blocks, stream);
}

// Fused backward of texture(interpolate(attr, rasterize(pos))): see texture_chain_grad_fast.  Extension of the op surface (the
// reference has no counterpart); the front end uses it only when it can prove the provenance of uv / uv_da, see
// nvdiffrast_b200/torch/ops.py (_fusable) and INTEGRATION.md.
extern "C" int b200dr_texture_chain_grad(const float* tex, const float* const* mip_ptrs, int max_level, const float* dy,
                                         const float* attr, const int32_t* attr_tri, int Va, int attr_per_image,
                                         const float* rast, const float* rast_db,
                                         const float* pos, const int32_t* pos_tri, int V, int grad_db,
                                         int N, int H, int W, int T, int texD, int texH, int texW, int C,
                                         int filter_mode, int boundary_mode,
                                         float* g_tex, float* g_attr, float* g_rast, float* g_rast_db, float* g_pos,
                                         void* scratch, size_t scratch_bytes, void* stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    const char* fn = "texture_chain_grad()";
    B200DR_CHECK(boundary_mode != B200DR_BOUNDARY_CUBE && C >= 1 && C <= 4, "%s: 2-D textures with at most 4 channels only", fn);
    B200DR_CHECK(filter_mode == B200DR_FILTER_LINEAR || filter_mode == B200DR_FILTER_LINEAR_MIPMAP_LINEAR || filter_mode == B200DR_FILTER_LINEAR_MIPMAP_NEAREST,
                 "%s: filter_mode unsupported", fn);
    const bool mip = (filter_mode != B200DR_FILTER_LINEAR);
    B200DR_CHECK(!mip || rast_db, "%s: mipmapped filtering needs rast_db (pixel differentials of the attributes)", fn);
    B200DR_CHECK(attr && attr_tri && rast && dy && Va > 0 && T > 0, "%s: missing input", fn);
    B200DR_CHECK((size_t)H * W < (1u << 30), "%s: image too large", fn);
    B200DR_CHECK(B200DR_ALIGNED16(rast) && B200DR_ALIGNED16(rast_db) && B200DR_ALIGNED16(dy) && (((uintptr_t)attr) & 7) == 0, "%s: input tensor misaligned", fn);
    B200DR_CHECK(!pos || (pos_tri && V > 0 && B200DR_ALIGNED16(pos)), "%s: pos / tri missing or misaligned", fn);
    B200DR_CHECK(B200DR_ALIGNED16(g_rast) && B200DR_ALIGNED16(g_rast_db) && (!g_rast || g_rast_db || !rast_db), "%s: g_rast outputs missing or misaligned", fn);
    B200DR_CHECK((((uintptr_t)g_attr) & 7) == 0 && B200DR_ALIGNED16(g_pos), "%s: gradient output misaligned", fn);
    TexParams p = {};
    // uv is recomputed in the kernel; a non-null placeholder tells the mode helpers that pixel differentials exist
    if (tex_fill(p, fn, tex, mip_ptrs, max_level, 0, mip ? rast_db : 0, 0, N, H, W, texD, texH, texW, C, filter_mode, boundary_mode)) return 1;
    p.dy = dy; p.tile2d = 0;
    size_t need = b200dr_texture_grad_scratch_bytes(texD, texH, texW, C, p.mip_max);
    B200DR_CHECK(scratch && scratch_bytes >= need && B200DR_ALIGNED16(scratch), "%s: scratch buffer too small or misaligned", fn);
    if (cudaMemsetAsync(scratch, 0, need, stream) != cudaSuccess) { b200dr_set_error("%s: cudaMemsetAsync failed", fn); return 1; }
    ResolveParams r = {};
    {
        char* s = (char*)scratch;
        for (int L = 0; L <= p.mip_max; L++)
        {
            size_t w = (texW >> L) > 1 ? (texW >> L) : 1, h = (texH >> L) > 1 ? (texH >> L) : 1;
            p.gpad[L] = (float*)s; r.gpad[L] = (const float*)s;
            s += w * h * texD * p.Cp * sizeof(float);
        }
    }
    ChainParams q = {};
    q.rast = rast; q.rast_db = rast_db; q.attr = attr; q.attr_tri = attr_tri; q.pos = pos; q.pos_tri = pos_tri;
    q.g_attr = g_attr; q.g_pos = g_pos; q.g_rast = g_rast; q.g_rast_db = g_rast_db;
    q.Va = Va; q.V = V; q.T = T; q.attr_per_image = attr_per_image; q.grad_db = grad_db;
    q.xs = 2.f / (float)W; q.xo = 1.f / (float)W - 1.f; q.ys = 2.f / (float)H; q.yo = 1.f / (float)H - 1.f;

    size_t groups = ((size_t)N * H * W + 31) / 32;
    size_t blocks = (groups + (TEXG_THREADS / 32) - 1) / (TEXG_THREADS / 32);
    size_t bcap = (size_t)b200dr_num_sms() * TCG_MINB * 4;
    q.strip_counter = (unsigned*)((char*)scratch + need - TEXG_TAIL_BYTES);
    const bool strips = TCG_STRIPS && (W % 32) == 0 && (H % TCG_STRIP_ROWS) == 0 && groups < 0xffffffffu;
    if (strips)
    {
        size_t nstrips = groups / TCG_STRIP_ROWS;
        blocks = (nstrips + (TEXG_THREADS / 32) - 1) / (TEXG_THREADS / 32);
        bcap = (size_t)b200dr_num_sms() * TCG_MINB * TCG_STRIP_WAVES;   // persistent: the work counter balances, no stride to choose
        if (blocks > bcap) blocks = bcap;
#if TCG_STRIPS
        if (pos) launch_chain_grad<true, true>(p, q, (unsigned)blocks, stream);
        else     launch_chain_grad<false, true>(p, q, (unsigned)blocks, stream);
#endif
    }
    else
    {
        if (blocks > bcap) blocks = b200dr_balanced_blocks(bcap, TEXG_THREADS / 32, W);
        if (pos) launch_chain_grad<true, false>(p, q, (unsigned)blocks, stream);
        else     launch_chain_grad<false, false>(p, q, (unsigned)blocks, stream);
    }
    B200DR_CHECK_LAUNCH(fn);

    r.C = C; r.Cp = p.Cp; r.texD = texD; r.mip_max = p.mip_max; r.baseW = texW; r.baseH = texH;
    r.level = 0; r.texW = texW; r.texH = texH; r.pull = (p.mip_max > 0) ? 1 : 0; r.g_out = g_tex;
    if (g_tex)
    {
        size_t total = (size_t)texW * texH * texD * C;
        size_t rb = (total + 255) / 256, rcap = (size_t)b200dr_num_sms() * 16;
        if (rb > rcap) rb = rcap;
        texture_grad_resolve_kernel<<<(unsigned)rb, 256, 0, stream>>>(r);
        B200DR_CHECK_LAUNCH("texture_chain_grad(resolve)");
    }
    return 0;
}

#ifdef TEX_COUNT_HITS
extern "C" int b200dr_debug_tex_counters(unsigned long long* out8, int reset)
{
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(out8, g_tex_counters, sizeof(unsigned long long) * 8);
    if (reset) { unsigned long long z[8] = {0}; cudaMemcpyToSymbol(g_tex_counters, z, sizeof(z)); }
    return 0;
}
#endif
