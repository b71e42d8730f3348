// texture.cu -- B200-native replacement of the reference's texture op, 2-D modes:
//   texture_construct_mip  (csrc/torch/torch_texture.cpp:98-169,  csrc/common/texture_kernel.cu:644-704)
//   texture_fwd[_mip]      (csrc/torch/torch_texture.cpp:174-416, csrc/common/texture_kernel.cu:709-800)
//   texture_grad_*         (csrc/torch/torch_texture.cpp:421-716, csrc/common/texture_kernel.cu:843-1154)
//   mip layout             (csrc/common/texture.cpp:62-102)
//
// One kernel per direction handles every 2-D filter/boundary mode (mode switches are warp-uniform), instead of the
// reference's 48-entry kernel table.  Per-pixel tensors are staged through shared memory for full-line accesses.
// Backward: texel gradients are accumulated into a channel-padded gradient pyramid with 128-bit vector reductions
// (REDG.E.ADD.F32x4: one atomic per texel instead of one per texel channel), then a resolve pass pulls the mip levels
// down to the base level and writes the gradient in the caller's channel layout.
#include "common.cuh"

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
};

static __device__ __forceinline__ int lvl_w(const TexParams& p, int l) { int v = p.texW >> l; return v > 1 ? v : 1; }
static __device__ __forceinline__ int lvl_h(const TexParams& p, int l) { int v = p.texH >> l; return v > 1 ? v : 1; }

// texture_kernel.cu:321-366 (2-D)
static __device__ __forceinline__ int index_nearest(const TexParams& p, float u, float v, int tz)
{
    int w = p.texW, h = p.texH;
    if (p.boundary == B200DR_BOUNDARY_WRAP) { u = u - (float)__float2int_rd(u); v = v - (float)__float2int_rd(v); }
    u = u * (float)w; v = v * (float)h;
    int iu = __float2int_rd(u), iv = __float2int_rd(v);
    if (p.boundary == B200DR_BOUNDARY_ZERO && (iu < 0 || iu >= w || iv < 0 || iv >= h)) return -1;
    iu = min(max(iu, 0), w - 1); iv = min(max(iv, 0), h - 1);
    return iu + w * (iv + tz * h);
}

// texture_kernel.cu:368-472 (2-D): four texel indices (-1 = outside in zero mode) and the bilinear weights.
static __device__ __forceinline__ float2 index_linear(const TexParams& p, float u, float v, int tz, int level, int4& tc)
{
    int w = lvl_w(p, level), h = lvl_h(p, level);
    bool clampU = false, clampV = false;
    if (p.boundary == B200DR_BOUNDARY_WRAP) { u = u - (float)__float2int_rd(u); v = v - (float)__float2int_rd(v); }
    u = u * (float)w - 0.5f; v = v * (float)h - 0.5f;
    if (p.boundary == B200DR_BOUNDARY_CLAMP)
    {
        u = fminf(fmaxf(u, 0.f), w - 1.f); v = fminf(fmaxf(v, 0.f), h - 1.f);
        clampU = (u == 0.f || u == w - 1.f); clampV = (v == 0.f || v == h - 1.f);
    }
    int iu0 = __float2int_rd(u), iv0 = __float2int_rd(v);
    int iu1 = iu0 + (clampU ? 0 : 1), iv1 = iv0 + (clampV ? 0 : 1);
    u -= (float)iu0; v -= (float)iv0;
    if (p.boundary == B200DR_BOUNDARY_WRAP)
    {
        if (iu0 < 0) iu0 += w;
        if (iv0 < 0) iv0 += h;
        if (iu1 >= w) iu1 -= w;
        if (iv1 >= h) iv1 -= h;
    }
    int iu0z = iu0 + tz * w * h, iu1z = iu1 + tz * w * h;
    tc.x = iu0z + w * iv0; tc.y = iu1z + w * iv0; tc.z = iu0z + w * iv1; tc.w = iu1z + w * iv1;
    if (p.boundary == B200DR_BOUNDARY_ZERO)
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
static __device__ __forceinline__ void mip_level(const TexParams& p, size_t pidx, int& level0, int& level1, float& flevel, float4& dw)
{
    level0 = 0; level1 = 0; flevel = 0.f;
    if (p.filter == B200DR_FILTER_NEAREST || p.filter == B200DR_FILTER_LINEAR) return;
    float fl = 0.f;
    if (p.uv_da)
    {
        float4 q = ld_stream_f4((const float4*)p.uv_da + pidx);
        float uscl = (float)p.texW, vscl = (float)p.texH;
        float dsdx = q.x * uscl, dsdy = q.y * uscl, dtdx = q.z * vscl, dtdy = q.w * vscl;
        float A = dsdx * dsdx + dtdx * dtdx;
        float B = dsdy * dsdy + dtdy * dtdy;
        float Cc = dsdx * dsdy + dtdx * dtdy;
        float l2b = 0.5f * (A + B);
        float l2n = 0.25f * (A - B) * (A - B) + Cc * Cc;
        float l2a = sqrtf(l2n);
        float lenMajorSqr = l2b + l2a;
        if (WANT_DW && p.filter == B200DR_FILTER_LINEAR_MIPMAP_LINEAR)
        {
            float d = 0.72134752f / (l2n + l2a * l2b);      // 0.5 / ln 2
            float AB = d * .5f * (A - B), Cw = d * Cc, l2aw = d * l2a;
            float4 r = make_float4(uscl * (dsdx * (l2aw + AB) + dsdy * Cw), uscl * (dsdy * (l2aw - AB) + dsdx * Cw),
                                   vscl * (dtdx * (l2aw + AB) + dtdy * Cw), vscl * (dtdy * (l2aw - AB) + dtdx * Cw));
            bool fin = isfinite(r.x) && isfinite(r.y) && isfinite(r.z) && isfinite(r.w);
            dw = fin ? r : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        fl = .5f * __log2f(lenMajorSqr);
    }
    if (p.bias) fl += ld_stream_f(p.bias + pidx);
    fl = fminf(fmaxf(fl, 0.f), (float)p.mip_max);
    level0 = __float2int_rd(fl);
    if (p.filter == B200DR_FILTER_LINEAR_MIPMAP_LINEAR && fl > 0.f)
    {
        level1 = min(level0 + 1, p.mip_max);
        fl -= (float)level0;
    }
    flevel = fl;
}

static __device__ __forceinline__ float texel(const float* t, int tc, int C, int c) { return tc >= 0 ? __ldg(t + (size_t)tc * C + c) : 0.f; }
static __device__ __forceinline__ float lerpf(float a, float b, float c) { return a + c * (b - a); }
static __device__ __forceinline__ float bilerp(float a, float b, float c, float d, float2 e) { return lerpf(lerpf(a, b, e.x), lerpf(c, d, e.x), e.y); }

// ------------------------------------------------------------------------------------------------
// Mip construction (texture_kernel.cu:644-704): level L from level L-1.

__global__ void __launch_bounds__(256) mip_build_kernel(const float* __restrict__ in, float* __restrict__ out, int wi, int hi, int wo, int ho, int D, int C)
{
    size_t total = (size_t)wo * ho * D * C;
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (size_t)gridDim.x * 256)
    {
        int c = (int)(i % C); size_t t = i / C;
        int x = (int)(t % wo); t /= wo;
        int y = (int)(t % ho); int z = (int)(t / ho);
        size_t i0 = (size_t)C * ((((size_t)x + (size_t)wi * y) << 1) + (size_t)z * wi * hi) + c;
        float v;
        if (wi == 1 || hi == 1)
        {
            size_t i1 = (hi == 1) ? i0 + C : i0 + (size_t)C * wi;
            v = .5f * (in[i0] + in[i1]);
        }
        else
        {
            size_t i1 = i0 + (size_t)C * wi;
            v = .25f * (in[i0] + in[i0 + C] + in[i1] + in[i1 + C]);
        }
        out[i] = v;
    }
}

// ------------------------------------------------------------------------------------------------
// Forward

__global__ void __launch_bounds__(TEX_THREADS) texture_fwd_kernel(const TexParams p)
{
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
            if (p.filter == B200DR_FILTER_NEAREST)
            {
                int tc = index_nearest(p, uv.x, uv.y, tz);
                for (int i = 0; i < C; i++) o[i] = texel(p.tex[0], tc, C, i);
            }
            else
            {
                int level0, level1; float flevel; float4 dw;
                mip_level<false>(p, pidx, level0, level1, flevel, dw);
                int4 tc0;
                float2 uv0 = index_linear(p, uv.x, uv.y, tz, level0, tc0);
                const float* t0 = p.tex[level0];
                if (p.filter != B200DR_FILTER_LINEAR_MIPMAP_LINEAR || !(flevel > 0.f))
                {
                    for (int i = 0; i < C; i++)
                        o[i] = bilerp(texel(t0, tc0.x, C, i), texel(t0, tc0.y, C, i), texel(t0, tc0.z, C, i), texel(t0, tc0.w, C, i), uv0);
                }
                else
                {
                    int4 tc1;
                    float2 uv1 = index_linear(p, uv.x, uv.y, tz, level1, tc1);
                    const float* t1 = p.tex[level1];
                    for (int i = 0; i < C; i++)
                    {
                        float a = bilerp(texel(t0, tc0.x, C, i), texel(t0, tc0.y, C, i), texel(t0, tc0.z, C, i), texel(t0, tc0.w, C, i), uv0);
                        float b = bilerp(texel(t1, tc1.x, C, i), texel(t1, tc1.y, C, i), texel(t1, tc1.z, C, i), texel(t1, tc1.w, C, i), uv1);
                        o[i] = lerpf(a, b, flevel);
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
// Lanes that hit the same texel of the same level in consecutive pixels (magnified textures, constant uv such as the
// background) are summed with shuffles first, so a texel receives one 128-bit vector reduction (REDG.E.ADD.F32x4) per
// run instead of one atomic per pixel and channel; 32 pixels on one texel cost one reduction, not 32 serialised ones.
static __device__ __forceinline__ void accum_texel_warp(float* g, int level, int tc, int C, int Cp, float w, const float* dy)
{
    int key = (tc >= 0) ? (tc ^ (level << 27)) : -1;
    WarpRuns runs = warp_runs(key);
    const bool dup = (runs.heads != 0xffffffffu);       // warp-uniform: any two neighbouring lanes on the same texel?
    float* q = g + (size_t)(tc >= 0 ? tc : 0) * Cp;
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
        if (dup)
        {
            v.x = warp_run_sum(v.x, runs);
            if (k + 1 < C) v.y = warp_run_sum(v.y, runs);
            if (k + 2 < C) v.z = warp_run_sum(v.z, runs);
            if (k + 3 < C) v.w = warp_run_sum(v.w, runs);
        }
        if (runs.is_head && tc >= 0) red_add4(q + k, v);
    }
}

__global__ void __launch_bounds__(TEX_THREADS) texture_grad_kernel(const TexParams p)
{
    extern __shared__ float s_dy[];
    const int tid = threadIdx.x;
    const int C = p.C, Cp = p.Cp;
    const size_t hw = (size_t)p.H * p.W;
    const size_t total = hw * p.N;
    const bool trilinear = (p.filter == B200DR_FILTER_LINEAR_MIPMAP_LINEAR);

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
            if (p.filter == B200DR_FILTER_NEAREST)
            {
                tc0.x = index_nearest(p, uv.x, uv.y, tz);
            }
            else
            {
                mip_level<true>(p, pidx, level0, level1, flevel, dw);
                uv0 = index_linear(p, uv.x, uv.y, tz, level0, tc0);
                two = trilinear && (flevel > 0.f);
                f0 = trilinear ? (1.f - flevel) : 1.f;
                if (two) uv1 = index_linear(p, uv.x, uv.y, tz, level1, tc1);
            }
        }

        // ---- texel gradients (warp-uniform control flow)
        if (__any_sync(0xffffffffu, act))
        {
            if (p.filter == B200DR_FILTER_NEAREST)
            {
                accum_texel_warp(p.gpad[0], 0, tc0.x, C, Cp, 1.f, pDy);
            }
            else
            {
                float uv011 = uv0.x * uv0.y, uv010 = uv0.x - uv011, uv001 = uv0.y - uv011, uv000 = 1.f - uv0.x - uv001;
                float* g0 = p.gpad[level0];
                accum_texel_warp(g0, level0, tc0.x, C, Cp, uv000 * f0, pDy);
                accum_texel_warp(g0, level0, tc0.y, C, Cp, uv010 * f0, pDy);
                accum_texel_warp(g0, level0, tc0.z, C, Cp, uv001 * f0, pDy);
                accum_texel_warp(g0, level0, tc0.w, C, Cp, uv011 * f0, pDy);
                if (__any_sync(0xffffffffu, two))
                {
                    float uv111 = uv1.x * uv1.y, uv110 = uv1.x - uv111, uv101 = uv1.y - uv111, uv100 = 1.f - uv1.x - uv101;
                    float* g1 = p.gpad[level1];
                    accum_texel_warp(g1, level1, two ? tc1.x : -1, C, Cp, uv100 * flevel, pDy);
                    accum_texel_warp(g1, level1, two ? tc1.y : -1, C, Cp, uv110 * flevel, pDy);
                    accum_texel_warp(g1, level1, two ? tc1.z : -1, C, Cp, uv101 * flevel, pDy);
                    accum_texel_warp(g1, level1, two ? tc1.w : -1, C, Cp, uv111 * flevel, pDy);
                }
            }
        }

        // ---- uv and mip-level gradients
        if (inb)
        {
            float gu = 0.f, gv = 0.f, df = 0.f;
            if (act && p.filter != B200DR_FILTER_NEAREST)
            {
                const float* t0 = p.tex[level0];
                const float* t1 = p.tex[level1];
                float sclu0 = (float)lvl_w(p, level0), sclv0 = (float)lvl_h(p, level0);
                float sclu1 = (float)lvl_w(p, level1), sclv1 = (float)lvl_h(p, level1);
                for (int i = 0; i < C; i++)
                {
                    float dy = pDy[i];
                    float dy0 = f0 * dy;
                    float a00 = texel(t0, tc0.x, C, i), a10 = texel(t0, tc0.y, C, i), a01 = texel(t0, tc0.z, C, i), a11 = texel(t0, tc0.w, C, i);
                    float ad = (a11 + a00 - a10 - a01);
                    gu += dy0 * ((a10 - a00) + uv0.y * ad) * sclu0;
                    gv += dy0 * ((a01 - a00) + uv0.x * ad) * sclv0;
                    if (two)
                    {
                        float dy1 = flevel * dy;
                        float b00 = texel(t1, tc1.x, C, i), b10 = texel(t1, tc1.y, C, i), b01 = texel(t1, tc1.z, C, i), b11 = texel(t1, tc1.w, C, i);
                        float bd = (b11 + b00 - b10 - b01);
                        gu += dy1 * ((b10 - b00) + uv1.y * bd) * sclu1;
                        gv += dy1 * ((b01 - b00) + uv1.x * bd) * sclv1;
                        float a = bilerp(a00, a10, a01, a11, uv0);
                        float b = bilerp(b00, b10, b01, b11, uv1);
                        df += (b - a) * dy;
                    }
                }
            }
            if (p.g_uv) st_stream_f2((float2*)p.g_uv + pidx, make_float2(gu, gv));
            if (p.g_bias) st_stream_f(p.g_bias + pidx, df);
            if (p.g_uv_da) st_stream_f4((float4*)p.g_uv_da + pidx, make_float4(dw.x * df, dw.y * df, dw.z * df, dw.w * df));
        }
        if (p.staged) __syncthreads();
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
    B200DR_CHECK(boundary != B200DR_BOUNDARY_CUBE, "%s: cube map mode is not implemented in this build", fn);
    B200DR_CHECK(N > 0 && H > 0 && W > 0, "%s: uv must have shape [>0, >0, >0, 2]", fn);
    B200DR_CHECK(texD > 0 && texH > 0 && texW > 0 && C > 0, "%s: tex must have shape[>0, >0, >0, >0]", fn);
    B200DR_CHECK(texD == 1 || texD == N, "%s: minibatch size mismatch between inputs tex, uv", fn);
    B200DR_CHECK(texW <= (1 << TEX_MAX_MIP) && texH <= (1 << TEX_MAX_MIP), "%s: texture size too large", fn);
    B200DR_CHECK(max_level >= 0 && max_level <= TEX_MAX_MIP, "%s: invalid mip level count", fn);
    bool mip = (filter == B200DR_FILTER_LINEAR_MIPMAP_NEAREST || filter == B200DR_FILTER_LINEAR_MIPMAP_LINEAR);
    B200DR_CHECK(!mip || uv_da || bias, "%s: mipmapping filter mode requires uv_da and/or mip_level_bias input", fn);
    B200DR_CHECK((((uintptr_t)uv) & 7) == 0, "%s: uv input tensor not aligned to float2", fn);
    B200DR_CHECK(B200DR_ALIGNED16(uv_da), "%s: uv_da input tensor not aligned to float4", fn);
    p.tex[0] = tex;
    p.mip_max = mip ? max_level : 0;
    for (int i = 1; i <= p.mip_max; i++)
    {
        B200DR_CHECK(mip_ptrs && mip_ptrs[i], "%s: mipmapping filter mode requires mip wrapper or mip stack input", fn);
        p.tex[i] = mip_ptrs[i];
    }
    p.uv = uv; p.uv_da = mip ? uv_da : 0; p.bias = mip ? bias : 0;
    p.N = N; p.H = H; p.W = W; p.texD = texD; p.texH = texH; p.texW = texW; p.C = C; p.Cp = (C + 3) & ~3;
    p.filter = filter; p.boundary = boundary;
    size_t total = (size_t)N * H * W;
    B200DR_CHECK(total < (1ull << 31) * 128, "%s: too many pixels", fn);
    p.chunks = (int)((total + TEX_THREADS - 1) / TEX_THREADS);
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
    B200DR_CHECK(max_level >= 0 && max_level <= TEX_MAX_MIP, "texture_construct_mip(): invalid max_mip_level");
    for (int L = 1; L <= max_level; L++)
    {
        const float* in = (L == 1) ? tex : mip + offsets[L - 1];
        float* out = mip + offsets[L];
        int wi = (texW >> (L - 1)) > 1 ? (texW >> (L - 1)) : 1, hi = (texH >> (L - 1)) > 1 ? (texH >> (L - 1)) : 1;
        int wo = (texW >> L) > 1 ? (texW >> L) : 1, ho = (texH >> L) > 1 ? (texH >> L) : 1;
        size_t total = (size_t)wo * ho * texD * C;
        size_t blocks = (total + 255) / 256;
        size_t cap = (size_t)b200dr_num_sms() * 16;
        if (blocks > cap) blocks = cap;
        mip_build_kernel<<<(unsigned)blocks, 256, 0, stream>>>(in, out, wi, hi, wo, ho, texD, C);
    }
    b200dr_count_launches(max_level - 1);   // max_level kernels, the check below counts one
    B200DR_CHECK_LAUNCH("texture_construct_mip()");
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
    size_t smem = (size_t)TEX_THREADS * C * sizeof(float);
    p.staged = smem <= 96 * 1024;
    if (!p.staged) smem = 0;
    if (smem > 48 * 1024) cudaFuncSetAttribute(texture_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    int grid = p.chunks;
    int cap = b200dr_num_sms() * 16;
    if (grid > cap) grid = cap;
    texture_fwd_kernel<<<grid, TEX_THREADS, smem, stream>>>(p);
    B200DR_CHECK_LAUNCH("texture_fwd()");
    return 0;
}

// Bytes of the channel-padded gradient pyramid used by b200dr_texture_grad (levels 0..max_level).
extern "C" size_t b200dr_texture_grad_scratch_bytes(int texD, int texH, int texW, int C, int max_level)
{
    size_t Cp = (size_t)((C + 3) & ~3), total = 0;
    for (int L = 0; L <= max_level; L++)
    {
        size_t w = (texW >> L) > 1 ? (texW >> L) : 1, h = (texH >> L) > 1 ? (texH >> L) : 1;
        total += w * h * texD * Cp * sizeof(float);
    }
    return total;
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
    B200DR_CHECK((((uintptr_t)g_uv) & 7) == 0, "texture_grad(): grad_uv output tensor not aligned to float2");
    B200DR_CHECK(B200DR_ALIGNED16(g_uv_da), "texture_grad(): grad_uv_da output tensor not aligned to float4");
    size_t need = b200dr_texture_grad_scratch_bytes(texD, texH, texW, C, p.mip_max);
    B200DR_CHECK(scratch && scratch_bytes >= need && B200DR_ALIGNED16(scratch), "texture_grad(): scratch buffer too small or misaligned");
    p.dy = dy; p.g_uv = (filter_mode != B200DR_FILTER_NEAREST) ? g_uv : 0;
    p.g_uv_da = (filter_mode == B200DR_FILTER_LINEAR_MIPMAP_LINEAR && uv_da) ? g_uv_da : 0;
    p.g_bias = (filter_mode == B200DR_FILTER_LINEAR_MIPMAP_LINEAR && mip_level_bias) ? g_bias : 0;

    // padded gradient pyramid
    cudaMemsetAsync(scratch, 0, need, stream);
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

    size_t smem = (size_t)TEX_THREADS * C * sizeof(float);
    p.staged = smem <= 96 * 1024;
    if (!p.staged) smem = 0;
    if (smem > 48 * 1024) cudaFuncSetAttribute(texture_grad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    int grid = p.chunks;
    int cap = b200dr_num_sms() * 16;
    if (grid > cap) grid = cap;
    texture_grad_kernel<<<grid, TEX_THREADS, smem, stream>>>(p);
    B200DR_CHECK_LAUNCH("texture_grad()");

    // resolve into the caller's layout
    r.C = C; r.Cp = p.Cp; r.texD = texD; r.mip_max = p.mip_max; r.baseW = texW; r.baseH = texH;
    int nlev = (pull_mips || p.mip_max == 0) ? 1 : p.mip_max + 1;
    for (int L = 0; L < nlev; L++)
    {
        r.level = L;
        r.texW = (texW >> L) > 1 ? (texW >> L) : 1; r.texH = (texH >> L) > 1 ? (texH >> L) : 1;
        r.pull = (L == 0 && pull_mips && p.mip_max > 0) ? 1 : 0;
        r.g_out = (L == 0) ? g_tex : (g_mip_ptrs ? g_mip_ptrs[L] : 0);
        if (!r.g_out) continue;
        size_t total = (size_t)r.texW * r.texH * texD * C;
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
