// interpolate.cu -- B200-native replacement of the reference's interpolate op:
//   interpolate_fwd[_da]  (csrc/torch/torch_interpolate.cpp:42-132,  csrc/common/interpolate.cu:15-126)
//   interpolate_grad[_da] (csrc/torch/torch_interpolate.cpp:137-248, csrc/common/interpolate.cu:131-274)
//
// Layout: the image batch is one flat array of P = N*H*W pixels; a warp owns 32 consecutive pixels, a CTA step
// owns 256.  Per-pixel AoS tensors with an arbitrary channel count A ([.., A] floats) are staged through shared
// memory so that every global access is a full-line 128-bit transaction regardless of A (the reference issues
// 4-byte stores with stride 4A from 8-pixel rows).  Vertex-attribute gradients are summed per run of equal
// triangle ids inside the warp (5 shuffles), accumulated in a per-CTA shared-memory copy of the gradient tensor when
// it fits, and flushed with one atomic per touched element per CTA.
#include "common.cuh"
#include "pixel_math.cuh"

namespace {

#define IP_THREADS DR_THREADS
#define IP_MAX_DIFF 32      // csrc/common/interpolate.h:18

struct InterpParams
{
    const int32_t* tri; const float* attr; const float* rast; const float* rast_db; const float* dy; const float* dda;
    float* out; float* out_da; float* g_attr; float* g_rast; float* g_rast_db;
    int N, H, W, V, T, A, D;
    int attr_bc, instance_mode, diff_all;
    int staged;             // 1: per-pixel channel tensors go through shared-memory staging
    int use_smem_acc;       // backward: 1 = accumulate g_attr in shared memory
    int chunks;             // 256-pixel chunks per image
    // fused backward through rasterize (interpolate_grad_fast<.., RASTER = true>): g_rast / g_rast_db stay in registers and go
    // straight through the rasterizer's pixel math into g_pos
    const float* pos; const int32_t* pos_tri; float* g_pos;    // [N,Vp,4]; [T,3]; [N,Vp,4] zero-initialised
    int Vp, grad_db, use_smem_pos;
    float xs, xo, ys, yo;   // pixel centre -> clip space: fx = xs * X + xo
    int diff[IP_MAX_DIFF];
};

static __device__ __forceinline__ int diff_index(const InterpParams& p, int i)
{
    int j = p.diff_all ? i : p.diff[i];
    if (j < 0) j += p.A;
    return j;
}

// Decode the triangle of a pixel and fetch its (validated, instance-adjusted) vertex indices.
static __device__ __forceinline__ bool pixel_triangle(const InterpParams& p, float idf, int n, int& vi0, int& vi1, int& vi2)
{
    int triIdx = float_to_triidx(idf) - 1;
    if (triIdx < 0 || triIdx >= p.T) return false;
    vi0 = __ldg(p.tri + triIdx * 3 + 0); vi1 = __ldg(p.tri + triIdx * 3 + 1); vi2 = __ldg(p.tri + triIdx * 3 + 2);
    if (vi0 < 0 || vi0 >= p.V || vi1 < 0 || vi1 >= p.V || vi2 < 0 || vi2 >= p.V) return false;
    if (p.instance_mode && !p.attr_bc) { vi0 += n * p.V; vi1 += n * p.V; vi2 += n * p.V; }
    return true;
}

// ------------------------------------------------------------------------------------------------
// Forward

template <bool DA>
__global__ void __launch_bounds__(IP_THREADS) interpolate_fwd_kernel(const InterpParams p)
{
    extern __shared__ float s_mem[];
    float* s_out = s_mem;                                   // [256 * A]
    float* s_da = s_mem + IP_THREADS * p.A;                 // [256 * 2D]
    const int tid = threadIdx.x;
    const int n = blockIdx.y;
    const size_t hw = (size_t)p.H * p.W;
    const int A = p.A, D2 = 2 * p.D;

    for (int c = blockIdx.x; c < p.chunks; c += gridDim.x)
    {
        size_t r0 = (size_t)c * IP_THREADS;
        int npx = (int)min((size_t)IP_THREADS, hw - r0);
        size_t pbase = (size_t)n * hw + r0;
        bool inb = tid < npx;

        float* o = p.staged ? (s_out + tid * A) : (p.out + (pbase + tid) * A);
        float* oda = DA ? (p.staged ? (s_da + tid * D2) : (p.out_da + (pbase + tid) * D2)) : 0;

        if (inb)
        {
            float4 r = ld_stream_f4((const float4*)p.rast + pbase + tid);
            int vi0, vi1, vi2;
            if (pixel_triangle(p, r.w, n, vi0, vi1, vi2))
            {
                const float* a0 = p.attr + (size_t)vi0 * A;
                const float* a1 = p.attr + (size_t)vi1 * A;
                const float* a2 = p.attr + (size_t)vi2 * A;
                float b0 = r.x, b1 = r.y, b2 = 1.f - r.x - r.y;
                for (int i = 0; i < A; i++)
                    o[i] = b0 * __ldg(a0 + i) + b1 * __ldg(a1 + i) + b2 * __ldg(a2 + i);
                if (DA)
                {
                    float4 db = ld_stream_f4((const float4*)p.rast_db + pbase + tid);
                    for (int i = 0; i < p.D; i++)
                    {
                        int j = diff_index(p, i);
                        float dsdx = 0.f, dsdy = 0.f;
                        if (j >= 0 && j < A)
                        {
                            float s2 = __ldg(a2 + j);
                            float dsdu = __ldg(a0 + j) - s2;
                            float dsdv = __ldg(a1 + j) - s2;
                            dsdx = db.x * dsdu + db.z * dsdv;
                            dsdy = db.y * dsdu + db.w * dsdv;
                        }
                        oda[2 * i] = dsdx; oda[2 * i + 1] = dsdy;
                    }
                }
            }
            else
            {
                for (int i = 0; i < A; i++) o[i] = 0.f;
                if (DA) for (int i = 0; i < D2; i++) oda[i] = 0.f;
            }
        }

        if (p.staged)
        {
            __syncthreads();
            coop_store(p.out + pbase * A, s_out, npx * A, tid);
            if (DA) coop_store(p.out_da + pbase * D2, s_da, npx * D2, tid);
            __syncthreads();
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Backward

template <bool DA>
__global__ void __launch_bounds__(IP_THREADS) interpolate_grad_kernel(const InterpParams p)
{
    extern __shared__ float s_mem[];
    const int tid = threadIdx.x;
    const int n = blockIdx.y;
    const size_t hw = (size_t)p.H * p.W;
    const int A = p.A, D2 = 2 * p.D;
    const bool per_image = p.instance_mode && !p.attr_bc;
    const int acc_n = p.V * A;

    float* gbase = p.g_attr + (per_image ? (size_t)n * acc_n : 0);
    float* s_acc = s_mem;                                           // [V*A] when use_smem_acc
    float* s_dy = s_mem + (p.use_smem_acc ? acc_n : 0);             // [256*A]
    float* s_dda = s_dy + (p.staged ? IP_THREADS * A : 0);          // [256*2D]
    float* acc = p.use_smem_acc ? s_acc : p.g_attr;                 // global mode: vertex ids already carry the image offset
    const int voff = (p.use_smem_acc && per_image) ? n * p.V : 0;   // shared copy is indexed by image-local vertex id

    if (p.use_smem_acc)
    {
        for (int i = tid; i < acc_n; i += IP_THREADS) s_acc[i] = 0.f;
        __syncthreads();
    }

    for (int c = blockIdx.x; c < p.chunks; c += gridDim.x)
    {
        size_t r0 = (size_t)c * IP_THREADS;
        int npx = (int)min((size_t)IP_THREADS, hw - r0);
        size_t pbase = (size_t)n * hw + r0;
        bool inb = tid < npx;

        if (p.staged)
        {
            coop_load(s_dy, p.dy + pbase * A, npx * A, tid);
            if (DA) coop_load(s_dda, p.dda + pbase * D2, npx * D2, tid);
            __syncthreads();
        }
        const float* pdy = p.staged ? (s_dy + tid * A) : (p.dy + (pbase + tid) * A);
        const float* pdda = DA ? (p.staged ? (s_dda + tid * D2) : (p.dda + (pbase + tid) * D2)) : 0;

        float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
        int vi0 = 0, vi1 = 0, vi2 = 0;
        bool valid = false;
        int key = -1;
        if (inb)
        {
            r = ld_stream_f4((const float4*)p.rast + pbase + tid);
            valid = pixel_triangle(p, r.w, n, vi0, vi1, vi2);
            if (valid) key = float_to_triidx(r.w);
        }

        if (__any_sync(0xffffffffu, valid))
        {
            WarpRuns runs = warp_runs(key);
            bool head = runs.is_head && valid;
            const float* a0 = p.attr + (size_t)vi0 * A;     // vi* are 0 for invalid lanes: loads stay in bounds
            const float* a1 = p.attr + (size_t)vi1 * A;
            const float* a2 = p.attr + (size_t)vi2 * A;
            float* ga0 = acc + (size_t)(vi0 - voff) * A;
            float* ga1 = acc + (size_t)(vi1 - voff) * A;
            float* ga2 = acc + (size_t)(vi2 - voff) * A;
            float b0 = r.x, b1 = r.y, b2 = 1.f - r.x - r.y;
            float gb0 = 0.f, gb1 = 0.f;
            float4 db = make_float4(0.f, 0.f, 0.f, 0.f);
            if (DA && valid) db = ld_stream_f4((const float4*)p.rast_db + pbase + tid);
            const bool merged = DA && p.diff_all;           // differentials of attribute i ride along with attribute i
            float gdudx = 0.f, gdudy = 0.f, gdvdx = 0.f, gdvdy = 0.f;

            for (int i = 0; i < A; i++)
            {
                float y = valid ? pdy[i] : 0.f;
                float s0 = __ldg(a0 + i), s1 = __ldg(a1 + i), s2 = __ldg(a2 + i);
                gb0 += y * (s0 - s2);
                gb1 += y * (s1 - s2);
                float t0 = b0 * y, t1 = b1 * y, t2 = b2 * y;
                if (merged)
                {
                    float dsdx = valid ? pdda[2 * i] : 0.f, dsdy = valid ? pdda[2 * i + 1] : 0.f;
                    float dsdu = s0 - s2, dsdv = s1 - s2;
                    gdudx += dsdu * dsdx; gdudy += dsdu * dsdy; gdvdx += dsdv * dsdx; gdvdy += dsdv * dsdy;
                    float du = dsdx * db.x + dsdy * db.y;
                    float dv = dsdx * db.z + dsdy * db.w;
                    t0 += du; t1 += dv; t2 += -du - dv;
                }
                t0 = warp_run_sum(t0, runs); t1 = warp_run_sum(t1, runs); t2 = warp_run_sum(t2, runs);
                if (head) { red_add(ga0 + i, t0); red_add(ga1 + i, t1); red_add(ga2 + i, t2); }
            }
            if (DA && !merged)
            {
                for (int i = 0; i < p.D; i++)
                {
                    int j = diff_index(p, i);
                    if (j < 0 || j >= A) continue;                  // uniform across the warp
                    float dsdx = valid ? pdda[2 * i] : 0.f, dsdy = valid ? pdda[2 * i + 1] : 0.f;
                    float s2 = __ldg(a2 + j);
                    float dsdu = __ldg(a0 + j) - s2, dsdv = __ldg(a1 + j) - s2;
                    gdudx += dsdu * dsdx; gdudy += dsdu * dsdy; gdvdx += dsdv * dsdx; gdvdy += dsdv * dsdy;
                    float du = dsdx * db.x + dsdy * db.y;
                    float dv = dsdx * db.z + dsdy * db.w;
                    float t0 = warp_run_sum(du, runs), t1 = warp_run_sum(dv, runs), t2 = warp_run_sum(-du - dv, runs);
                    if (head) { red_add(ga0 + j, t0); red_add(ga1 + j, t1); red_add(ga2 + j, t2); }
                }
            }
            if (inb)
            {
                st_stream_f4((float4*)p.g_rast + pbase + tid, valid ? make_float4(gb0, gb1, 0.f, 0.f) : make_float4(0.f, 0.f, 0.f, 0.f));
                if (DA) st_stream_f4((float4*)p.g_rast_db + pbase + tid, valid ? make_float4(gdudx, gdudy, gdvdx, gdvdy) : make_float4(0.f, 0.f, 0.f, 0.f));
            }
        }
        else if (inb)
        {
            st_stream_f4((float4*)p.g_rast + pbase + tid, make_float4(0.f, 0.f, 0.f, 0.f));
            if (DA) st_stream_f4((float4*)p.g_rast_db + pbase + tid, make_float4(0.f, 0.f, 0.f, 0.f));
        }
        if (p.staged) __syncthreads();      // staging buffers are reused by the next chunk
    }

    if (p.use_smem_acc)
    {
        __syncthreads();
        for (int i = tid; i < acc_n; i += IP_THREADS)
        {
            float v = s_acc[i];
            if (v != 0.f) red_add_global(gbase + i, v);
        }
    }
}

// ================================================================================================
// Fast paths: A in {1,2,3,4}, pixel differentials for none or all attributes.
//
// Warp-autonomous: a warp owns groups of 32 consecutive pixels, two groups in flight per iteration, no block barriers in
// the loop.  Channel count is a template parameter, so attribute rows live in registers and every per-pixel tensor is
// moved with one vector access per lane (K = 1, 2, 4, 8 floats) or through a warp-private shared-memory transpose
// (K = 3, 6).  Per-pixel tensors that only matter under geometry (rast_db, dy, dda) are fetched after the triangle id is
// known: empty regions cost 16 bytes per pixel instead of the full algorithmic traffic.
// ================================================================================================

#ifndef ACC_SMEM_MAX_V
#define ACC_SMEM_MAX_V (1 << 30)
#endif
// forward: groups of 32 pixels in flight per warp.  Without pixel differentials 4 (A = 3 at 2048^2 x 16: 0.346 -> 0.322 ms;
// 1: 0.370); with them 2 (c3, A = 2 + differentials: 4 groups cost 0.544 -> 0.624 ms in registers)
#ifndef IPF_UNROLL_PLAIN
#define IPF_UNROLL_PLAIN 4
#endif
#ifndef IPF_UNROLL_DA
#define IPF_UNROLL_DA 2
#endif
// backward (r02 call 13, same box): 4 groups in flight at <= 64 registers (4 CTAs/SM) against 2 groups unconstrained: A = 3 target
// 0.710 -> 0.677 ms, c2 0.509 -> 0.482, c3 op-by-op (A = 2 + differentials) 0.846 -> 0.808; 3 groups: 0.690 / 0.487 / 0.817.
// call 18, A = 3 target: 6 groups 0.859, 8 groups at 3 CTAs 1.168 (code size), 5 CTAs/SM 0.668, dy requested for all groups
// before the first group's gathers 0.699 -> none adopted
#ifndef IPG_UNROLL_PLAIN
#define IPG_UNROLL_PLAIN 4
#endif
#ifndef IPG_UNROLL_DA
#define IPG_UNROLL_DA 4
#endif
#ifndef IPG_MINB_PLAIN
#define IPG_MINB_PLAIN 4
#endif
#ifndef IPG_MINB_DA
#define IPG_MINB_DA 4
#endif

template <int A>
static __device__ __forceinline__ void load_row(const float* a, float (&v)[A])
{
    if (A == 2) { float2 t = __ldg((const float2*)a); v[0] = t.x; v[1] = t.y; }
    else if (A == 4) { float4 t = __ldg((const float4*)a); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
    else { _Pragma("unroll") for (int i = 0; i < A; i++) v[i] = __ldg(a + i); }
}

template <int A, bool DA>
__global__ void __launch_bounds__(IP_THREADS) interpolate_fwd_fast(const InterpParams p)
{
    constexpr int SK = (A == 3) ? (DA ? 6 : 3) : 1;                    // staging floats per pixel (only K = 3 / 6 need it)
    __shared__ __align__(16) float s_stage[IP_THREADS / 32][32 * SK];
    constexpr int IPF_UNROLL = DA ? IPF_UNROLL_DA : IPF_UNROLL_PLAIN;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n = blockIdx.y;
    const int hw = p.H * p.W;
    const int groups = (hw + 31) >> 5;
    const int wstride = gridDim.x * (IP_THREADS / 32);
    const size_t img = (size_t)n * hw;
    const bool row_aligned = (((size_t)hw * A) & 3) == 0;

    for (int g0 = blockIdx.x * (IP_THREADS / 32) + warp; g0 < groups; g0 += IPF_UNROLL * wstride)
    {
        float4 r[IPF_UNROLL], db[IPF_UNROLL];
        bool inb[IPF_UNROLL], valid[IPF_UNROLL];
        int vi0[IPF_UNROLL], vi1[IPF_UNROLL], vi2[IPF_UNROLL];
        #pragma unroll
        for (int u = 0; u < IPF_UNROLL; u++)
        {
            int px = (g0 + u * wstride) * 32 + lane;
            inb[u] = (g0 + u * wstride < groups) && (px < hw);
            r[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (inb[u]) r[u] = ld_stream_f4((const float4*)p.rast + img + px);
        }
        #pragma unroll
        for (int u = 0; u < IPF_UNROLL; u++)
        {
            int px = (g0 + u * wstride) * 32 + lane;
            vi0[u] = vi1[u] = vi2[u] = 0;
            valid[u] = inb[u] && pixel_triangle(p, r[u].w, n, vi0[u], vi1[u], vi2[u]);
            db[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (DA && valid[u]) db[u] = ld_stream_f4((const float4*)p.rast_db + img + px);
        }
        #pragma unroll
        for (int u = 0; u < IPF_UNROLL; u++)
        {
            int g = g0 + u * wstride;
            if (g >= groups) continue;                                  // warp-uniform
            float o[A], oda[DA ? 2 * A : 1];
            #pragma unroll
            for (int i = 0; i < A; i++) o[i] = 0.f;
            #pragma unroll
            for (int i = 0; i < (DA ? 2 * A : 1); i++) oda[i] = 0.f;
            if (valid[u])
            {
                float a0[A], a1[A], a2[A];
                load_row<A>(p.attr + (size_t)vi0[u] * A, a0);
                load_row<A>(p.attr + (size_t)vi1[u] * A, a1);
                load_row<A>(p.attr + (size_t)vi2[u] * A, a2);
                float b0 = r[u].x, b1 = r[u].y, b2 = 1.f - r[u].x - r[u].y;
                #pragma unroll
                for (int i = 0; i < A; i++)
                {
                    o[i] = b0 * a0[i] + b1 * a1[i] + b2 * a2[i];
                    if (DA)
                    {
                        float dsdu = a0[i] - a2[i], dsdv = a1[i] - a2[i];
                        oda[2 * i] = db[u].x * dsdu + db[u].z * dsdv;
                        oda[2 * i + 1] = db[u].y * dsdu + db[u].w * dsdv;
                    }
                }
            }
            size_t pbase = img + (size_t)g * 32;
            bool full = row_aligned && (g * 32 + 32 <= hw);
            warp_store_aos<A>(p.out + pbase * A, o, lane, inb[u], full, s_stage[warp]);
            if (DA)
            {
                float (&od)[DA ? 2 * A : 1] = oda;
                warp_store_aos<(DA ? 2 * A : 1)>(p.out_da + pbase * (2 * A), od, lane, inb[u], full, s_stage[warp]);
            }
        }
    }
}

// RASTER: the rasterizer's backward is applied in the same pass (b200dr_interpolate_chain_grad): per pixel, (g_rast.xy, g_rast_db)
// go through pm::raster_grad instead of memory, and the vertex-position gradients are summed over the same triangle runs.
// r02 call 14 (c2 / c4 fused kernel): 4 CTAs/SM at 64 registers 0.638 / 0.715 ms; 3 CTAs at 80 registers 0.810 / 0.767; 2 CTAs the same
// call 15: 2 groups in flight instead of 4 (half the code: ncu showed instruction-fetch stalls) 0.593 / 0.632; 5 CTAs/SM 0.607 / 0.749
#ifndef IPC_MINB
#define IPC_MINB 4
#endif
#ifndef IPC_UNROLL
#define IPC_UNROLL 2
#endif
#define IPC_SMEM_MAX_VP 1024    // position gradients of one image in shared memory up to this many vertices (as raster_grad_kernel)
template <int A, bool DA, bool RASTER>
__global__ void __launch_bounds__(IP_THREADS, RASTER ? IPC_MINB : (DA ? IPG_MINB_DA : IPG_MINB_PLAIN)) interpolate_grad_fast(const InterpParams p)
{
    constexpr int IPG_UNROLL = RASTER ? IPC_UNROLL : (DA ? IPG_UNROLL_DA : IPG_UNROLL_PLAIN);
    extern __shared__ __align__(16) float s_dyn[];
    constexpr int SK = (A == 3) ? (DA ? 6 : 3) : 1;
    float* s_stage = s_dyn;                                             // [8 warps][32 * SK]
    float* s_acc = s_dyn + (IP_THREADS / 32) * 32 * SK;                 // [V * A] when use_smem_acc
    float* s_pos = s_acc + (p.use_smem_acc ? p.V * A : 0);              // RASTER: [Vp * 4] when use_smem_pos (this image's g_pos)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = blockIdx.y;
    const int hw = p.H * p.W;
    const int groups = (hw + 31) >> 5;
    const int wstride = gridDim.x * (IP_THREADS / 32);
    const size_t img = (size_t)n * hw;
    const bool row_aligned = (((size_t)hw * A) & 3) == 0;
    const bool per_image = p.instance_mode && !p.attr_bc;
    const int acc_n = p.V * A;
    float* gbase = p.g_attr + (per_image ? (size_t)n * acc_n : 0);
    float* acc = p.use_smem_acc ? s_acc : p.g_attr;
    const int voff = (p.use_smem_acc && per_image) ? n * p.V : 0;
    float* sw = s_stage + warp * 32 * SK;

    if (p.use_smem_acc)
        for (int i = tid; i < acc_n; i += IP_THREADS) s_acc[i] = 0.f;
    if (RASTER && p.use_smem_pos)
        for (int i = tid; i < p.Vp * 4; i += IP_THREADS) s_pos[i] = 0.f;
    if (p.use_smem_acc || (RASTER && p.use_smem_pos)) __syncthreads();

    for (int g0 = blockIdx.x * (IP_THREADS / 32) + warp; g0 < groups; g0 += IPG_UNROLL * wstride)
    {
        float4 r[IPG_UNROLL];
        bool inb[IPG_UNROLL];
        #pragma unroll
        for (int u = 0; u < IPG_UNROLL; u++)
        {
            int px = (g0 + u * wstride) * 32 + lane;
            inb[u] = (g0 + u * wstride < groups) && (px < hw);
            r[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (inb[u]) r[u] = ld_stream_f4((const float4*)p.rast + img + px);
        }
        #pragma unroll
        for (int u = 0; u < IPG_UNROLL; u++)
        {
            int g = g0 + u * wstride;
            if (g >= groups) continue;                                  // warp-uniform
            int px = g * 32 + lane;
            size_t pbase = img + (size_t)g * 32;
            int vi0 = 0, vi1 = 0, vi2 = 0;
            bool valid = inb[u] && pixel_triangle(p, r[u].w, n, vi0, vi1, vi2);
            float4 gr = make_float4(0.f, 0.f, 0.f, 0.f), gdb = gr;
            if (__any_sync(0xffffffffu, valid))
            {
                bool full = row_aligned && (g * 32 + 32 <= hw);
                float dy[A], dda[DA ? 2 * A : 1];
                #pragma unroll
                for (int i = 0; i < A; i++) dy[i] = 0.f;
                #pragma unroll
                for (int i = 0; i < (DA ? 2 * A : 1); i++) dda[i] = 0.f;
                float4 db = make_float4(0.f, 0.f, 0.f, 0.f);
                // K = 3 / 6 go through the warp transpose (all lanes take part); vector widths are fetched only under geometry
                warp_load_aos<A>(p.dy + pbase * A, dy, lane, (SK == 1) ? valid : inb[u], full, sw);
                if (DA)
                {
                    warp_load_aos<(DA ? 2 * A : 1)>(p.dda + pbase * (2 * A), dda, lane, (SK == 1) ? valid : inb[u], full, sw);
                    if (valid) db = ld_stream_f4((const float4*)p.rast_db + img + px);
                }
                float a0[A], a1[A], a2[A];
                load_row<A>(p.attr + (size_t)vi0 * A, a0);             // vi* are 0 for lanes without geometry: in bounds
                load_row<A>(p.attr + (size_t)vi1 * A, a1);
                load_row<A>(p.attr + (size_t)vi2 * A, a2);
                WarpRuns runs = warp_runs(valid ? float_to_triidx(r[u].w) : -1);
                bool head = runs.is_head && valid;
                float* ga0 = acc + (size_t)(vi0 - voff) * A;
                float* ga1 = acc + (size_t)(vi1 - voff) * A;
                float* ga2 = acc + (size_t)(vi2 - voff) * A;
                float b0 = r[u].x, b1 = r[u].y, b2 = 1.f - r[u].x - r[u].y;
                float t0[A], t1[A], t2[A];
                #pragma unroll
                for (int i = 0; i < A; i++)
                {
                    float y = valid ? dy[i] : 0.f;
                    float dsdu = a0[i] - a2[i], dsdv = a1[i] - a2[i];
                    gr.x += y * dsdu;
                    gr.y += y * dsdv;
                    t0[i] = b0 * y; t1[i] = b1 * y; t2[i] = b2 * y;
                    if (DA)
                    {
                        float dsdx = valid ? dda[2 * i] : 0.f, dsdy = valid ? dda[2 * i + 1] : 0.f;
                        gdb.x += dsdu * dsdx; gdb.y += dsdu * dsdy; gdb.z += dsdv * dsdx; gdb.w += dsdv * dsdy;
                        float du = dsdx * db.x + dsdy * db.y;
                        float dv = dsdx * db.z + dsdy * db.w;
                        t0[i] += du; t1[i] += dv; t2[i] += -du - dv;
                    }
                    t0[i] = warp_run_sum(t0[i], runs); t1[i] = warp_run_sum(t1[i], runs); t2[i] = warp_run_sum(t2[i], runs);
                }
                if (head)
                {
                    // global accumulation of 2- and 4-channel attributes: one vector reduction per vertex (REDG.F32x2 / x4);
                    // the shared copy takes scalar adds (an fp32 add to shared memory is a compare-and-swap loop)
                    if (!p.use_smem_acc && A == 2)
                    {
                        red_add2(ga0, make_float2(t0[0], t0[1 % A])); red_add2(ga1, make_float2(t1[0], t1[1 % A])); red_add2(ga2, make_float2(t2[0], t2[1 % A]));
                    }
                    else if (!p.use_smem_acc && A == 4)
                    {
                        red_add4(ga0, make_float4(t0[0], t0[1 % A], t0[2 % A], t0[3 % A])); red_add4(ga1, make_float4(t1[0], t1[1 % A], t1[2 % A], t1[3 % A]));
                        red_add4(ga2, make_float4(t2[0], t2[1 % A], t2[2 % A], t2[3 % A]));
                    }
                    else if (!p.use_smem_acc && A == 3)
                    {
                        // 12-byte rows: the 8-byte aligned pair is (0,1) for even vertices and (1,2) for odd ones
                        red_add3(ga0, t0[0], t0[1 % A], t0[2 % A]); red_add3(ga1, t1[0], t1[1 % A], t1[2 % A]); red_add3(ga2, t2[0], t2[1 % A], t2[2 % A]);
                    }
                    else if (p.use_smem_acc)
                    {
                        #pragma unroll
                        for (int i = 0; i < A; i++) { red_add_shared(ga0 + i, t0[i]); red_add_shared(ga1 + i, t1[i]); red_add_shared(ga2 + i, t2[i]); }
                    }
                    else
                    {
                        #pragma unroll
                        for (int i = 0; i < A; i++) { red_add_global(ga0 + i, t0[i]); red_add_global(ga1 + i, t1[i]); red_add_global(ga2 + i, t2[i]); }
                    }
                }
                if (!valid) { gr = make_float4(0.f, 0.f, 0.f, 0.f); gdb = gr; }
                if (RASTER)
                {
                    // rasterize backward on the pixel's own (g_rast.xy, g_rast_db): csrc/common/rasterize.cu:119-273 as restated in
                    // pixel_math.cuh.  The runs are the attribute runs (same triangle id); position indices come from rasterize's tri.
                    int pi0 = 0, pi1 = 0, pi2 = 0;
                    if (valid)
                    {
                        const int32_t* pt = p.pos_tri + (size_t)(float_to_triidx(r[u].w) - 1) * 3;
                        pi0 = __ldg(pt); pi1 = __ldg(pt + 1); pi2 = __ldg(pt + 2);
                    }
                    const bool pidx_ok = valid && !(pi0 < 0 || pi0 >= p.Vp || pi1 < 0 || pi1 >= p.Vp || pi2 < 0 || pi2 >= p.Vp);
                    // pixels whose incoming gradients are all +-0 are skipped (csrc/common/rasterize.cu:143-148)
                    const unsigned nz_dy = (__float_as_uint(gr.x) | __float_as_uint(gr.y)) << 1;
                    const unsigned nz_db = (DA && p.grad_db) ? ((__float_as_uint(gdb.x) | __float_as_uint(gdb.y) | __float_as_uint(gdb.z) | __float_as_uint(gdb.w)) << 1) : 0u;
                    float gv9[9];
                    #pragma unroll
                    for (int i = 0; i < 9; i++) gv9[i] = 0.f;
                    if (pidx_ok && (nz_dy | nz_db) != 0u)
                    {
                        const float4* vb = (const float4*)p.pos + (size_t)n * p.Vp;
                        const unsigned py = (unsigned)px / (unsigned)p.W, pxx = (unsigned)px - py * (unsigned)p.W;
                        const float fx = p.xs * (float)pxx + p.xo, fy = p.ys * (float)py + p.yo;
                        pm::PixelTri t;
                        pm::pixel_tri(t, __ldg(vb + pi0), __ldg(vb + pi1), __ldg(vb + pi2), fx, fy);
                        float gx[3], gy[3], gw[3];
                        pm::raster_grad<DA>(t, fx, fy, p.xs, p.ys, gr.x, gr.y, gdb, nz_db != 0u, gx, gy, gw);
                        #pragma unroll
                        for (int k = 0; k < 3; k++) { gv9[3 * k] = gx[k]; gv9[3 * k + 1] = gy[k]; gv9[3 * k + 2] = gw[k]; }
                    }
                    #pragma unroll
                    for (int i = 0; i < 9; i++) gv9[i] = warp_run_sum(gv9[i], runs);
                    if (head && pidx_ok)
                    {
                        if (p.use_smem_pos)
                        {
                            red_add_shared(s_pos + 4 * pi0, gv9[0]); red_add_shared(s_pos + 4 * pi0 + 1, gv9[1]); red_add_shared(s_pos + 4 * pi0 + 3, gv9[2]);
                            red_add_shared(s_pos + 4 * pi1, gv9[3]); red_add_shared(s_pos + 4 * pi1 + 1, gv9[4]); red_add_shared(s_pos + 4 * pi1 + 3, gv9[5]);
                            red_add_shared(s_pos + 4 * pi2, gv9[6]); red_add_shared(s_pos + 4 * pi2 + 1, gv9[7]); red_add_shared(s_pos + 4 * pi2 + 3, gv9[8]);
                        }
                        else
                        {
                            float* gp = p.g_pos + (size_t)n * p.Vp * 4;
                            red_add4(gp + 4 * pi0, make_float4(gv9[0], gv9[1], 0.f, gv9[2]));
                            red_add4(gp + 4 * pi1, make_float4(gv9[3], gv9[4], 0.f, gv9[5]));
                            red_add4(gp + 4 * pi2, make_float4(gv9[6], gv9[7], 0.f, gv9[8]));
                        }
                    }
                }
            }
            if (!RASTER && inb[u])
            {
                st_stream_f4((float4*)p.g_rast + img + px, gr);
                if (DA) st_stream_f4((float4*)p.g_rast_db + img + px, gdb);
            }
        }
    }

    if (p.use_smem_acc || (RASTER && p.use_smem_pos)) __syncthreads();
    if (p.use_smem_acc)
    {
        for (int i = tid; i < acc_n; i += IP_THREADS)
        {
            float v = s_acc[i];
            if (v != 0.f) red_add_global(gbase + i, v);
        }
    }
    if (RASTER && p.use_smem_pos)
    {
        float* gp = p.g_pos + (size_t)n * p.Vp * 4;
        for (int i = tid; i < p.Vp * 4; i += IP_THREADS)
        {
            float v = s_pos[i];
            if (v != 0.f) red_add_global(gp + i, v);
        }
    }
}

// true when the templated kernels apply
static bool fast_ok(const InterpParams& p, bool da, const void* attr)
{
    if (p.A < 1 || p.A > 4) return false;
    if (da && !(p.diff_all && p.D == p.A)) return false;
    if (p.A == 2 && (((uintptr_t)attr) & 7)) return false;
    if (p.A == 4 && (((uintptr_t)attr) & 15)) return false;
    if ((size_t)p.H * p.W >= (1u << 30)) return false;
    return true;
}

#define IP_DISPATCH(KERNEL, A_, DA_, GRID, SMEM, STREAM, P)                                  \
    do {                                                                                    \
        switch ((A_) * 2 + ((DA_) ? 1 : 0)) {                                               \
        case 2: KERNEL<1, false><<<GRID, IP_THREADS, SMEM, STREAM>>>(P); break;             \
        case 3: KERNEL<1, true ><<<GRID, IP_THREADS, SMEM, STREAM>>>(P); break;             \
        case 4: KERNEL<2, false><<<GRID, IP_THREADS, SMEM, STREAM>>>(P); break;             \
        case 5: KERNEL<2, true ><<<GRID, IP_THREADS, SMEM, STREAM>>>(P); break;             \
        case 6: KERNEL<3, false><<<GRID, IP_THREADS, SMEM, STREAM>>>(P); break;             \
        case 7: KERNEL<3, true ><<<GRID, IP_THREADS, SMEM, STREAM>>>(P); break;             \
        case 8: KERNEL<4, false><<<GRID, IP_THREADS, SMEM, STREAM>>>(P); break;             \
        case 9: KERNEL<4, true ><<<GRID, IP_THREADS, SMEM, STREAM>>>(P); break;             \
        } } while (0)

typedef void (*IpgKernel)(const InterpParams);
// the backward kernel for (A, differentials, fused rasterize backward)
static IpgKernel grad_fast_kernel(int A, bool da, bool raster)
{
    switch (A * 4 + (da ? 2 : 0) + (raster ? 1 : 0)) {
    case  4: return interpolate_grad_fast<1, false, false>;  case  5: return interpolate_grad_fast<1, false, true>;
    case  6: return interpolate_grad_fast<1, true,  false>;  case  7: return interpolate_grad_fast<1, true,  true>;
    case  8: return interpolate_grad_fast<2, false, false>;  case  9: return interpolate_grad_fast<2, false, true>;
    case 10: return interpolate_grad_fast<2, true,  false>;  case 11: return interpolate_grad_fast<2, true,  true>;
    case 12: return interpolate_grad_fast<3, false, false>;  case 13: return interpolate_grad_fast<3, false, true>;
    case 14: return interpolate_grad_fast<3, true,  false>;  case 15: return interpolate_grad_fast<3, true,  true>;
    case 16: return interpolate_grad_fast<4, false, false>;  case 17: return interpolate_grad_fast<4, false, true>;
    case 18: return interpolate_grad_fast<4, true,  false>;  case 19: return interpolate_grad_fast<4, true,  true>;
    }
    return 0;
}

// Launch of the templated backward (p filled; p.pos != 0 selects the fused rasterize backward).  Shared accumulators (g_attr, and
// this image's g_pos when fused) are used when they fit; the grid is then exactly the resident wave (every CTA zeroes and flushes
// its copies once), rounded DOWN so that no CTA is left over for a second wave; global accumulation: a few waves.
static int launch_grad_fast(InterpParams& p, bool da, cudaStream_t stream, const char* fn)
{
    const bool raster = p.pos != 0;
    const int A = p.A, N = p.N;
    size_t stage_b = (size_t)(IP_THREADS / 32) * 32 * ((A == 3) ? (da ? 6 : 3) : 1) * sizeof(float);
    size_t acc_b = (size_t)p.V * A * sizeof(float);
    size_t pos_b = raster ? (size_t)p.Vp * 4 * sizeof(float) : 0;
    p.use_smem_acc = (acc_b + stage_b <= 64 * 1024 && p.V <= ACC_SMEM_MAX_V) ? 1 : 0;
    p.use_smem_pos = (raster && p.Vp <= IPC_SMEM_MAX_VP && pos_b + stage_b + (p.use_smem_acc ? acc_b : 0) <= 64 * 1024) ? 1 : 0;
    size_t smem_f = stage_b + (p.use_smem_acc ? acc_b : 0) + (p.use_smem_pos ? pos_b : 0);
    IpgKernel k = grad_fast_kernel(A, da, raster);
    if (smem_f > 48 * 1024) B200DR_CUDA(cudaFuncSetAttribute((const void*)k, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024), fn);
    int groups = (int)(((size_t)p.H * p.W + 31) / 32);
    const int unroll = raster ? IPC_UNROLL : (da ? IPG_UNROLL_DA : IPG_UNROLL_PLAIN);
    int gx = (groups + (IP_THREADS / 32) * unroll - 1) / ((IP_THREADS / 32) * unroll);
    int capx = (b200dr_num_sms() * 12 + N - 1) / N;
    if (p.use_smem_acc || p.use_smem_pos)
    {
        int resident = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, (const void*)k, IP_THREADS, smem_f);
        capx = max(b200dr_num_sms() * max(resident, 1) / N, 1);
    }
    if (gx > capx) gx = capx;
    if (gx < 1) gx = 1;
    k<<<dim3(gx, N), IP_THREADS, smem_f, stream>>>(p);
    return 0;
}

static int fill_params(InterpParams& p, const char* fn, const float* attr, const float* rast, const int32_t* tri, const float* rast_db,
                       int N, int H, int W, int V, int T, int A, int attr_bc, int instance_mode,
                       int diff_all, const int32_t* diff_list, int n_diff)
{
    B200DR_CHECK(N > 0 && H > 0 && W > 0, "%s: rast must have shape[>0, >0, >0, 4]", fn);
    B200DR_CHECK(V > 0 && A > 0, "%s: attr must have shape [>0, >0] or [>0, >0, >0]", fn);
    B200DR_CHECK(T > 0, "%s: tri must have shape [>0, 3]", fn);
    B200DR_CHECK(N <= 65535, "%s: minibatch size above 65535 is not supported", fn);
    B200DR_CHECK(B200DR_ALIGNED16(rast), "%s: rast input tensor not aligned to float4", fn);
    B200DR_CHECK(B200DR_ALIGNED16(rast_db), "%s: rast_db input tensor not aligned to float4", fn);
    B200DR_CHECK(diff_all || n_diff <= IP_MAX_DIFF, "%s: too many entries in diff_attrs list (increase IP_MAX_DIFF_ATTRS)", fn);
    B200DR_CHECK(n_diff == 0 || rast_db, "%s: rast_db is required when diff_attrs is given", fn);
    p.tri = tri; p.attr = attr; p.rast = rast; p.rast_db = rast_db;
    p.N = N; p.H = H; p.W = W; p.V = V; p.T = T; p.A = A;
    p.attr_bc = attr_bc; p.instance_mode = instance_mode; p.diff_all = diff_all;
    p.D = diff_all ? A : n_diff;
    if (!diff_all) for (int i = 0; i < n_diff; i++) p.diff[i] = diff_list[i];
    p.chunks = (int)(((size_t)H * W + IP_THREADS - 1) / IP_THREADS);
    return 0;
}

} // namespace

extern "C" int b200dr_interpolate_fwd(const float* attr, const float* rast, const int32_t* tri, const float* rast_db,
                                      int N, int H, int W, int V, int T, int A, int attr_bc, int instance_mode,
                                      int diff_all, const int32_t* diff_list, int n_diff,
                                      float* out, float* out_da, void* stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    InterpParams p = {};
    bool da = (diff_all || n_diff > 0);
    if (fill_params(p, da ? "interpolate_fwd_da()" : "interpolate_fwd()", attr, rast, tri, da ? rast_db : 0, N, H, W, V, T, A, attr_bc, instance_mode,
                    da ? diff_all : 0, diff_list, da ? n_diff : 0)) return 1;
    if (!da) p.D = 0;
    p.out = out; p.out_da = out_da;
    if (fast_ok(p, da, attr) && B200DR_ALIGNED16(out) && (!da || (((uintptr_t)out_da) & 31) == 0))
    {
        int groups = (int)(((size_t)H * W + 31) / 32);
        const int unroll = da ? IPF_UNROLL_DA : IPF_UNROLL_PLAIN;
        int gx = (groups + (IP_THREADS / 32) * unroll - 1) / ((IP_THREADS / 32) * unroll);
        int capx = (b200dr_num_sms() * 12 + N - 1) / N;
        if (gx > capx) gx = capx;
        if (gx < 1) gx = 1;
        IP_DISPATCH(interpolate_fwd_fast, A, da, dim3(gx, N), 0, stream, p);
        B200DR_CHECK_LAUNCH("interpolate_fwd()");
        return 0;
    }
    size_t smem = (size_t)IP_THREADS * (A + 2 * p.D) * sizeof(float);
    p.staged = smem <= 96 * 1024;
    if (!p.staged) smem = 0;
    int gx = p.chunks;
    int cap = (b200dr_num_sms() * 16 + N - 1) / N;
    if (gx > cap) gx = cap;
    dim3 grid(gx, N);
    if (da)
    {
        if (smem > 48 * 1024) B200DR_CUDA(cudaFuncSetAttribute(interpolate_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024), "interpolate()");
        interpolate_fwd_kernel<true><<<grid, IP_THREADS, smem, stream>>>(p);
    }
    else
    {
        if (smem > 48 * 1024) B200DR_CUDA(cudaFuncSetAttribute(interpolate_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024), "interpolate()");
        interpolate_fwd_kernel<false><<<grid, IP_THREADS, smem, stream>>>(p);
    }
    B200DR_CHECK_LAUNCH("interpolate_fwd()");
    return 0;
}

extern "C" int b200dr_interpolate_grad(const float* attr, const float* rast, const int32_t* tri, const float* dy,
                                       const float* rast_db, const float* dda,
                                       int N, int H, int W, int V, int T, int A, int attr_bc, int instance_mode,
                                       int diff_all, const int32_t* diff_list, int n_diff,
                                       float* g_attr, float* g_rast, float* g_rast_db, void* stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    InterpParams p = {};
    bool da = (diff_all || n_diff > 0);
    if (fill_params(p, da ? "interpolate_grad_da()" : "interpolate_grad()", attr, rast, tri, da ? rast_db : 0, N, H, W, V, T, A, attr_bc, instance_mode,
                    da ? diff_all : 0, diff_list, da ? n_diff : 0)) return 1;
    if (!da) p.D = 0;
    B200DR_CHECK(B200DR_ALIGNED16(g_rast) && B200DR_ALIGNED16(g_rast_db), "interpolate_grad(): gradient output tensors not aligned to float4");
    p.dy = dy; p.dda = dda; p.g_attr = g_attr; p.g_rast = g_rast; p.g_rast_db = g_rast_db;

    if (fast_ok(p, da, attr) && B200DR_ALIGNED16(dy) && B200DR_ALIGNED16(g_attr) && (!da || (((uintptr_t)dda) & 31) == 0))
    {
        if (launch_grad_fast(p, da, stream, "interpolate_grad()")) return 1;
        B200DR_CHECK_LAUNCH("interpolate_grad()");
        return 0;
    }
    size_t stage = (size_t)IP_THREADS * (A + 2 * p.D) * sizeof(float);
    p.staged = stage <= 32 * 1024;
    if (!p.staged) stage = 0;
    size_t accb = (size_t)V * A * sizeof(float);
    p.use_smem_acc = (accb + stage <= 96 * 1024) ? 1 : 0;
    size_t smem = stage + (p.use_smem_acc ? accb : 0);

    int gx;
    if (p.use_smem_acc) gx = (b200dr_num_sms() * 4 + N - 1) / N;
    else                gx = (b200dr_num_sms() * 16 + N - 1) / N;
    if (gx > p.chunks) gx = p.chunks;
    if (gx < 1) gx = 1;
    dim3 grid(gx, N);
    if (da)
    {
        if (smem > 48 * 1024) B200DR_CUDA(cudaFuncSetAttribute(interpolate_grad_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024), "interpolate()");
        interpolate_grad_kernel<true><<<grid, IP_THREADS, smem, stream>>>(p);
    }
    else
    {
        if (smem > 48 * 1024) B200DR_CUDA(cudaFuncSetAttribute(interpolate_grad_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024), "interpolate()");
        interpolate_grad_kernel<false><<<grid, IP_THREADS, smem, stream>>>(p);
    }
    B200DR_CHECK_LAUNCH("interpolate_grad()");
    return 0;
}

// EXTENSION (include/b200dr.h): backward of interpolate(attr, rasterize(pos)) in one pass.
extern "C" int b200dr_interpolate_chain_grad(const float* attr, const float* rast, const int32_t* tri, const float* dy,
                                             const float* rast_db, const float* dda,
                                             const float* pos, const int32_t* pos_tri, int Vp, int grad_db,
                                             int N, int H, int W, int V, int T, int A, int attr_bc, int instance_mode, int diff_all,
                                             float* g_attr, float* g_pos, void* stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    const char* fn = "interpolate_chain_grad()";
    InterpParams p = {};
    const bool da = diff_all != 0;
    if (fill_params(p, fn, attr, rast, tri, da ? rast_db : 0, N, H, W, V, T, A, attr_bc, instance_mode, da ? 1 : 0, 0, 0)) return 1;
    if (!da) p.D = 0;
    B200DR_CHECK(pos && pos_tri && g_pos && g_attr && dy && Vp > 0, "%s: missing input or output", fn);
    B200DR_CHECK(!da || (rast_db && dda), "%s: rast_db and dda are required with differentials", fn);
    B200DR_CHECK(fast_ok(p, da, attr), "%s: at most 4 attributes (all of them differentiated, or none), 8/16-byte aligned rows", fn);
    B200DR_CHECK(B200DR_ALIGNED16(dy) && B200DR_ALIGNED16(g_attr) && B200DR_ALIGNED16(pos) && B200DR_ALIGNED16(g_pos) && (!da || (((uintptr_t)dda) & 31) == 0),
                 "%s: tensor misaligned", fn);
    p.dy = dy; p.dda = dda; p.g_attr = g_attr;
    p.pos = pos; p.pos_tri = pos_tri; p.g_pos = g_pos; p.Vp = Vp; p.grad_db = grad_db;
    p.xs = 2.f / (float)W; p.xo = 1.f / (float)W - 1.f; p.ys = 2.f / (float)H; p.yo = 1.f / (float)H - 1.f;
    if (launch_grad_fast(p, da, stream, fn)) return 1;
    B200DR_CHECK_LAUNCH(fn);
    return 0;
}
