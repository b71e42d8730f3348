// antialias.cu -- B200-native replacement of the reference's antialias op:
//   antialias_construct_topology_hash (csrc/torch/torch_antialias.cpp:25-63,  csrc/common/antialias.cu:47-160)
//   antialias_fwd                     (csrc/torch/torch_antialias.cpp:68-155, csrc/common/antialias.cu:165-382)
//   antialias_grad                    (csrc/torch/torch_antialias.cpp:160-241, csrc/common/antialias.cu:387-556)
//
// Differences in structure (results are the same):
//   * topology is a per-triangle table of the three opposite ("wing") vertices, built once through a transient edge
//     hash; the per-pixel-pair analysis then needs one 16-byte load instead of three hash probes;
//   * the dense pass is fused: one kernel streams color -> out (the reference clones the tensor first), compares
//     triangle ids with the right/up neighbours and appends work items with one atomic per warp;
//   * position gradients are summed per run of equal (triangle, edge) inside the warp before the atomics.
#include "common.cuh"
#include "pixel_math.cuh"

namespace {

// ------------------------------------------------------------------------------------------------
// Topology

struct EdgeSlot { unsigned long long key; int a, b; };   // 16 bytes; a, b = opposite vertex + 1 (0 = empty)

static __device__ __forceinline__ unsigned long long mix64(unsigned long long x)
{
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
    return x;
}
static __device__ __forceinline__ unsigned long long edge_key(int va, int vb)
{
    return (unsigned long long)(unsigned)(min(va, vb) + 1) | ((unsigned long long)(unsigned)(max(va, vb) + 1) << 32);
}

static __device__ __forceinline__ void edge_insert(EdgeSlot* tab, unsigned long long mask, int va, int vb, int vn)
{
    if (va == vb) return;
    unsigned long long k = edge_key(va, vb);
    unsigned long long i = mix64(k) & mask;
    for (;;)
    {
        unsigned long long prev = atomicCAS(&tab[i].key, 0ULL, k);
        if (prev == 0ULL || prev == k) break;
        i = (i + 1) & mask;
    }
    int old = atomicCAS(&tab[i].a, 0, vn + 1);
    if (old != 0 && old != vn + 1) atomicCAS(&tab[i].b, 0, vn + 1);     // third and later triangles on an edge are dropped
}

static __device__ __forceinline__ int edge_find(const EdgeSlot* tab, unsigned long long mask, int va, int vb, int vr)
{
    if (va == vb) return -1;
    unsigned long long k = edge_key(va, vb);
    unsigned long long i = mix64(k) & mask;
    for (;;)
    {
        unsigned long long kk = tab[i].key;
        if (kk == k || kk == 0ULL) break;
        i = (i + 1) & mask;
    }
    int a = tab[i].a - 1, b = tab[i].b - 1;
    if (a == vr) return b;
    if (b == vr) return a;
    return -1;
}

static __device__ __forceinline__ bool load_tri(const int32_t* tri, int t, int V, int& v0, int& v1, int& v2)
{
    v0 = __ldg(tri + t * 3 + 0); v1 = __ldg(tri + t * 3 + 1); v2 = __ldg(tri + t * 3 + 2);
    return !(v0 < 0 || v0 >= V || v1 < 0 || v1 >= V || v2 < 0 || v2 >= V);
}

__global__ void __launch_bounds__(256) topology_insert_kernel(const int32_t* tri, int T, int V, EdgeSlot* tab, unsigned long long mask)
{
    int t = blockIdx.x * 256 + threadIdx.x;
    if (t >= T) return;
    int v0, v1, v2;
    if (!load_tri(tri, t, V, v0, v1, v2)) return;
    if (v0 == v1 || v1 == v2 || v2 == v0) return;
    edge_insert(tab, mask, v1, v2, v0);
    edge_insert(tab, mask, v2, v0, v1);
    edge_insert(tab, mask, v0, v1, v2);
}

__global__ void __launch_bounds__(256) topology_wings_kernel(const int32_t* tri, int T, int V, const EdgeSlot* tab, unsigned long long mask, int4* wings)
{
    int t = blockIdx.x * 256 + threadIdx.x;
    if (t >= T) return;
    int v0, v1, v2;
    int4 w = make_int4(-1, -1, -1, 0);
    if (load_tri(tri, t, V, v0, v1, v2))
    {
        w.x = edge_find(tab, mask, v2, v1, v0);
        w.y = edge_find(tab, mask, v0, v2, v1);
        w.z = edge_find(tab, mask, v1, v0, v2);
    }
    wings[t] = w;
}

static unsigned long long topology_slots(int T)
{
    unsigned long long n = 64;
    while (n < (unsigned long long)T) n <<= 1;
    return n * 4;       // <= 3T distinct edges: load factor <= 0.75, typically < 0.4
}

// ------------------------------------------------------------------------------------------------
// Work items.  int4: x = px | py << 16, y = pz, z = flags, w = alpha bits (0 = no blend).

#define AA_FLAG_DOWN_BIT 2
#define AA_FLAG_TRI1_BIT 3
#define AA_EDGE_MASK     3

struct AAParams
{
    const float* color; const float* rast; const int32_t* tri; const float* pos; const int4* wings;
    float* out; const float* dy; float* g_color; float* g_pos;
    int4* items; int* counters;         // counters[0] = number of items
    int N, H, W, C, V, T, instance_mode;
    int vec_ok;                         // color/out are 16-byte aligned
    float xh, yh;
};

// ------------------------------------------------------------------------------------------------
// Dense pass: out = color, find id discontinuities (csrc/common/antialias.cu:165-214).

// A thread owns a vertical strip of AA_ROWS pixels of one column (a warp: 32 columns x AA_ROWS rows), a CTA of 8 warps a
// 128 x 16 pixel tile.  All of a thread's loads (AA_ROWS + 1 triangle ids, AA_ROWS colour vectors) are issued before the
// first use, the "next row" neighbour is the thread's own next register, the right neighbour a shuffle; no shared memory,
// no block barrier, one atomic per warp for all the work items of its 32 x AA_ROWS pixels.
#ifndef AA_ROWS
#define AA_ROWS   8
#endif
#ifndef AA_MINB
#define AA_MINB   1
#endif
#define AA_WARPS_X 4
#define AA_WARPS_Y 2
#define AA_TILE_W (32 * AA_WARPS_X)
#define AA_TILE_H (AA_ROWS * AA_WARPS_Y)

__global__ void __launch_bounds__(256, AA_MINB) aa_discontinuity_kernel(const AAParams p)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int x0 = blockIdx.x * AA_TILE_W + (warp % AA_WARPS_X) * 32;       // first column of this warp
    const int y0 = blockIdx.y * AA_TILE_H + (warp / AA_WARPS_X) * AA_ROWS;   // first row of this warp
    const int pz = blockIdx.z;
    const int px = x0 + lane;
    if (x0 >= p.W || y0 >= p.H) return;                                     // whole warp outside the image
    const bool colin = px < p.W;
    const size_t img = (size_t)pz * p.H * p.W;
    const float* idp = p.rast + (img + (size_t)y0 * p.W + min(px, p.W - 1)) * 4 + 3;

    // ---- issue every load of this strip: ids of rows y0 .. y0+AA_ROWS (clamped), right-hand column for lane 31, colours
    float id[AA_ROWS + 1], idr[AA_ROWS];
    #pragma unroll
    for (int r = 0; r <= AA_ROWS; r++)
    {
        int y = min(y0 + r, p.H - 1);
        id[r] = __ldg(idp + (size_t)(y - y0) * p.W * 4);
    }
    const bool need_right = (lane == 31) && (px < p.W - 1);
    #pragma unroll
    for (int r = 0; r < AA_ROWS; r++)
    {
        idr[r] = 0.f;
        if (need_right && y0 + r < p.H) idr[r] = __ldg(idp + (size_t)r * p.W * 4 + 4);
    }

    const int nflt = min(32, p.W - x0) * p.C;                               // floats of one row segment of this warp
    const bool vec = p.vec_ok && ((p.W * p.C) & 3) == 0 && ((x0 * p.C) & 3) == 0;
    if (vec && nflt <= 128)
    {
        float4 c[AA_ROWS];
        const int q = lane * 4;
        #pragma unroll
        for (int r = 0; r < AA_ROWS; r++)
            if (y0 + r < p.H && q < nflt) c[r] = ld_stream_f4((const float4*)(p.color + (img + (size_t)(y0 + r) * p.W + x0) * p.C) + lane);
        #pragma unroll
        for (int r = 0; r < AA_ROWS; r++)
            if (y0 + r < p.H && q < nflt) st_stream_f4((float4*)(p.out + (img + (size_t)(y0 + r) * p.W + x0) * p.C) + lane, c[r]);
    }
    else
    {
        for (int r = 0; r < AA_ROWS && y0 + r < p.H; r++)
        {
            size_t base = (img + (size_t)(y0 + r) * p.W + x0) * p.C;
            if (vec) for (int i = lane; i < (nflt >> 2); i += 32) st_stream_f4((float4*)(p.out + base) + i, ld_stream_f4((const float4*)(p.color + base) + i));
            else     for (int i = lane; i < nflt; i += 32) st_stream_f(p.out + base + i, ld_stream_f(p.color + base + i));
        }
    }

    // ---- discontinuities (csrc/common/antialias.cu:165-214): right neighbour (clamped at the image edge), next row
    unsigned m1[AA_ROWS], m2[AA_ROWS];
    int total = 0;
    #pragma unroll
    for (int r = 0; r < AA_ROWS; r++)
    {
        float right = __shfl_down_sync(0xffffffffu, id[r], 1);
        if (lane == 31) right = idr[r];
        bool rowin = (y0 + r < p.H);
        bool f1 = colin && rowin && (px < p.W - 1) && (right != id[r]);
        bool f2 = colin && rowin && (y0 + r < p.H - 1) && (id[r + 1] != id[r]);
        m1[r] = __ballot_sync(0xffffffffu, f1);
        m2[r] = __ballot_sync(0xffffffffu, f2);
        total += __popc(m1[r]) + __popc(m2[r]);
    }
    if (total == 0) return;                                                  // warp-uniform
    int base = 0;
    if (lane == 0) base = atomicAdd(p.counters, total);
    base = __shfl_sync(0xffffffffu, base, 0);
    const unsigned lt = (1u << lane) - 1u;
    #pragma unroll
    for (int r = 0; r < AA_ROWS; r++)
    {
        int pxy = px | ((y0 + r) << 16);
        if ((m1[r] >> lane) & 1u) p.items[base + __popc(m1[r] & lt)] = make_int4(pxy, pz, 0, 0);
        base += __popc(m1[r]);
        if ((m2[r] >> lane) & 1u) p.items[base + __popc(m2[r] & lt)] = make_int4(pxy, pz, 1 << AA_FLAG_DOWN_BIT, 0);
        base += __popc(m2[r]);
    }
}

// ------------------------------------------------------------------------------------------------
// Sparse pass: silhouette analysis and blending (csrc/common/antialias.cu:219-382).

__global__ void __launch_bounds__(256) aa_analysis_kernel(const AAParams p)
{
    const int count = min(p.counters[0], 2 * p.N * p.H * p.W);
    for (int idx = blockIdx.x * 256 + threadIdx.x; idx < count; idx += gridDim.x * 256)
    {
        int4 item = p.items[idx];
        int px = item.x & 0xFFFF, py = (int)((unsigned)item.x >> 16), pz = item.y;
        int d = (item.z >> AA_FLAG_DOWN_BIT) & 1;

        int pixel0 = px + p.W * (py + p.H * pz);
        int pixel1 = pixel0 + (d ? p.W : 1);
        float2 zt0 = __ldg((const float2*)p.rast + ((size_t)pixel0 << 1) + 1);
        float2 zt1 = __ldg((const float2*)p.rast + ((size_t)pixel1 << 1) + 1);
        int tri0 = float_to_triidx(zt0.y) - 1;
        int tri1 = float_to_triidx(zt1.y) - 1;

        // nearer of the two surfaces (background loses)
        int tri = (tri0 >= 0) ? tri0 : tri1;
        if (tri0 >= 0 && tri1 >= 0) tri = (zt0.x < zt1.x) ? tri0 : tri1;
        if (tri == tri1) { px += 1 - d; py += d; }
        if (tri < 0 || tri >= p.T) continue;

        int vi[3];
        if (!load_tri(p.tri, tri, p.V, vi[0], vi[1], vi[2])) continue;
        const int4 wg = __ldg(p.wings + tri);
        const int wv[3] = { wg.x, wg.y, wg.z };
        const float4* vb = (const float4*)p.pos + (p.instance_mode ? (size_t)pz * p.V : 0);
        float4 vtx[3], wing[3];
        #pragma unroll
        for (int i = 0; i < 3; i++) vtx[i] = __ldg(vb + vi[i]);
        #pragma unroll
        for (int i = 0; i < 3; i++) wing[i] = (wv[i] < 0) ? vtx[i] : __ldg(vb + wv[i]);   // open edge: counts as a silhouette

        const pm::AAEdgeResult r = pm::aa_edge(vtx, wing, px, py, p.xh, p.yh, d != 0, tri == tri0);
        if (r.blend)
        {
            const float alpha = r.alpha;
            const int di = r.edge;
            const float ds = (tri == tri0) ? 1.f : -1.f;
            const float* c0 = p.color + (size_t)pixel0 * p.C;
            const float* c1 = p.color + (size_t)pixel1 * p.C;
            float* o = p.out + (size_t)(alpha > 0.f ? pixel0 : pixel1) * p.C;
            for (int i = 0; i < p.C; i++) red_add_global(o + i, alpha * (__ldg(c1 + i) - __ldg(c0 + i)));

            unsigned flags = (unsigned)di | ((unsigned)d << AA_FLAG_DOWN_BIT) | ((__float_as_uint(ds) >> 31) << AA_FLAG_TRI1_BIT);
            ((int2*)(p.items + idx))[1] = make_int2((int)flags, __float_as_int(alpha));
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Backward (csrc/common/antialias.cu:387-556).

__global__ void __launch_bounds__(256) aa_copy_kernel(const float* __restrict__ src, float* __restrict__ dst, size_t n)
{
    size_t n4 = n >> 2;
    const float4* s4 = (const float4*)src; float4* d4 = (float4*)dst;
    size_t stride = (size_t)gridDim.x * 256;
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n4; i += stride) st_stream_f4(d4 + i, ld_stream_f4(s4 + i));
    size_t i = (n4 << 2) + (size_t)blockIdx.x * 256 + threadIdx.x;
    if (i < n) dst[i] = src[i];
}

__global__ void __launch_bounds__(256) aa_grad_kernel(const AAParams p)
{
    const int count = min(p.counters[0], 2 * p.N * p.H * p.W);
    const int rounds = (count + gridDim.x * 256 - 1) / (gridDim.x * 256);
    for (int r = 0; r < rounds; r++)                   // whole warps iterate together (warp collectives below)
    {
        int idx = (r * gridDim.x + blockIdx.x) * 256 + threadIdx.x;
        int4 item = (idx < count) ? p.items[idx] : make_int4(0, 0, 0, 0);
        bool active = (idx < count) && (item.w != 0);

        int px = item.x & 0xFFFF, py = (int)((unsigned)item.x >> 16), pz = item.y;
        int d = (item.z >> AA_FLAG_DOWN_BIT) & 1;
        float alpha = __int_as_float(item.w);
        int tri1f = (item.z >> AA_FLAG_TRI1_BIT) & 1;
        int di = item.z & AA_EDGE_MASK;
        int pixel0 = px + p.W * (py + p.H * pz);
        int pixel1 = pixel0 + (d ? p.W : 1);
        int tri = -1;
        if (active)
        {
            tri = float_to_triidx(__ldg(p.rast + ((size_t)(tri1f ? pixel1 : pixel0) << 2) + 3)) - 1;
            if (tri < 0 || tri >= p.T) active = false;
        }
        if (tri1f) { px += 1 - d; py += d; }

        float dd = 0.f;
        if (active)
        {
            float* g0 = p.g_color + (size_t)pixel0 * p.C;
            float* g1 = p.g_color + (size_t)pixel1 * p.C;
            const float* pdy = p.dy + (size_t)(alpha > 0.f ? pixel0 : pixel1) * p.C;
            const float* c0 = p.color + (size_t)pixel0 * p.C;
            const float* c1 = p.color + (size_t)pixel1 * p.C;
            for (int i = 0; i < p.C; i++)
            {
                float dy = __ldg(pdy + i);
                if (dy != 0.f)
                {
                    dd += dy * (__ldg(c1 + i) - __ldg(c0 + i));
                    float v = alpha * dy;
                    red_add_global(g0 + i, -v);
                    red_add_global(g1 + i, v);
                }
            }
            if (dd == 0.f) active = false;
        }

        int vi1 = 0, vi2 = 0;
        if (active)
        {
            int i1 = (di < 2) ? (di + 1) : 0;
            int i2 = (i1 < 2) ? (i1 + 1) : 0;
            vi1 = __ldg(p.tri + 3 * tri + i1);
            vi2 = __ldg(p.tri + 3 * tri + i2);
            if (vi1 < 0 || vi1 >= p.V || vi2 < 0 || vi2 >= p.V) active = false;
            else if (p.instance_mode) { vi1 += pz * p.V; vi2 += pz * p.V; }
        }

        float g[6] = { 0.f, 0.f, 0.f, 0.f, 0.f, 0.f };
        if (active)
        {
            float g1[3], g2[3];
            pm::aa_edge_grad(__ldg((const float4*)p.pos + vi1), __ldg((const float4*)p.pos + vi2), px, py, p.xh, p.yh, d != 0, alpha, dd, g1, g2);
            g[0] = g1[0]; g[1] = g1[1]; g[2] = g1[2]; g[3] = g2[0]; g[4] = g2[1]; g[5] = g2[2];
        }

        // consecutive items often hit the same edge (same two vertices): sum such runs before the atomics
        if (__any_sync(0xffffffffu, active))
        {
            WarpRuns runs = warp_runs2(active ? vi1 : -1, active ? vi2 : -1);
            #pragma unroll
            for (int i = 0; i < 6; i++) g[i] = warp_run_sum(g[i], runs);
            if (runs.is_head && active)
            {
                float* q1 = p.g_pos + 4 * (size_t)vi1; float* q2 = p.g_pos + 4 * (size_t)vi2;
                red_add_global(q1 + 0, g[0]); red_add_global(q1 + 1, g[1]); red_add_global(q1 + 3, g[2]);
                red_add_global(q2 + 0, g[3]); red_add_global(q2 + 1, g[4]); red_add_global(q2 + 3, g[5]);
            }
        }
    }
}

static int aa_fill(AAParams& p, const char* fn, const float* color, const float* rast, const float* pos, const int32_t* tri,
                   int N, int H, int W, int C, int V, int T, int instance_mode)
{
    B200DR_CHECK(N > 0 && H > 0 && W > 0 && C > 0, "%s: color must have shape[>0, >0, >0, >0]", fn);
    B200DR_CHECK(V > 0 && T > 0, "%s: pos/tri must be non-empty", fn);
    B200DR_CHECK(H < 65536 && W < 65536 && N <= 65535, "%s: image size or minibatch too large", fn);
    B200DR_CHECK((size_t)N * H * W < (1ull << 30), "%s: too many pixels for 32-bit work indexing", fn);
    B200DR_CHECK(B200DR_ALIGNED16(pos), "%s: pos input tensor not aligned to float4", fn);
    B200DR_CHECK((((uintptr_t)rast) & 7) == 0, "%s: raster_out input tensor not aligned to float2", fn);
    p.color = color; p.rast = rast; p.pos = pos; p.tri = tri;
    p.N = N; p.H = H; p.W = W; p.C = C; p.V = V; p.T = T; p.instance_mode = instance_mode;
    p.xh = .5f * (float)W; p.yh = .5f * (float)H;
    return 0;
}

} // namespace

extern "C" size_t b200dr_topology_scratch_bytes(int T)
{
    return (size_t)topology_slots(T) * sizeof(EdgeSlot);
}

extern "C" int b200dr_antialias_construct_topology(const int32_t* tri, int T, int V, int32_t* wings,
                                                   void* scratch, size_t scratch_bytes, void* stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    B200DR_CHECK(T > 0, "antialias_construct_topology_hash(): tri must have shape [>0, 3]");
    B200DR_CHECK(scratch && scratch_bytes >= b200dr_topology_scratch_bytes(T), "antialias_construct_topology_hash(): scratch buffer too small");
    B200DR_CHECK(B200DR_ALIGNED16(scratch) && B200DR_ALIGNED16(wings), "antialias_construct_topology_hash(): internal tensor not aligned to int4");
    unsigned long long slots = topology_slots(T);
    B200DR_CUDA(cudaMemsetAsync(scratch, 0, slots * sizeof(EdgeSlot), stream), "antialias()");
    int blocks = (T + 255) / 256;
    topology_insert_kernel<<<blocks, 256, 0, stream>>>(tri, T, V, (EdgeSlot*)scratch, slots - 1);
    topology_wings_kernel<<<blocks, 256, 0, stream>>>(tri, T, V, (const EdgeSlot*)scratch, slots - 1, (int4*)wings);
    b200dr_count_launches(1);   // two kernels, one check below
    B200DR_CHECK_LAUNCH("antialias_construct_topology_hash()");
    return 0;
}

extern "C" int b200dr_antialias_fwd(const float* color, const float* rast, const float* pos, const int32_t* tri, const int32_t* wings,
                                    int N, int H, int W, int C, int V, int T, int instance_mode,
                                    float* out, void* work_items, int32_t* work_counters, void* stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    AAParams p = {};
    if (aa_fill(p, "antialias_fwd()", color, rast, pos, tri, N, H, W, C, V, T, instance_mode)) return 1;
    B200DR_CHECK(B200DR_ALIGNED16(work_items), "antialias_fwd(): work_buffer internal tensor not aligned to int4");
    p.wings = (const int4*)wings; p.out = out; p.vec_ok = B200DR_ALIGNED16(color) && B200DR_ALIGNED16(out); p.items = (int4*)work_items; p.counters = work_counters;
    B200DR_CUDA(cudaMemsetAsync(work_counters, 0, 4 * sizeof(int), stream), "antialias()");
    dim3 grid((W + AA_TILE_W - 1) / AA_TILE_W, (H + AA_TILE_H - 1) / AA_TILE_H, N);
    aa_discontinuity_kernel<<<grid, 256, 0, stream>>>(p);
    B200DR_CHECK_LAUNCH("antialias_fwd(discontinuity)");
    aa_analysis_kernel<<<b200dr_num_sms() * 8, 256, 0, stream>>>(p);
    B200DR_CHECK_LAUNCH("antialias_fwd(analysis)");
    return 0;
}

extern "C" int b200dr_antialias_grad(const float* color, const float* rast, const float* pos, const int32_t* tri, const float* dy,
                                     const void* work_items, const int32_t* work_counters,
                                     int N, int H, int W, int C, int V, int T, int instance_mode,
                                     float* g_color, float* g_pos, void* stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    AAParams p = {};
    if (aa_fill(p, "antialias_grad()", color, rast, pos, tri, N, H, W, C, V, T, instance_mode)) return 1;
    p.dy = dy; p.g_color = g_color; p.g_pos = g_pos; p.items = (int4*)work_items; p.counters = (int*)work_counters;
    size_t n = (size_t)N * H * W * C;
    if (B200DR_ALIGNED16(dy) && B200DR_ALIGNED16(g_color))
    {
        size_t blocks = ((n >> 2) + 255) / 256;
        size_t cap = (size_t)b200dr_num_sms() * 32;
        if (blocks > cap) blocks = cap;
        if (blocks < 1) blocks = 1;
        aa_copy_kernel<<<(unsigned)blocks, 256, 0, stream>>>(dy, g_color, n);
        B200DR_CHECK_LAUNCH("antialias_grad(copy)");
    }
    else
        B200DR_CUDA(cudaMemcpyAsync(g_color, dy, n * sizeof(float), cudaMemcpyDeviceToDevice, stream), "antialias()");
    aa_grad_kernel<<<b200dr_num_sms() * 8, 256, 0, stream>>>(p);
    B200DR_CHECK_LAUNCH("antialias_grad()");
    return 0;
}
