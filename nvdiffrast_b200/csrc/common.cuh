// Shared device/host helpers for libb200dr (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "b200dr.h"

// ------------------------------------------------------------------------------------------------
// Host-side error plumbing (C ABI returns int, message kept thread-local).

void b200dr_set_error(const char* fmt, ...);

#define B200DR_CHECK(cond, ...)                                        \
    do { if (!(cond)) { b200dr_set_error(__VA_ARGS__); return 1; } } while (0)

#define B200DR_CHECK_LAUNCH(name)                                                            \
    do { b200dr_count_launches(1); cudaError_t e_ = cudaGetLastError();                                               \
         if (e_ != cudaSuccess) { b200dr_set_error("%s: Cuda error: %s", name, cudaGetErrorString(e_)); return 1; } } while (0)

// every CUDA runtime call that can fail is checked: B200DR_CUDA(cudaMemsetAsync(...), "op()")
#define B200DR_CUDA(call, name)                                                                   \
    do { cudaError_t e_ = (call);                                                                 \
         if (e_ != cudaSuccess) { b200dr_set_error("%s: Cuda error: %s (%s)", name, cudaGetErrorString(e_), #call); return 1; } } while (0)

#define B200DR_ALIGNED16(p) ((((uintptr_t)(p)) & 15) == 0)

void b200dr_count_launches(int n);

// B200: 148 SMs.  Queried once per device at first use.
int b200dr_num_sms(void);

// Grid size of a persistent, warp-autonomous kernel over the flat pixel array.  A warp visits the 32-pixel groups g, g + S,
// g + 2S, ... with S = blocks * warps_per_cta.  When S is a multiple of the groups per image row, every warp stays in ONE
// column of the image for its whole life -- and columns differ a lot in cost (background at the borders, geometry in the
// middle), so the middle columns' warps set the kernel time while the others idle.  Returns the largest block count <= cap
// (searching 16 below) whose stride visits as many distinct columns as the CTA shape allows.
static inline size_t b200dr_balanced_blocks(size_t cap, int warps_per_cta, int W)
{
    if (cap <= 1 || W <= 0 || (W & 31)) return cap;            // rows that are not whole groups drift on their own
    const size_t gpr = (size_t)W / 32;
    if (gpr <= 1) return cap;
    size_t best = cap, best_gcd = (size_t)-1;
    for (size_t b = cap; b > 0 && b + 16 > cap; b--)
    {
        size_t x = (b * (size_t)warps_per_cta) % gpr, y = gpr;
        while (x) { size_t t = y % x; y = x; x = t; }           // y = gcd(stride mod gpr, gpr); stride = 0 gives gpr (pinned)
        if (y < best_gcd) { best_gcd = y; best = b; }
    }
    return best;
}

#ifdef __CUDACC__

// ------------------------------------------------------------------------------------------------
// Constants shared with the oracle (csrc/common/cudaraster/impl/Constants.hpp:69-71).

#define DR_DEPTH_MIN 17600u
#define DR_DEPTH_MAX (0xFFFFFFFFu - 17600u)

// ------------------------------------------------------------------------------------------------
// Triangle id <-> float codec (csrc/common/common.h:192-193).

static __device__ __forceinline__ int   float_to_triidx(float x) { if (x <= 16777216.f) return (int)x; return __float_as_int(x) - 0x4a800000; }
static __device__ __forceinline__ float triidx_to_float(int x)   { if (x <= 0x01000000) return (float)x; return __int_as_float(0x4a800000 + x); }

// ------------------------------------------------------------------------------------------------
// Streaming loads / stores.  Per-pixel tensors are touched exactly once per kernel: keep them out of L1
// (L2::evict_first needs 256-bit accesses on sm_100a; plain no_allocate is used here).

static __device__ __forceinline__ float4 ld_stream_f4(const float4* p)
{
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
static __device__ __forceinline__ float2 ld_stream_f2(const float2* p)
{
    float2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0,%1}, [%2];" : "=f"(r.x), "=f"(r.y) : "l"(p));
    return r;
}
static __device__ __forceinline__ float ld_stream_f(const float* p)
{
    float r;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(r) : "l"(p));
    return r;
}
static __device__ __forceinline__ void st_stream_f4(float4* p, float4 v)
{
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
static __device__ __forceinline__ void st_stream_f2(float2* p, float2 v)
{
    asm volatile("st.global.L1::no_allocate.v2.f32 [%0], {%1,%2};" :: "l"(p), "f"(v.x), "f"(v.y) : "memory");
}
static __device__ __forceinline__ void st_stream_f(float* p, float v)
{
    asm volatile("st.global.L1::no_allocate.f32 [%0], %1;" :: "l"(p), "f"(v) : "memory");
}
static __device__ __forceinline__ void st_stream_u4(uint4* p, uint4 v)
{
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// 256-bit accesses (LDG.E.256 / STG.E.256, new on sm_100): two float4 pixels per instruction, and the only access
// width that accepts an L2 eviction priority.  Address must be 32-byte aligned.
static __device__ __forceinline__ void ld_stream_f8(const float* p, float4& a, float4& b)
{
    asm volatile("ld.global.nc.L1::no_allocate.L2::evict_first.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "l"(p));
}
static __device__ __forceinline__ void st_stream_f8(float* p, float4 a, float4 b)
{
    asm volatile("st.global.L1::no_allocate.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 :: "l"(p), "f"(a.x), "f"(a.y), "f"(a.z), "f"(a.w), "f"(b.x), "f"(b.y), "f"(b.z), "f"(b.w) : "memory");
}

// ------------------------------------------------------------------------------------------------
// Warp-cooperative AoS access: lane l owns K consecutive floats at base[l*K .. l*K+K) (32 pixels x K channels,
// contiguous).  K = 1, 2, 4, 8 map to one vector access per lane.  Other K go through a warp-private shared-memory
// transpose so that global memory still sees full 128-bit coalesced accesses; `s` points to >= 32*K floats owned by this
// warp (16-byte aligned).  `full`: all 32 pixels valid and `base` 16-byte aligned; otherwise a scalar path is taken.

template <int K>
static __device__ __forceinline__ void warp_store_aos(float* base, const float (&v)[K], int lane, bool valid, bool full, float* s)
{
    if (K == 1) { if (valid) st_stream_f(base + lane, v[0]); }
    else if (K == 2) { if (valid) st_stream_f2((float2*)base + lane, make_float2(v[0], v[1])); }
    else if (K == 4) { if (valid) st_stream_f4((float4*)base + lane, make_float4(v[0], v[1], v[2], v[3])); }
    else if (K == 8) { if (valid) st_stream_f8(base + 8 * lane, make_float4(v[0], v[1], v[2], v[3]), make_float4(v[4], v[5], v[6], v[7])); }
    else if (full)
    {
        #pragma unroll
        for (int i = 0; i < K; i++) s[lane * K + i] = v[i];
        __syncwarp();
        #pragma unroll
        for (int j = 0; j < (K * 32 + 127) / 128; j++)
        {
            int q = lane + 32 * j;
            if (q * 4 < K * 32) st_stream_f4((float4*)base + q, *(const float4*)(s + 4 * q));
        }
        __syncwarp();
    }
    else if (valid)
    {
        #pragma unroll
        for (int i = 0; i < K; i++) st_stream_f(base + lane * K + i, v[i]);
    }
}

template <int K>
static __device__ __forceinline__ void warp_load_aos(const float* base, float (&v)[K], int lane, bool valid, bool full, float* s)
{
    if (K == 1) { if (valid) v[0] = ld_stream_f(base + lane); }
    else if (K == 2) { if (valid) { float2 t = ld_stream_f2((const float2*)base + lane); v[0] = t.x; v[1] = t.y; } }
    else if (K == 4) { if (valid) { float4 t = ld_stream_f4((const float4*)base + lane); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; } }
    else if (K == 8) { if (valid) { float4 a, b; ld_stream_f8(base + 8 * lane, a, b); v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w; } }
    else if (full)
    {
        #pragma unroll
        for (int j = 0; j < (K * 32 + 127) / 128; j++)
        {
            int q = lane + 32 * j;
            if (q * 4 < K * 32) *(float4*)(s + 4 * q) = ld_stream_f4((const float4*)base + q);
        }
        __syncwarp();
        #pragma unroll
        for (int i = 0; i < K; i++) v[i] = s[lane * K + i];
        __syncwarp();
    }
    else if (valid)
    {
        #pragma unroll
        for (int i = 0; i < K; i++) v[i] = ld_stream_f(base + lane * K + i);
    }
}

// Split form of warp_load_aos for software pipelining: `issue` starts the global loads of a group (results stay in
// registers), `finish` turns them into the lane's K values (through the warp transpose when K is not a power of two).
template <int K> struct AosRaw { float4 q[(K == 1 || K == 2 || K == 4) ? 1 : (K == 8 ? 2 : (K * 32 + 127) / 128)]; };

template <int K>
static __device__ __forceinline__ void warp_load_aos_issue(const float* base, AosRaw<K>& raw, int lane, bool valid, bool full)
{
    raw.q[0] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (K == 1) { if (valid) raw.q[0].x = ld_stream_f(base + lane); }
    else if (K == 2) { if (valid) { float2 t = ld_stream_f2((const float2*)base + lane); raw.q[0].x = t.x; raw.q[0].y = t.y; } }
    else if (K == 4) { if (valid) raw.q[0] = ld_stream_f4((const float4*)base + lane); }
    else if (K == 8) { raw.q[1] = raw.q[0]; if (valid) ld_stream_f8(base + 8 * lane, raw.q[0], raw.q[1]); }
    else if (full)
    {
        #pragma unroll
        for (int j = 0; j < (K * 32 + 127) / 128; j++)
        {
            int q = lane + 32 * j;
            raw.q[j] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (q * 4 < K * 32) raw.q[j] = ld_stream_f4((const float4*)base + q);
        }
    }
}

template <int K>
static __device__ __forceinline__ void warp_load_aos_finish(const float* base, const AosRaw<K>& raw, float (&v)[K], int lane, bool valid, bool full, float* s)
{
    if (K == 1) { v[0] = raw.q[0].x; }
    else if (K == 2) { v[0] = raw.q[0].x; v[1 % K] = raw.q[0].y; }
    else if (K == 4) { v[0] = raw.q[0].x; v[1 % K] = raw.q[0].y; v[2 % K] = raw.q[0].z; v[3 % K] = raw.q[0].w; }
    else if (K == 8)
    {
        v[0] = raw.q[0].x; v[1 % K] = raw.q[0].y; v[2 % K] = raw.q[0].z; v[3 % K] = raw.q[0].w;
        v[4 % K] = raw.q[(K == 8) ? 1 : 0].x; v[5 % K] = raw.q[(K == 8) ? 1 : 0].y; v[6 % K] = raw.q[(K == 8) ? 1 : 0].z; v[7 % K] = raw.q[(K == 8) ? 1 : 0].w;
    }
    else if (full)
    {
        #pragma unroll
        for (int j = 0; j < (K * 32 + 127) / 128; j++)
        {
            int q = lane + 32 * j;
            if (q * 4 < K * 32) *(float4*)(s + 4 * q) = raw.q[j];
        }
        __syncwarp();
        #pragma unroll
        for (int i = 0; i < K; i++) v[i] = s[lane * K + i];
        __syncwarp();
    }
    else
    {
        #pragma unroll
        for (int i = 0; i < K; i++) v[i] = valid ? ld_stream_f(base + lane * K + i) : 0.f;
    }
}

// ------------------------------------------------------------------------------------------------
// Warp-level segmented ("run") reduction.
//
// Gradient kernels see 32 consecutive pixels per warp; pixels of one triangle form runs.  Instead of one atomic
// per pixel per value (or the reference's __match_any + shared-memory atomics, csrc/common/common.h:198-241),
// every run is summed with 5 shuffles and only the run's head lane issues the atomic.

struct WarpRuns
{
    unsigned heads;     // bit l set: lane l starts a run
    int      run_end;   // first lane after this lane's run (1..32)
    bool     is_head;
};

// key: any int identifying the group of a lane (all lanes of the warp must call).
static __device__ __forceinline__ WarpRuns warp_runs(int key)
{
    const int lane = threadIdx.x & 31;
    int prev = __shfl_up_sync(0xffffffffu, key, 1);
    bool head = (lane == 0) || (prev != key);
    WarpRuns r;
    r.heads = __ballot_sync(0xffffffffu, head);
    unsigned above = (lane == 31) ? 0u : (r.heads >> (lane + 1));
    r.run_end = above ? (lane + __ffs(above)) : 32;
    r.is_head = head;
    return r;
}

// Same, runs of equal (key0, key1) pairs.
static __device__ __forceinline__ WarpRuns warp_runs2(int key0, int key1)
{
    const int lane = threadIdx.x & 31;
    int prev0 = __shfl_up_sync(0xffffffffu, key0, 1);
    int prev1 = __shfl_up_sync(0xffffffffu, key1, 1);
    bool head = (lane == 0) || (prev0 != key0) || (prev1 != key1);
    WarpRuns r;
    r.heads = __ballot_sync(0xffffffffu, head);
    unsigned above = (lane == 31) ? 0u : (r.heads >> (lane + 1));
    r.run_end = above ? (lane + __ffs(above)) : 32;
    r.is_head = head;
    return r;
}

// Sum v over each run; the result is valid in the run's head lane.
// (Measured, no gain: folding the per-step predicate into an FFMA with a 0/1 factor -- SHFL + FFMA instead of SHFL + FADD + FSEL.)
static __device__ __forceinline__ float warp_run_sum(float v, const WarpRuns& r)
{
    const int lane = threadIdx.x & 31;
    #pragma unroll
    for (int d = 1; d < 32; d <<= 1)
    {
        float t = __shfl_down_sync(0xffffffffu, v, d);
        if (lane + d < r.run_end) v += t;
    }
    return v;
}

// (Measured and rejected, profiles/r02_experiments.md: bounding the number of shuffle steps by the warp's longest run --
// __reduce_max_sync + a warp-uniform early exit -- made every gradient kernel 12-30 % slower than the fixed five steps.)

// ------------------------------------------------------------------------------------------------
// Block-cooperative staging copies (all streaming kernels run DR_THREADS threads per CTA).

#define DR_THREADS 256

// Cooperative copy of n contiguous floats shared -> global.  Uses 128-bit stores on the 16-byte aligned body.
static __device__ __forceinline__ void coop_store(float* __restrict__ g, const float* __restrict__ s, int n, int tid)
{
    int head = (int)(((16 - ((uintptr_t)g & 15)) & 15) >> 2);
    if (head > n) head = n;
    if (tid < head) st_stream_f(g + tid, s[tid]);
    int body = (n - head) >> 2;
    float4* g4 = (float4*)(g + head);
    for (int i = tid; i < body; i += DR_THREADS)
    {
        const float* q = s + head + 4 * i;
        st_stream_f4(g4 + i, make_float4(q[0], q[1], q[2], q[3]));
    }
    int done = head + 4 * body;
    if (tid < n - done) st_stream_f(g + done + tid, s[done + tid]);
}

// Cooperative copy of n contiguous floats global -> shared (128-bit loads on the aligned body).
static __device__ __forceinline__ void coop_load(float* __restrict__ s, const float* __restrict__ g, int n, int tid)
{
    int head = (int)(((16 - ((uintptr_t)g & 15)) & 15) >> 2);
    if (head > n) head = n;
    if (tid < head) s[tid] = ld_stream_f(g + tid);
    int body = (n - head) >> 2;
    const float4* g4 = (const float4*)(g + head);
    for (int i = tid; i < body; i += DR_THREADS)
    {
        float4 v = ld_stream_f4(g4 + i);
        float* q = s + head + 4 * i;
        q[0] = v.x; q[1] = v.y; q[2] = v.z; q[3] = v.w;
    }
    int done = head + 4 * body;
    if (tid < n - done) s[done + tid] = ld_stream_f(g + done + tid);
}

// Runs capped at 4 lanes (an extra head every 4 lanes): two shuffle steps instead of five.  Used where a 2-4x cut in
// the number of atomics is what matters (duplicate texels of neighbouring pixels), not a full 32-lane reduction.
static __device__ __forceinline__ WarpRuns warp_runs_max4(int key, int key2)
{
    const int lane = threadIdx.x & 31;
    int prev = __shfl_up_sync(0xffffffffu, key, 1);
    int prev2 = __shfl_up_sync(0xffffffffu, key2, 1);
    bool head = ((lane & 3) == 0) || (prev != key) || (prev2 != key2);
    WarpRuns r;
    r.heads = __ballot_sync(0xffffffffu, head);
    unsigned above = (lane == 31) ? 0u : (r.heads >> (lane + 1));
    r.run_end = above ? (lane + __ffs(above)) : 32;
    r.is_head = head;
    return r;
}
static __device__ __forceinline__ float warp_run_sum_max4(float v, const WarpRuns& r)
{
    const int lane = threadIdx.x & 31;
    float t = __shfl_down_sync(0xffffffffu, v, 1);
    if (lane + 1 < r.run_end) v += t;
    t = __shfl_down_sync(0xffffffffu, v, 2);
    if (lane + 2 < r.run_end) v += t;
    return v;
}

// fp32 add without return value (RED).  Works on shared and global addresses (generic pointer: the compiler emits an address
// space query and both code paths -- use the two below where the space is known).
static __device__ __forceinline__ void red_add(float* p, float v) { atomicAdd(p, v); }
// global memory only: REDG.E.ADD.F32, fire and forget
static __device__ __forceinline__ void red_add_global(float* p, float v)
{
    asm volatile("red.global.add.f32 [%0], %1;" :: "l"(p), "f"(v) : "memory");
}
// shared memory only (fp32 adds to shared memory are compare-and-swap loops, ATOMS.CAST.SPIN; no address-space query here)
static __device__ __forceinline__ void red_add_shared(float* p, float v)
{
    asm volatile("red.shared.add.f32 [%0], %1;" :: "r"((uint32_t)__cvta_generic_to_shared(p)), "f"(v) : "memory");
}
// 128-bit vector reduction (global memory only, 16-byte aligned): REDG.E.ADD.F32x4 on sm_100a.
static __device__ __forceinline__ void red_add4(float* p, float4 v)
{
    asm volatile("red.global.v4.f32.add [%0], {%1,%2,%3,%4};" :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// 64-bit vector reduction (global memory only, 8-byte aligned): REDG.E.ADD.F32x2.
static __device__ __forceinline__ void red_add2(float* p, float2 v)
{
    asm volatile("red.global.v2.f32.add [%0], {%1,%2};" :: "l"(p), "f"(v.x), "f"(v.y) : "memory");
}

// three consecutive floats at a 4-byte aligned global address: one 64-bit vector reduction on the 8-byte aligned pair + one scalar
static __device__ __forceinline__ void red_add3(float* p, float a, float b, float c)
{
    if ((((uintptr_t)p) & 7) == 0) { red_add2(p, make_float2(a, b)); red_add(p + 2, c); }
    else                           { red_add(p, a); red_add2(p + 1, make_float2(b, c)); }
}

// ---- L2 prefetch of a warp's NEXT pixel group: measured and removed from every short-iteration kernel (profiles/r02_experiments.md
// section 2: the bulk form cp.async.bulk.prefetch.L2 goes through the TMA unit and slows the texture kernels by 3-26 %, the per-lane
// prefetch.global.L2 costs the texture forward 18 % and gains <= 2 % elsewhere).  Only the fused chain backward, whose iterations
// run for hundreds of instructions, keeps a per-lane prefetch (texture.cu, texture_chain_grad_fast).

// ---- TMA bulk copy (1-D, global -> shared) with an mbarrier transaction count: cp.async.bulk / UBLKCP on sm_100a.
// dst, src 16-byte aligned, bytes a multiple of 16.  One thread issues; every consumer waits on the barrier's phase.
static __device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
static __device__ __forceinline__ void mbar_init(unsigned long long* bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
static __device__ __forceinline__ void tma_load_1d(void* dst, const void* src, uint32_t bytes, unsigned long long* bar)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// several copies under one transaction count: mbar_expect_tx(total bytes) once, then tma_copy_1d per piece
static __device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
static __device__ __forceinline__ void tma_copy_1d(void* dst, const void* src, uint32_t bytes, unsigned long long* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
static __device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity)
{
    uint32_t ok;
    do
    {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!ok);
}

#endif // __CUDACC__
