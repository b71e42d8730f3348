// Micro-benchmark: global reduction (RED) throughput on B200 for the access patterns of the texture-gradient kernel.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
__device__ __forceinline__ void red4(float* p, float4 v) { asm volatile("red.global.v4.f32.add [%0], {%1,%2,%3,%4};" :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory"); }
__device__ __forceinline__ void red2(float* p, float2 v) { asm volatile("red.global.v2.f32.add [%0], {%1,%2};" :: "l"(p), "f"(v.x), "f"(v.y) : "memory"); }

// mode 0: scalar x3 (C=3 layout, 12 B/texel); 1: v4 (16 B/texel); 2: scalar x4 (16 B/texel); 3: v2 x2 (16B/texel)
// pattern: pixel p of a 2048-wide image maps to texel (x/mag, y/mag) of a texW-wide texture, 4 bilinear neighbours
template <int MODE>
__global__ void k(float* g, int texW, int W, float mag, long npx)
{
    long p = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= npx) return;
    int x = (int)(p % W), y = (int)((p / W) % W);
    int tx = (int)(x / mag) % texW, ty = (int)(y / mag) % texW;
    for (int j = 0; j < 2; j++) for (int i = 0; i < 2; i++)
    {
        int t = ((ty + j) % texW) * texW + (tx + i) % texW;
        if (MODE == 0) { atomicAdd(g + 3 * t, 1.f); atomicAdd(g + 3 * t + 1, 1.f); atomicAdd(g + 3 * t + 2, 1.f); }
        if (MODE == 1) red4(g + 4 * t, make_float4(1.f, 1.f, 1.f, 0.f));
        if (MODE == 2) { atomicAdd(g + 4 * t, 1.f); atomicAdd(g + 4 * t + 1, 1.f); atomicAdd(g + 4 * t + 2, 1.f); atomicAdd(g + 4 * t + 3, 1.f); }
        if (MODE == 3) { red2(g + 4 * t, make_float2(1.f, 1.f)); red2(g + 4 * t + 2, make_float2(1.f, 0.f)); }
    }
}
// random spread pattern
template <int MODE>
__global__ void kr(float* g, int ntex, long npx)
{
    long p = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= npx) return;
    unsigned h = (unsigned)p * 2654435761u; h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
    int t = h % ntex;
    if (MODE == 0) { atomicAdd(g + 3 * t, 1.f); atomicAdd(g + 3 * t + 1, 1.f); atomicAdd(g + 3 * t + 2, 1.f); }
    if (MODE == 1) red4(g + 4 * t, make_float4(1.f, 1.f, 1.f, 0.f));
}
int main()
{
    int texW = 1024, W = 2048; long npx = (long)W * W * 8;
    float* g; cudaMalloc(&g, (size_t)texW * texW * 4 * 4 + 1024); cudaMemset(g, 0, (size_t)texW * texW * 16);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    float mags[] = {0.5f, 1.f, 2.5f, 8.f};
    const char* names[] = {"scalar x3 (12B)", "v4 (16B)", "scalar x4 (16B)", "v2 x2 (16B)"};
    for (float mag : mags) for (int mode = 0; mode < 4; mode++)
    {
        float ms = 0;
        for (int rep = 0; rep < 2; rep++)
        {
            cudaEventRecord(a);
            int blocks = (int)((npx + 255) / 256);
            if (mode == 0) k<0><<<blocks, 256>>>(g, texW, W, mag, npx);
            if (mode == 1) k<1><<<blocks, 256>>>(g, texW, W, mag, npx);
            if (mode == 2) k<2><<<blocks, 256>>>(g, texW, W, mag, npx);
            if (mode == 3) k<3><<<blocks, 256>>>(g, texW, W, mag, npx);
            cudaEventRecord(b); cudaEventSynchronize(b); cudaEventElapsedTime(&ms, a, b);
        }
        printf("mag %.1f  %-16s %8.3f ms  (%.1f G texel-updates/s)\n", mag, names[mode], ms, npx * 4 / ms / 1e6);
    }
    for (int mode = 0; mode < 2; mode++)
    {
        float ms = 0;
        for (int rep = 0; rep < 2; rep++)
        {
            cudaEventRecord(a);
            int blocks = (int)((npx + 255) / 256);
            if (mode == 0) kr<0><<<blocks, 256>>>(g, texW * texW, npx); else kr<1><<<blocks, 256>>>(g, texW * texW, npx);
            cudaEventRecord(b); cudaEventSynchronize(b); cudaEventElapsedTime(&ms, a, b);
        }
        printf("random  %-16s %8.3f ms  (%.1f G texel-updates/s)\n", names[mode], ms, npx / ms / 1e6);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
