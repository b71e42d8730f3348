/* Plain-C use of the drop-in boundary (no torch, no C++): rasterize one triangle at 64x64 through libb200dr.so and
 * print the covered-pixel count.  Build on a machine with the CUDA runtime:
 *
 *   gcc -std=c99 -Iinclude examples/c_abi_triangle.c -o c_abi_triangle \
 *       -Lnvdiffrast_b200 -lb200dr -L/usr/local/cuda/lib64 -lcudart -I/usr/local/cuda/include \
 *       -Wl,-rpath,$PWD/nvdiffrast_b200
 *
 * The same calls are what a binding in any host language makes: device pointers, sizes, a stream.
 */
#include <stdio.h>
#include <stdlib.h>
#include <cuda_runtime_api.h>
#include "b200dr.h"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA: %s\n", cudaGetErrorString(e_)); return 1; } } while (0)

int main(void)
{
    const int N = 1, V = 3, T = 1, H = 64, W = 64;
    const float pos[12] = { -0.8f, -0.8f, 0.f, 1.f,   0.8f, -0.8f, 0.f, 1.f,   0.f, 0.8f, 0.f, 1.f };   /* samples/torch/triangle.py */
    const int tri[3] = { 0, 1, 2 };
    float *d_pos, *d_rast, *d_db; int *d_tri; void* d_scratch;
    size_t px = (size_t)N * H * W;
    int path = b200dr_rasterize_pick_path(N, T, H, W);
    size_t sbytes = b200dr_rasterize_scratch_bytes(N, T, H, W, path);

    CK(cudaMalloc((void**)&d_pos, sizeof(pos)));   CK(cudaMemcpy(d_pos, pos, sizeof(pos), cudaMemcpyHostToDevice));
    CK(cudaMalloc((void**)&d_tri, sizeof(tri)));   CK(cudaMemcpy(d_tri, tri, sizeof(tri), cudaMemcpyHostToDevice));
    CK(cudaMalloc((void**)&d_rast, px * 16));      CK(cudaMalloc((void**)&d_db, px * 16));
    CK(cudaMalloc(&d_scratch, sbytes));

    if (b200dr_rasterize_fwd(d_pos, d_tri, NULL, N, V, T, H, W, /*instance_mode=*/1, /*max_tris=*/T, NULL, NULL,
                             d_scratch, sbytes, path, d_rast, d_db, /*stream=*/NULL))
    {
        fprintf(stderr, "b200dr: %s\n", b200dr_last_error());
        return 1;
    }
    float* rast = (float*)malloc(px * 16);
    CK(cudaMemcpy(rast, d_rast, px * 16, cudaMemcpyDeviceToHost));
    size_t covered = 0;
    for (size_t i = 0; i < px; i++) covered += rast[4 * i + 3] > 0.f;
    printf("libb200dr v%d: %zu of %zu pixels covered, %llu kernels launched\n", b200dr_version(), covered, px, b200dr_launch_count());
    free(rast);
    cudaFree(d_pos); cudaFree(d_tri); cudaFree(d_rast); cudaFree(d_db); cudaFree(d_scratch);
    return 0;
}
