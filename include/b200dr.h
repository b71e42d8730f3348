/*
 * b200dr.h -- C ABI of libb200dr.so, the B200-native (sm_100a) differentiable rasterizer hot path.
 *
 * This is the drop-in boundary for the reference's pybind module `_nvdiffrast_c`
 * (csrc/torch/torch_bindings.cpp:43-71): one C entry point per native op the reference's
 * nvdiffrast/torch/ops.py calls.  Plain device pointers, ints and a CUDA stream handle; no torch
 * types.  Ownership: every input/output buffer is allocated by the caller (the PyTorch front end,
 * nvdiffrast_b200/torch/ops.py, so the caching allocator and stream semantics are torch's); the
 * library never allocates device memory, never synchronises the stream and never touches the host
 * except to launch kernels.  All tensors are contiguous, row-major, fp32 / int32, 16-byte aligned.
 *
 * Return value: 0 on success, non-zero on error; b200dr_last_error() returns a thread-local
 * message in the style of the reference's NVDR_CHECK text (csrc/common/framework.h:24-25).
 */
#ifndef B200DR_H
#define B200DR_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200DR_VERSION 1

/* Filter / boundary enums: csrc/common/texture.h:26-36, nvdiffrast/torch/ops.py:415-420. */
#define B200DR_FILTER_NEAREST                 0
#define B200DR_FILTER_LINEAR                  1
#define B200DR_FILTER_LINEAR_MIPMAP_NEAREST   2
#define B200DR_FILTER_LINEAR_MIPMAP_LINEAR    3
#define B200DR_BOUNDARY_CUBE                  0
#define B200DR_BOUNDARY_WRAP                  1
#define B200DR_BOUNDARY_CLAMP                 2
#define B200DR_BOUNDARY_ZERO                  3
#define B200DR_TEX_MAX_MIP_LEVEL              16

/* Rasterizer path selection (both produce identical results). */
#define B200DR_RAST_AUTO     0
#define B200DR_RAST_BINS     1   /* per-bin scan + register depth/id keys + fused pixel shader   */
#define B200DR_RAST_SCATTER  2   /* triangle-parallel 64-bit atomicMin key buffer + resolve pass */

/* ---- misc ------------------------------------------------------------------------------ */
int         b200dr_version(void);
const char* b200dr_last_error(void);
/* replaces get_log_level / set_log_level (torch_bindings.cpp:50-51) */
int         b200dr_get_log_level(void);
void        b200dr_set_log_level(int level);
/* number of CUDA kernels this library has launched so far in this process */
unsigned long long b200dr_launch_count(void);

/* ---- rasterize ------------------------------------------------------------------------- */
/* Scratch bytes needed by b200dr_rasterize_fwd for the given problem (caller allocates). */
size_t b200dr_rasterize_scratch_bytes(int N, int max_tris_per_image, int H, int W, int path);
/* Resolved path (BINS or SCATTER) that AUTO would pick. */
int    b200dr_rasterize_pick_path(int N, int max_tris_per_image, int H, int W);

/*
 * Replaces rasterize_fwd_cuda (csrc/torch/torch_rasterize.cpp:43-166) including the whole
 * CudaRaster pipeline (csrc/common/cudaraster/impl/RasterImpl.cpp:250-368) and the pixel shader
 * (csrc/common/rasterize.cu:15-114).
 *   pos            [N,V,4] (instance_mode=1) or [V,4] (range mode)
 *   tri            [T,3]
 *   ranges_dev     range mode only: DEVICE copy of the [N,2] (start,count) CPU tensor, else NULL
 *   max_tris       max triangles drawn per image (T in instance mode, max count in range mode)
 *   peel_depth_in  NULL, or previous layer's depth [N,H,W] (fragments with depth <= it are killed;
 *                  impl/FineRaster.inl:349)
 *   depth_out      NULL, or receives this layer's depth [N,H,W] (needed only by DepthPeeler)
 *   scratch        b200dr_rasterize_scratch_bytes() bytes, 16-byte aligned, need not be initialised
 *   rast, rast_db  [N,H,W,4] outputs
 */
int b200dr_rasterize_fwd(const float* pos, const int32_t* tri, const int32_t* ranges_dev,
                         int N, int V, int T, int H, int W, int instance_mode, int max_tris,
                         const uint32_t* peel_depth_in, uint32_t* depth_out,
                         void* scratch, size_t scratch_bytes, int path,
                         float* rast, float* rast_db, void* stream);

/* Replaces rasterize_grad / rasterize_grad_db (torch_rasterize.cpp:171-262, rasterize.cu:119-277).
 * ddb may be NULL.  grad_pos has the shape of pos and MUST be zero-filled by the caller. */
int b200dr_rasterize_grad(const float* pos, const int32_t* tri, const float* rast, const float* dy, const float* ddb,
                          int N, int V, int T, int H, int W, int instance_mode,
                          float* grad_pos, void* stream);

/* ---- interpolate ----------------------------------------------------------------------- */
/* Replaces interpolate_fwd / interpolate_fwd_da (torch_interpolate.cpp:42-132, interpolate.cu:15-126).
 *   attr [V,A] | [N,V,A] | [1,V,A] (attr_bc=1); n_diff = number of differentiated attributes
 *   (A when diff_all), diff_list = host array of n_diff indices (ignored when diff_all);
 *   rast_db / out_da may be NULL when n_diff == 0.  out [N,H,W,A], out_da [N,H,W,2*n_diff]. */
int b200dr_interpolate_fwd(const float* attr, const float* rast, const int32_t* tri, const float* rast_db,
                           int N, int H, int W, int V, int T, int A, int attr_bc, int instance_mode,
                           int diff_all, const int32_t* diff_list, int n_diff,
                           float* out, float* out_da, void* stream);

/* Replaces interpolate_grad / interpolate_grad_da (torch_interpolate.cpp:137-248, interpolate.cu:131-274).
 * g_attr (shape of attr) MUST be zero-filled by the caller; g_rast, g_rast_db [N,H,W,4] are fully written. */
int b200dr_interpolate_grad(const float* attr, const float* rast, const int32_t* tri, const float* dy,
                            const float* rast_db, const float* dda,
                            int N, int H, int W, int V, int T, int A, int attr_bc, int instance_mode,
                            int diff_all, const int32_t* diff_list, int n_diff,
                            float* g_attr, float* g_rast, float* g_rast_db, void* stream);

/* ---- antialias ------------------------------------------------------------------------- */
/* Replaces antialias_construct_topology_hash (torch_antialias.cpp:25-63, antialias.cu:47-160).
 * Our topology object is a per-triangle table of the three opposite ("wing") vertices:
 * wings [T,4] int32 = (op0, op1, op2, 0), -1 where an edge has no second triangle.
 *   scratch: b200dr_topology_scratch_bytes(T) bytes, need not be initialised. */
size_t b200dr_topology_scratch_bytes(int T);
int    b200dr_antialias_construct_topology(const int32_t* tri, int T, int V, int32_t* wings,
                                           void* scratch, size_t scratch_bytes, void* stream);

/* Replaces antialias_fwd (torch_antialias.cpp:68-155, antialias.cu:165-382).
 *   out [N,H,W,C] is fully written (copy of color + blends)
 *   work_items: int4 x (2*N*H*W) capacity; work_counters: int32[4], zeroed here. */
int b200dr_antialias_fwd(const float* color, const float* rast, const float* pos, const int32_t* tri, const int32_t* wings,
                         int N, int H, int W, int C, int V, int T, int instance_mode,
                         float* out, void* work_items, int32_t* work_counters, void* stream);

/* Replaces antialias_grad (torch_antialias.cpp:160-241, antialias.cu:387-556).
 *   g_color [N,H,W,C] fully written (copy of dy + blends); g_pos (shape of pos) MUST be zero-filled by the caller. */
int b200dr_antialias_grad(const float* color, const float* rast, const float* pos, const int32_t* tri, const float* dy,
                          const void* work_items, const int32_t* work_counters,
                          int N, int H, int W, int C, int V, int T, int instance_mode,
                          float* g_color, float* g_pos, void* stream);

/* ---- texture (2-D and cube maps) -------------------------------------------------------- */
/* Cube maps (boundary_mode B200DR_BOUNDARY_CUBE, texture_kernel.cu:31-317): tex is [texD,6,texW,texW,Cp], every mip level
 * likewise; the six faces of a level are separate 2-D slices, so b200dr_texture_mip_info / _construct_mip /
 * _grad_scratch_bytes / _pad are called with depth 6*texD, while b200dr_texture_fwd / _grad take texD itself (the
 * minibatch), uv [N,H,W,3] direction vectors, uv_da / g_uv_da [N,H,W,6] (8-byte aligned).  At most 4 channels. */
/* Texels are consumed from a CHANNEL-PADDED pyramid: every level (0 = base texture .. max_level) holds
 * Cp = (C + 3) & ~3 floats per texel, so a bilinear tap is one 128-bit load per corner.  The caller pads level 0 with
 * b200dr_texture_pad (when C % 4 != 0) and builds or pads the mip levels; all level pointers are 16-byte aligned. */

/* Mip layout identical to calculateMipInfo (csrc/common/texture.cpp:62-102) for a texture with C floats per texel
 * (pass Cp to lay out the padded pyramid).  Returns the number of floats in the packed mip stack, fills
 * offsets[0..16] and *max_level; returns -1 if an odd extent cannot be halved. */
long long b200dr_texture_mip_info(int texW, int texH, int texD, int C, int max_mip_level, int* offsets, int* max_level);

/* dst[t][0..Cp) = src[t][0..C), zero padded; ntexels = texD*texH*texW of that level. */
int b200dr_texture_pad(const float* src, float* dst, long long ntexels, int C, void* stream);

/* Replaces texture_construct_mip (torch_texture.cpp:98-169, texture_kernel.cu:644-704) on padded data:
 * tex = padded level 0, mip = padded packed stack laid out by b200dr_texture_mip_info(.., Cp, ..), C = Cp. */
int b200dr_texture_construct_mip(const float* tex, float* mip, int texW, int texH, int texD, int C,
                                 const int* offsets, int max_level, void* stream);

/* Replaces texture_fwd / texture_fwd_mip (torch_texture.cpp:174-416, texture_kernel.cu:709-800).
 *   tex = padded level 0; mip_ptrs[1..max_level] = padded mip levels (mip_ptrs[0] ignored); C = real channel count. */
int b200dr_texture_fwd(const float* tex, const float* const* mip_ptrs, int max_level,
                       const float* uv, const float* uv_da, const float* mip_level_bias,
                       int N, int H, int W, int texD, int texH, int texW, int C,
                       int filter_mode, int boundary_mode, float* out, void* stream);

/* Replaces texture_grad_{nearest,linear,linear_mipmap_nearest,linear_mipmap_linear}
 * (torch_texture.cpp:421-716, texture_kernel.cu:843-1154).
 *   scratch: b200dr_texture_grad_scratch_bytes() bytes (channel-padded gradient pyramid + a 256-byte bookkeeping tail, zeroed here).
 *   tex / mip_ptrs: padded levels as in b200dr_texture_fwd.  g_tex (UNPADDED, shape of the user's tex) is fully written.  pull_mips != 0 (internal mip stack): mip-level gradients are pulled
 *   down into g_tex and g_mip_ptrs is ignored; pull_mips == 0 with max_level > 0 (custom mip stack): level l's
 *   gradient is written, UNPADDED, to g_mip_ptrs[l] (1..max_level).
 *   g_uv [N,H,W,2] (NULL for nearest), g_uv_da [N,H,W,4] / g_bias [N,H,W] may be NULL. */
size_t b200dr_texture_grad_scratch_bytes(int texD, int texH, int texW, int C, int max_level);
int b200dr_texture_grad(const float* tex, const float* const* mip_ptrs, int max_level,
                        const float* uv, const float* uv_da, const float* mip_level_bias, const float* dy,
                        int N, int H, int W, int texD, int texH, int texW, int C,
                        int filter_mode, int boundary_mode,
                        float* g_tex, float* const* g_mip_ptrs, int pull_mips,
                        float* g_uv, float* g_uv_da, float* g_bias,
                        void* scratch, size_t scratch_bytes, void* stream);

/* EXTENSION (no counterpart in the reference's binding; SURVEY.md 8(f)-4 "kernel fusion behind the same API"): the backward of
 * texture(interpolate(attr, rasterize(pos))) in ONE pass.  Equivalent to texture_grad_* followed by interpolate_grad_da
 * (torch_interpolate.cpp:137-248) and, when pos != NULL, rasterize_grad_db (torch_rasterize.cpp:171-262), without the g_uv /
 * g_uv_da / g_rast / g_rast_db round trips through memory: uv and uv_da are recomputed from attr [Va,2] (or [N,Va,2] with
 * attr_per_image) through attr_tri, rast and rast_db.  Preconditions the caller guarantees: 2-D texture modes, C <= 4, no
 * mip_level_bias, internal mip stack; uv (and uv_da) of the forward call were exactly interpolate's outputs on these inputs
 * with diff_attrs = all.  Outputs (each may be NULL when that gradient is not needed): g_tex (unpadded, fully written),
 * g_attr (zero-initialised by the caller, accumulated), g_pos [N,V,4] (zero-initialised, accumulated; instanced mode only),
 * or -- pos == NULL -- g_rast / g_rast_db [N,H,W,4] (fully written).  grad_db: rasterize() was called with grad_db=True.
 * rast_db may be NULL for filter_mode linear.  scratch as in b200dr_texture_grad. */
int b200dr_texture_chain_grad(const float* tex, const float* const* mip_ptrs, int max_level, const float* dy,
                              const float* attr, const int32_t* attr_tri, int Va, int attr_per_image,
                              const float* rast, const float* rast_db,
                              const float* pos, const int32_t* pos_tri, int V, int grad_db,
                              int N, int H, int W, int T, int texD, int texH, int texW, int C,
                              int filter_mode, int boundary_mode,
                              float* g_tex, float* g_attr, float* g_rast, float* g_rast_db, float* g_pos,
                              void* scratch, size_t scratch_bytes, void* stream);

/* EXTENSION (same status as b200dr_texture_chain_grad): the backward of interpolate(attr, rasterize(pos)) in ONE pass.
 * Equivalent to interpolate_grad[_da] (torch_interpolate.cpp:137-248) followed by rasterize_grad[_db]
 * (torch_rasterize.cpp:171-262) without the g_rast / g_rast_db round trip: 28 instead of 76 bytes per pixel for three plain
 * attributes.  Preconditions the caller guarantees: rast (and rast_db) are rasterize's outputs for pos [N,Vp,4] / pos_tri [T,3]
 * (instanced mode); 1 <= A <= 4; differentials for all attributes (diff_all = 1: rast_db, dda required) or none (diff_all = 0).
 * attr / tri / attr_bc / instance_mode as in b200dr_interpolate_grad.  grad_db: rasterize() was called with grad_db=True.
 * g_attr (shape of attr) and g_pos [N,Vp,4] MUST be zero-filled by the caller; both are accumulated into. */
int b200dr_interpolate_chain_grad(const float* attr, const float* rast, const int32_t* tri, const float* dy,
                                  const float* rast_db, const float* dda,
                                  const float* pos, const int32_t* pos_tri, int Vp, int grad_db,
                                  int N, int H, int W, int V, int T, int A, int attr_bc, int instance_mode, int diff_all,
                                  float* g_attr, float* g_pos, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B200DR_H */
