"""PyTorch front end of the B200-native rasterizer: the `nvdiffrast.torch` op surface on top of libb200dr.so.

Public names, argument meaning, defaults, return shapes and error behaviour follow the reference front end
(nvdiffrast/torch/ops.py:47-559); everything below the C ABI (include/b200dr.h) is new.  This module owns what the
reference's C++ binding layer did (csrc/torch/torch_*.cpp): input validation, output allocation through torch's caching
allocator, current-stream lookup.  There is no CPU path and no fallback: CPU tensors raise, a missing CUDA library raises.
"""
import ctypes
import os
import warnings

import numpy as np
import torch

from .. import _lib

__all__ = ['RasterizeCudaContext', 'get_log_level', 'set_log_level', 'rasterize', 'DepthPeeler', 'interpolate',
           'texture', 'texture_construct_mip', 'antialias', 'antialias_construct_topology_hash', 'RasterizeGLContext']

_FILTER_MODES = {'nearest': 0, 'linear': 1, 'linear-mipmap-nearest': 2, 'linear-mipmap-linear': 3}
_BOUNDARY_MODES = {'cube': 0, 'wrap': 1, 'clamp': 2, 'zero': 3}
_TEX_MAX_MIP_LEVEL = 16

# ----------------------------------------------------------------------------------------------
# helpers


def _ptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _stream(device):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _fail(fn, msg):
    raise RuntimeError("%s(): %s" % (fn, msg))


def _check_inputs(fn, f32=(), i32=(), contiguous=True):
    """Device / layout / dtype checks with the reference's wording (csrc/torch/torch_common.inl:20-28).
    f32, i32: lists of (name, tensor)."""
    named = [x for x in list(f32) + list(i32) if x[1] is not None]
    names = ", ".join(n for n, _ in named)
    dev = None
    for _, t in named:
        if not t.is_cuda or (dev is not None and t.device != dev):
            _fail(fn, "Inputs %s must reside on the same GPU device" % names)
        dev = t.device
    if contiguous:
        for _, t in named:
            if not t.is_contiguous():
                _fail(fn, "Inputs %s must be contiguous tensors" % names)
    for n, t in f32:
        if t is not None and t.dtype != torch.float32:
            _fail(fn, "Inputs %s must be float32 tensors" % ", ".join(k for k, v in f32 if v is not None))
    for n, t in i32:
        if t is not None and t.dtype != torch.int32:
            _fail(fn, "Inputs %s must be int32 tensors" % ", ".join(k for k, v in i32 if v is not None))
    return dev


def _require(cond, fn, msg):
    if not cond:
        _fail(fn, msg)


# ----------------------------------------------------------------------------------------------
# Log level (nvdiffrast/torch/ops.py:18-41).


# ----------------------------------------------------------------------------------------------
# Cross-op fusion behind the unchanged API.
#
# The op boundaries force the backward pass to write g_uv / g_uv_da / g_rast / g_rast_db to HBM only to read them straight
# back in the next op.  `rasterize()` and `interpolate()` therefore tag their output tensors with where they came from; when
# `texture()` is handed tensors that are provably the untouched outputs of interpolate(attr, rast, tri, rast_db, 'all') -- and
# those the untouched outputs of rasterize(pos, ...) -- it samples exactly as before but registers ONE backward node that
# carries the gradient from the colour all the way to (tex, attr, pos) in a single kernel (b200dr_texture_chain_grad).  It
# reads uv / uv_da as constants, so the interpolate / rasterize nodes receive no gradient from this path and are not even
# executed unless another consumer (antialias's own pos gradient is separate) needs them; autograd sums the paths as usual.
# Anything the proof does not cover (views, in-place edits, hooks or retain_grad on the intermediates, custom mip stacks,
# mip_level_bias, cube maps, more than 4 channels, attribute counts other than 2) takes the op-by-op path.
#
# The same proof one level down: `interpolate()` handed the untouched outputs of an instanced-mode rasterize(pos, ...) registers
# a backward node that goes from (dy, dda) to (g_attr, g_pos) in one kernel (b200dr_interpolate_chain_grad; up to 4 attributes,
# differentials for all of them or none) -- chains without a texture (BASELINE's c2 / c4) and chains whose texture op cannot be
# fused (cube maps, > 4 channels) skip the g_rast / g_rast_db round trip this way.

_fusion_enabled = os.environ.get("B200DR_FUSION", "1") not in ("0", "false", "off")


def set_fusion(enabled):
    """Enable / disable the fused backward nodes (texture -> interpolate -> rasterize, interpolate -> rasterize); default: enabled
    (env B200DR_FUSION=0 disables)."""
    global _fusion_enabled
    _fusion_enabled = bool(enabled)


def get_fusion():
    return _fusion_enabled


class _Provenance:
    """Attached to op outputs as `_b200dr_prov` (outputs reference the record, the record references only the op's inputs:
    no reference cycle).  `versions`: in-place version counters of the outputs when they were produced."""
    __slots__ = ("kind", "inputs", "versions", "meta")

    def __init__(self, kind, inputs, outputs, **meta):
        self.kind, self.inputs, self.meta = kind, inputs, meta
        self.versions = tuple(o._version for o in outputs)
        # Only outputs that take part in differentiation are tagged: under no_grad (or with no differentiable input) nothing can be
        # fused, and the record would merely extend the lifetime of the op's inputs to that of its outputs.
        if _fusion_enabled and torch.is_grad_enabled() and any(o.requires_grad for o in outputs):
            for o in outputs:
                o._b200dr_prov = self


def _pristine(t, prov, slot):
    """t is output `slot` of the op recorded in prov, unmodified, and nobody watches its gradient."""
    return (getattr(t, "_b200dr_prov", None) is prov and t._version == prov.versions[slot]
            and not t.retains_grad and not getattr(t, "_backward_hooks", None))


def get_log_level():
    '''Get current log level (0 = Info, 1 = Warning, 2 = Error, 3 = Fatal; default 1).'''
    return _lib.load().b200dr_get_log_level()


def set_log_level(level):
    '''Set log level: messages of this severity or higher are printed.'''
    _lib.load().b200dr_set_log_level(int(level))


# ----------------------------------------------------------------------------------------------
# Rasterizer context (nvdiffrast/torch/ops.py:47-68, csrc/torch/torch_types.h:15-23).


class RasterizeCudaContext:
    def __init__(self, device=None):
        '''Create a rasterizer context on `device` (torch.device, string or int; default: current CUDA device).

        The context owns the rasterizer's scratch memory (allocated through torch, grown on demand) and, during depth
        peeling, the two depth layers.  One context per device; not thread-safe (same as the reference).'''
        _lib.load()
        if device is None:
            idx = torch.cuda.current_device()
        else:
            with torch.cuda.device(device):
                idx = torch.cuda.current_device()
        self.cuda_device_idx = idx
        self.device = torch.device('cuda', idx)
        self.active_depth_peeler = None
        self._scratch = None
        self._depth = [None, None]
        self.raster_path = 0   # B200DR_RAST_AUTO; tests may force 1 (bins) or 2 (scatter)

    def _get_scratch(self, nbytes):
        if self._scratch is None or self._scratch.numel() < nbytes:
            self._scratch = None
            self._scratch = torch.empty(int(nbytes), dtype=torch.uint8, device=self.device)
            if get_log_level() <= 0:
                print("nvdiffrast_b200: rasterizer scratch grown to %.1f MB" % (nbytes / 2**20))
        return self._scratch

    def _get_depth(self, slot, shape):
        d = self._depth[slot]
        if d is None or tuple(d.shape) != tuple(shape):
            d = torch.empty(shape, dtype=torch.int32, device=self.device)
            self._depth[slot] = d
        return d


def _rasterize_fwd(raster_ctx, pos, tri, resolution, ranges, peeling_idx):
    """Host side of rasterize_fwd_cuda (csrc/torch/torch_rasterize.cpp:43-166)."""
    fn = "rasterize_fwd_cuda"
    lib = _lib.load()
    dev = _check_inputs(fn, f32=[("pos", pos)], i32=[("tri", tri)])
    if ranges.is_cuda:
        _fail(fn, "Inputs ranges must reside on CPU")
    _require(ranges.is_contiguous(), fn, "Inputs pos, tri, ranges must be contiguous tensors")
    _require(ranges.dtype == torch.int32, fn, "Inputs tri, ranges must be int32 tensors")
    _require(dev.index == raster_ctx.cuda_device_idx, fn, "CudaRaster context must must reside on the same device as input tensors")

    instance_mode = pos.dim() > 2
    if instance_mode:
        _require(pos.dim() == 3 and pos.shape[0] > 0 and pos.shape[1] > 0 and pos.shape[2] == 4, fn, "instance mode - pos must have shape [>0, >0, 4]")
    else:
        _require(pos.dim() == 2 and pos.shape[0] > 0 and pos.shape[1] == 4, fn, "range mode - pos must have shape [>0, 4]")
        _require(ranges.dim() == 2 and ranges.shape[0] > 0 and ranges.shape[1] == 2, fn, "range mode - ranges must have shape [>0, 2]")
    _require(tri.dim() == 2 and tri.shape[0] > 0 and tri.shape[1] == 3, fn, "tri must have shape [>0, 3]")
    H, W = int(resolution[0]), int(resolution[1])
    _require(H > 0 and W > 0, fn, "resolution must be [>0, >0]")

    T = tri.shape[0]
    if instance_mode:
        N, V = pos.shape[0], pos.shape[1]
        max_tris = T
        ranges_dev = None
    else:
        N, V = ranges.shape[0], pos.shape[0]
        max_tris = max(int(ranges[:, 1].max().item()), 1)
        ranges_dev = ranges.to(dev, non_blocking=False)

    with torch.cuda.device(dev):
        path = raster_ctx.raster_path or lib.b200dr_rasterize_pick_path(N, max_tris, H, W)
        scratch = raster_ctx._get_scratch(lib.b200dr_rasterize_scratch_bytes(N, max_tris, H, W, path))
        out = torch.empty((N, H, W, 4), dtype=torch.float32, device=dev)
        out_db = torch.empty((N, H, W, 4), dtype=torch.float32, device=dev)
        peel_in = depth_out = None
        if peeling_idx >= 0:        # DepthPeeler: keep this layer's depth; layers > 0 test against the previous one
            depth_out = raster_ctx._get_depth(peeling_idx & 1, (N, H, W))
            if peeling_idx > 0:
                peel_in = raster_ctx._get_depth((peeling_idx - 1) & 1, (N, H, W))
        _lib.call("b200dr_rasterize_fwd", _ptr(pos), _ptr(tri), _ptr(ranges_dev), N, V, T, H, W, int(instance_mode), max_tris,
                                            _ptr(peel_in), _ptr(depth_out), _ptr(scratch), scratch.numel(), path,
                                            _ptr(out), _ptr(out_db), _stream(dev))
    return out, out_db


def _rasterize_grad(pos, tri, out, dy, ddb):
    """Host side of rasterize_grad / rasterize_grad_db (csrc/torch/torch_rasterize.cpp:171-262)."""
    fn = "rasterize_grad_db" if ddb is not None else "rasterize_grad"
    lib = _lib.load()
    dev = _check_inputs(fn, f32=[("pos", pos), ("out", out)], i32=[("tri", tri)])
    _check_inputs(fn, f32=[("pos", pos), ("dy", dy), ("ddb", ddb)], contiguous=False)
    instance_mode = pos.dim() > 2
    _require(out.dim() == 4, fn, "tensor out must be rank-4")
    N, H, W = out.shape[0], out.shape[1], out.shape[2]
    _require(N > 0 and H > 0 and W > 0, fn, "resolution must be [>0, >0, >0]")
    if instance_mode:
        _require(pos.dim() == 3 and pos.shape[0] == N and pos.shape[1] > 0 and pos.shape[2] == 4, fn, "pos must have shape [depth, >0, 4]")
    else:
        _require(pos.dim() == 2 and pos.shape[0] > 0 and pos.shape[1] == 4, fn, "pos must have shape [>0, 4]")
    _require(tri.dim() == 2 and tri.shape[0] > 0 and tri.shape[1] == 3, fn, "tri must have shape [>0, 3]")
    _require(tuple(out.shape) == (N, H, W, 4), fn, "out must have shape [depth, height, width, 4]")
    _require(tuple(dy.shape) == (N, H, W, 4), fn, "dy must have shape [depth, height, width, 4]")
    if ddb is not None:
        _require(tuple(ddb.shape) == (N, H, W, 4), fn, "ddb must have shape [depth, height, width, 4]")
        ddb = ddb.contiguous()
    dy = dy.contiguous()
    V = pos.shape[1] if instance_mode else pos.shape[0]
    with torch.cuda.device(dev):
        grad = torch.zeros_like(pos)
        _lib.call("b200dr_rasterize_grad", _ptr(pos), _ptr(tri), _ptr(out), _ptr(dy), _ptr(ddb), N, V, tri.shape[0], H, W,
                                             int(instance_mode), _ptr(grad), _stream(dev))
    return grad


class _rasterize_func(torch.autograd.Function):
    @staticmethod
    def forward(ctx, raster_ctx, pos, tri, resolution, ranges, grad_db, peeling_idx):
        out, out_db = _rasterize_fwd(raster_ctx, pos, tri, resolution, ranges, peeling_idx)
        ctx.save_for_backward(pos, tri, out)
        ctx.saved_grad_db = grad_db
        # Unused outputs arrive as None instead of materialised zero tensors: an unused rast_db costs no bandwidth.
        ctx.set_materialize_grads(False)
        return out, out_db

    @staticmethod
    def backward(ctx, dy, ddb):
        pos, tri, out = ctx.saved_tensors
        if dy is None and (ddb is None or not ctx.saved_grad_db):
            return None, torch.zeros_like(pos), None, None, None, None, None
        if dy is None:
            dy = torch.zeros_like(out)
        g_pos = _rasterize_grad(pos, tri, out, dy, ddb if ctx.saved_grad_db else None)
        return None, g_pos, None, None, None, None, None


def _cpu_ranges(ranges):
    """Instanced mode passes an empty [0, 2] int32 CPU tensor down to the op (csrc/torch/torch_rasterize.cpp:60-67)."""
    return torch.empty(size=(0, 2), dtype=torch.int32, device='cpu') if ranges is None else ranges


def rasterize(glctx, pos, tri, resolution, ranges=None, grad_db=True):
    '''Rasterize triangles.

    Args:
        glctx: `RasterizeCudaContext`.
        pos: float32 clip-space vertex positions, [num_vertices, 4] (range mode) or
             [minibatch_size, num_vertices, 4] (instanced mode), on the GPU.
        tri: int32 [num_triangles, 3], on the GPU.
        resolution: (height, width).
        ranges: range mode only: int32 CPU tensor [minibatch_size, 2] of (start, count) into `tri`.
        grad_db: propagate gradients of the barycentric image-space derivatives into `pos`.

    Returns:
        (rast, rast_db), both [minibatch_size, height, width, 4]: (u, v, z/w, triangle_id + 1) and
        (du/dX, du/dY, dv/dX, dv/dY).  Empty pixels are all zeros.
    '''
    assert isinstance(glctx, RasterizeCudaContext)
    assert grad_db is True or grad_db is False
    assert isinstance(pos, torch.Tensor) and isinstance(tri, torch.Tensor)
    assert ranges is None or isinstance(ranges, torch.Tensor)
    resolution, ranges = tuple(resolution), _cpu_ranges(ranges)
    if glctx.active_depth_peeler is not None:
        # The reference *returns* this error object (nvdiffrast/torch/ops.py:131-132); kept for drop-in behaviour.
        return RuntimeError("Cannot call rasterize() during depth peeling operation, use rasterize_next_layer() instead")
    rast, rast_db = _rasterize_func.apply(glctx, pos, tri, resolution, ranges, grad_db, -1)
    _Provenance("rasterize", (pos, tri), (rast, rast_db), grad_db=grad_db)
    return rast, rast_db


class DepthPeeler:
    """Depth peeling: `with DepthPeeler(glctx, pos, tri, resolution) as peeler: rast, db = peeler.rasterize_next_layer()`.
    Every call returns the next-nearest surface layer (arguments as in `rasterize()`; same protocol and error texts as
    nvdiffrast/torch/ops.py:141-204).  The context keeps two ping-pong depth layers while a peeler is active: layer k kills
    fragments at or in front of layer k-1's depth."""

    def __init__(self, glctx, pos, tri, resolution, ranges=None, grad_db=True):
        assert isinstance(glctx, RasterizeCudaContext)
        assert grad_db is True or grad_db is False
        assert isinstance(pos, torch.Tensor) and isinstance(tri, torch.Tensor)
        assert ranges is None or isinstance(ranges, torch.Tensor)
        self.raster_ctx = glctx
        # everything rasterize_next_layer() forwards, frozen at construction
        self._args = (pos, tri, tuple(resolution), _cpu_ranges(ranges), grad_db)
        self._layer = None                              # None: not entered; k: next layer to produce

    # read-only views of the frozen arguments (the reference exposes them as attributes)
    pos = property(lambda self: self._args[0] if self._args else None)
    tri = property(lambda self: self._args[1] if self._args else None)
    resolution = property(lambda self: self._args[2] if self._args else None)
    ranges = property(lambda self: self._args[3] if self._args else None)
    grad_db = property(lambda self: self._args[4] if self._args else None)
    peeling_idx = property(lambda self: self._layer)

    def __enter__(self):
        ctx = self.raster_ctx
        if ctx is None:
            raise RuntimeError("Cannot re-enter a terminated depth peeling operation")
        if ctx.active_depth_peeler is not None:
            raise RuntimeError("Cannot have multiple depth peelers active simultaneously in a rasterization context")
        ctx.active_depth_peeler = self
        self._layer = 0
        return self

    def __exit__(self, *exc):
        assert self.raster_ctx.active_depth_peeler is self
        self.raster_ctx.active_depth_peeler = None
        self.raster_ctx, self._args, self._layer = None, None, None      # terminated: cannot be re-entered
        return None

    def rasterize_next_layer(self):
        '''Rasterize the next depth layer: as `rasterize()`, minus the surface points already reported.'''
        assert self.raster_ctx.active_depth_peeler is self
        assert self._layer >= 0
        pos, tri, resolution, ranges, grad_db = self._args
        result = _rasterize_func.apply(self.raster_ctx, pos, tri, resolution, ranges, grad_db, self._layer)
        _Provenance("rasterize", (pos, tri), result, grad_db=grad_db)
        self._layer += 1
        return result


# ----------------------------------------------------------------------------------------------
# Interpolate (nvdiffrast/torch/ops.py:211-291, csrc/torch/torch_interpolate.cpp).


def _interp_dims(fn, attr, rast, tri, rast_db):
    _require(rast.dim() == 4 and min(rast.shape[:3]) > 0 and rast.shape[3] == 4, fn, "rast must have shape[>0, >0, >0, 4]")
    _require(tri.dim() == 2 and tri.shape[0] > 0 and tri.shape[1] == 3, fn, "tri must have shape [>0, 3]")
    _require(attr.dim() in (2, 3) and min(attr.shape) > 0, fn, "attr must have shape [>0, >0, >0] or [>0, >0]")
    instance_mode = attr.dim() > 2
    if instance_mode:
        _require(attr.shape[0] == rast.shape[0] or attr.shape[0] == 1, fn, "minibatch size mismatch between inputs rast, attr")
    if rast_db is not None:
        _require(rast_db.dim() == 4 and min(rast_db.shape[:3]) > 0 and rast_db.shape[3] == 4, fn, "rast_db must have shape[>0, >0, >0, 4]")
        _require(rast_db.shape[1] == rast.shape[1] and rast_db.shape[2] == rast.shape[2], fn, "spatial size mismatch between inputs rast and rast_db")
        _require(rast_db.shape[0] == rast.shape[0], fn, "minibatch size mismatch between inputs rast, rast_db")
    N, H, W = rast.shape[0], rast.shape[1], rast.shape[2]
    V, A = (attr.shape[1], attr.shape[2]) if instance_mode else (attr.shape[0], attr.shape[1])
    attr_bc = int(instance_mode and attr.shape[0] == 1 and N > 1)
    return N, H, W, V, A, int(instance_mode), attr_bc


def _diff_args(diff_attrs_all, diff_attrs_list, A):
    if diff_attrs_all:
        return 1, None, A
    arr = (ctypes.c_int32 * max(len(diff_attrs_list), 1))(*diff_attrs_list)
    return 0, arr, len(diff_attrs_list)


def _interpolate_fwd(attr, rast, tri, rast_db, diff_attrs_all, diff_attrs_list):
    da = rast_db is not None
    fn = "interpolate_fwd_da" if da else "interpolate_fwd"
    lib = _lib.load()
    dev = _check_inputs(fn, f32=[("attr", attr), ("rast", rast), ("rast_db", rast_db)], i32=[("tri", tri)])
    N, H, W, V, A, inst, bc = _interp_dims(fn, attr, rast, tri, rast_db)
    dall, dlist, D = _diff_args(diff_attrs_all, diff_attrs_list, A) if da else (0, None, 0)
    if da and not dall:
        _require(D <= 32, fn, "too many entries in diff_attrs list (increase IP_MAX_DIFF_ATTRS)")
    with torch.cuda.device(dev):
        out = torch.empty((N, H, W, A), dtype=torch.float32, device=dev)
        out_da = torch.empty((N, H, W, 2 * D), dtype=torch.float32, device=dev)
        _lib.call("b200dr_interpolate_fwd", _ptr(attr), _ptr(rast), _ptr(tri), _ptr(rast_db), N, H, W, V, tri.shape[0], A, bc, inst,
                                              dall, dlist, D, _ptr(out), _ptr(out_da) if D else None, _stream(dev))
    return out, out_da


def _interpolate_grad(attr, rast, tri, dy, rast_db, dda, diff_attrs_all, diff_attrs_list):
    da = rast_db is not None
    fn = "interpolate_grad_da" if da else "interpolate_grad"
    lib = _lib.load()
    dev = _check_inputs(fn, f32=[("attr", attr), ("rast", rast), ("rast_db", rast_db)], i32=[("tri", tri)])
    _check_inputs(fn, f32=[("attr", attr), ("dy", dy), ("dda", dda)], contiguous=False)
    N, H, W, V, A, inst, bc = _interp_dims(fn, attr, rast, tri, rast_db)
    dall, dlist, D = _diff_args(diff_attrs_all, diff_attrs_list, A) if da else (0, None, 0)
    _require(tuple(dy.shape) == (N, H, W, A), fn, "dy must have shape [minibatch_size, height, width, num_attributes]")
    dy = dy.contiguous()
    if da:
        _require(tuple(dda.shape) == (N, H, W, 2 * D), fn, "dda must have shape [minibatch_size, height, width, 2 * num_diff_attributes]")
        dda = dda.contiguous()
    with torch.cuda.device(dev):
        g_attr = torch.zeros_like(attr)
        g_rast = torch.empty_like(rast)
        g_rast_db = torch.empty_like(rast_db) if da else None
        _lib.call("b200dr_interpolate_grad", _ptr(attr), _ptr(rast), _ptr(tri), _ptr(dy), _ptr(rast_db), _ptr(dda), N, H, W, V, tri.shape[0], A, bc, inst,
                                               dall, dlist, D, _ptr(g_attr), _ptr(g_rast), _ptr(g_rast_db), _stream(dev))
    return g_attr, g_rast, g_rast_db


class _interpolate_func_da(torch.autograd.Function):
    @staticmethod
    def forward(ctx, attr, rast, tri, rast_db, diff_attrs_all, diff_attrs_list):
        out, out_da = _interpolate_fwd(attr, rast, tri, rast_db, diff_attrs_all, diff_attrs_list)
        ctx.save_for_backward(attr, rast, tri, rast_db)
        ctx.saved_misc = diff_attrs_all, diff_attrs_list
        return out, out_da

    @staticmethod
    def backward(ctx, dy, dda):
        attr, rast, tri, rast_db = ctx.saved_tensors
        diff_attrs_all, diff_attrs_list = ctx.saved_misc
        g_attr, g_rast, g_rast_db = _interpolate_grad(attr, rast, tri, dy, rast_db, dda, diff_attrs_all, diff_attrs_list)
        return g_attr, g_rast, None, g_rast_db, None, None


class _interpolate_func(torch.autograd.Function):
    @staticmethod
    def forward(ctx, attr, rast, tri):
        out, out_da = _interpolate_fwd(attr, rast, tri, None, 0, [])
        ctx.save_for_backward(attr, rast, tri)
        return out, out_da

    @staticmethod
    def backward(ctx, dy, _):
        attr, rast, tri = ctx.saved_tensors
        g_attr, g_rast, _unused = _interpolate_grad(attr, rast, tri, dy, None, None, 0, [])
        return g_attr, g_rast, None


def _aligned(t, nbytes):
    """t, or a fresh copy when its address is not a multiple of nbytes (views with a storage offset)."""
    return t if t is None or t.data_ptr() % nbytes == 0 else t.clone(memory_format=torch.contiguous_format)


def _interpolate_chain_grad(attr, rast, tri, dy, rast_db, dda, pos, pos_tri, grad_db):
    """Host side of b200dr_interpolate_chain_grad (include/b200dr.h): -> (g_attr, g_pos)."""
    da = rast_db is not None
    fn = "interpolate_chain_grad"
    _lib.load()
    dev = _check_inputs(fn, f32=[("attr", attr), ("rast", rast), ("rast_db", rast_db), ("pos", pos)], i32=[("tri", tri), ("pos_tri", pos_tri)])
    _check_inputs(fn, f32=[("attr", attr), ("dy", dy), ("dda", dda if da else None)], contiguous=False)
    N, H, W, V, A, inst, bc = _interp_dims(fn, attr, rast, tri, rast_db)
    _require(pos.dim() == 3 and pos.shape[0] == N and pos.shape[2] == 4, fn, "pos must have shape [minibatch_size, num_vertices, 4]")
    _require(pos_tri.shape == tri.shape, fn, "attribute and position index buffers must describe the same triangles")
    _require(tuple(dy.shape) == (N, H, W, A), fn, "dy must have shape [minibatch_size, height, width, num_attributes]")
    dy = _aligned(dy.contiguous(), 16)
    if da:
        _require(tuple(dda.shape) == (N, H, W, 2 * A), fn, "dda must have shape [minibatch_size, height, width, 2 * num_attributes]")
        dda = _aligned(dda.contiguous(), 32)
    with torch.cuda.device(dev):
        g_attr = torch.zeros_like(attr)
        g_pos = torch.zeros_like(pos)
        _lib.call("b200dr_interpolate_chain_grad", _ptr(attr), _ptr(rast), _ptr(tri), _ptr(dy), _ptr(rast_db), _ptr(dda) if da else None,
                  _ptr(pos), _ptr(pos_tri), pos.shape[1], int(grad_db), N, H, W, V, tri.shape[0], A, bc, inst, int(da),
                  _ptr(g_attr), _ptr(g_pos), _stream(dev))
    return g_attr, g_pos


class _interpolate_chain_func(torch.autograd.Function):
    """interpolate() whose backward also runs rasterize's (one kernel).  Differentiable inputs: attr and pos; rast / rast_db enter
    as constants (they are proven to be the untouched outputs of rasterize(pos, pos_tri))."""

    @staticmethod
    def forward(ctx, attr, pos, rast, rast_db, tri, pos_tri, grad_db):
        out, out_da = _interpolate_fwd(attr, rast, tri, rast_db, 1 if rast_db is not None else 0, [])
        ctx.save_for_backward(attr, pos, rast, tri, pos_tri, rast_db if rast_db is not None else torch.tensor([]))
        ctx.saved_misc = grad_db, rast_db is not None
        return out, out_da

    @staticmethod
    def backward(ctx, dy, dda):
        attr, pos, rast, tri, pos_tri, rast_db = ctx.saved_tensors
        grad_db, da = ctx.saved_misc
        g_attr, g_pos = _interpolate_chain_grad(attr, rast, tri, dy, rast_db if da else None, dda if da else None, pos, pos_tri, grad_db)
        return g_attr, g_pos, None, None, None, None, None


def _interp_fusion_plan(attr, rast, tri, rast_db, want_all, which):
    """-> (pos, pos_tri, grad_db) when interpolate's backward can carry on through rasterize's in one kernel, else None."""
    if not (_fusion_enabled and torch.is_grad_enabled()) or which:
        return None
    if not (attr.is_cuda and attr.dtype == torch.float32 and attr.is_contiguous() and attr.dim() in (2, 3) and 1 <= attr.shape[-1] <= 4
            and attr.data_ptr() % 16 == 0):
        return None
    rprov = getattr(rast, "_b200dr_prov", None)
    if rprov is None or rprov.kind != "rasterize" or not _pristine(rast, rprov, 0):
        return None
    pos, pos_tri = rprov.inputs
    if pos.dim() != 3 or not pos.requires_grad or pos.data_ptr() % 16 != 0 or not pos.is_contiguous():
        return None
    if not (isinstance(tri, torch.Tensor) and tri.shape == pos_tri.shape):
        return None
    if want_all and (rast_db is None or not _pristine(rast_db, rprov, 1)):
        return None
    return pos, pos_tri, bool(rprov.meta["grad_db"])


def _diff_attr_arg(diff_attrs):
    """`diff_attrs` of interpolate(): None | 'all' | 1-D sequence of attribute indices  ->  (all?, [indices])
    (same accepted forms as nvdiffrast/torch/ops.py:276-284)."""
    if isinstance(diff_attrs, str):
        assert diff_attrs == 'all', "diff_attrs: a list of attribute indices or 'all'"
        return True, []
    if diff_attrs is None:
        return False, []
    idx = np.asarray(diff_attrs, np.int32)
    assert idx.ndim == 1, "diff_attrs must be one-dimensional"
    return False, idx.tolist()


def _mip_limit(max_mip_level):
    """None -> -1 (no limit); otherwise a non-negative int."""
    if max_mip_level is None:
        return -1
    level = int(max_mip_level)
    assert level >= 0, "max_mip_level must be >= 0"
    return level


def interpolate(attr, rast, tri, rast_db=None, diff_attrs=None):
    """Interpolate vertex attributes.

    Args:
        attr: float32 [num_vertices, num_attributes] (range mode) or [minibatch_size, num_vertices, num_attributes]
              (instanced mode; a leading dimension of 1 broadcasts over the minibatch).
        rast: main output of `rasterize()`.
        tri: int32 [num_triangles, 3].
        rast_db: (optional) second output of `rasterize()`; enables image-space derivatives of attributes.
        diff_attrs: (optional) list of attribute indices to differentiate, or 'all'.

    Returns:
        (out, out_da): [minibatch_size, height, width, num_attributes] and
        [minibatch_size, height, width, 2 * len(diff_attrs)] ((dA/dX, dA/dY) pairs; last dimension 0 if unused).
    """
    want_all, which = _diff_attr_arg(diff_attrs)
    for name, t in (("attr", attr), ("rast", rast), ("tri", tri)):
        assert isinstance(t, torch.Tensor), "%s must be a tensor" % name
    if want_all or which:
        assert isinstance(rast_db, torch.Tensor), "diff_attrs needs rast_db"
    else:
        rast_db = None
    plan = _interp_fusion_plan(attr, rast, tri, rast_db, want_all, which)
    if plan is not None:
        # rast / rast_db enter as constants (detached): this node owns the path (out, out_da) -> (attr, pos)
        out = _interpolate_chain_func.apply(attr, plan[0], rast.detach(), rast_db.detach() if want_all else None, tri, plan[1], plan[2])
    elif want_all or which:
        out = _interpolate_func_da.apply(attr, rast, tri, rast_db, int(want_all), which)
    else:
        out = _interpolate_func.apply(attr, rast, tri)
    _Provenance("interpolate", (attr, rast, tri, rast_db), out, diff_all=want_all)
    return out


# ----------------------------------------------------------------------------------------------
# Texture (nvdiffrast/torch/ops.py:298-468, csrc/torch/torch_texture.cpp).


class TextureMipWrapper:
    """Opaque mip stack returned by `texture_construct_mip()` (csrc/torch/torch_types.h:28-35).
    `mip` holds mip levels 1..n packed back to back, each texel padded to a multiple of 4 channels."""

    def __init__(self, mip=None, max_mip_level=0, texture_size=(), cube_mode=False):
        self.mip = mip
        self.max_mip_level = max_mip_level
        self.texture_size = tuple(texture_size)
        self.cube_mode = cube_mode

    def level(self, i):
        """Unpadded view [minibatch, height, width, channels] of mip level i >= 1 (for inspection / tests)."""
        D, Ht, Wt, C = self.texture_size[0], self.texture_size[-3], self.texture_size[-2], self.texture_size[-1]
        Cp = (C + 3) & ~3
        depth = 6 * D if self.cube_mode else D
        _, offs, level_max = _mip_info("texture_construct_mip", (depth, Ht, Wt, Cp), self.max_mip_level)
        assert 1 <= i <= level_max
        w, h = max(Wt >> i, 1), max(Ht >> i, 1)
        lvl = self.mip[offs[i]:offs[i] + depth * h * w * Cp]
        return (lvl.view(D, 6, h, w, Cp) if self.cube_mode else lvl.view(D, h, w, Cp))[..., :C]


def _mip_info(fn, tex_shape, limit):
    """Mip layout of calculateMipInfo (csrc/common/texture.cpp:62-102) -> (total floats, offsets, level_max)."""
    lib = _lib.load()
    D, Ht, Wt, C = tex_shape
    offs = (ctypes.c_int * (_TEX_MAX_MIP_LEVEL + 1))()
    level_max = ctypes.c_int(0)
    total = lib.b200dr_texture_mip_info(Wt, Ht, D, C, int(limit), offs, ctypes.byref(level_max))
    if total < 0:
        _fail(fn, "Texture mipmap construction failed: a dimension of size >1 that is not divisible by 2 was encountered "
                  "(texture size %dx%d). Use max_mip_level to limit the number of mip levels" % (Wt, Ht))
    return int(total), offs, int(level_max.value)


def _tex_pad(t):
    """[..., C] -> [..., Cp] with Cp = C rounded up to a multiple of 4 (zero padded); `t` itself when C % 4 == 0."""
    C = t.shape[-1]
    Cp = (C + 3) & ~3
    if Cp == C:
        return t
    out = torch.empty(tuple(t.shape[:-1]) + (Cp,), dtype=torch.float32, device=t.device)
    _lib.call("b200dr_texture_pad", _ptr(t), _ptr(out), t.numel() // C, C, _stream(t.device))
    return out


def _texture_construct_mip(tex, max_mip_level, cube_mode):
    fn = "texture_construct_mip"
    _lib.load()
    _require(max_mip_level >= -1, fn, "invalid max_mip_level")
    dev = _check_inputs(fn, f32=[("tex", tex)])
    if cube_mode:
        _require(tex.dim() == 5 and min(tex.shape) > 0 and tex.shape[1] == 6, fn, "tex must have shape[>0, 6, >0, >0, >0] in cube map mode")
        _require(tex.shape[2] == tex.shape[3], fn, "texture shape must be square in cube map mode")
        D, Ht, Wt, C = 6 * tex.shape[0], tex.shape[2], tex.shape[3], tex.shape[4]     # faces are separate 2-D slices
    else:
        _require(tex.dim() == 4 and min(tex.shape) > 0, fn, "tex must have shape[>0, >0, >0, >0]")
        D, Ht, Wt, C = tex.shape
    Cp = (C + 3) & ~3
    total, offs, level_max = _mip_info(fn, (D, Ht, Wt, Cp), max_mip_level)
    with torch.cuda.device(dev):
        mip = torch.empty((total,), dtype=torch.float32, device=dev)
        if level_max > 0:
            lvl0 = _tex_pad(tex)
            _lib.call("b200dr_texture_construct_mip", _ptr(lvl0), _ptr(mip), Wt, Ht, D, Cp, offs, level_max, _stream(dev))
    return TextureMipWrapper(mip, max_mip_level, tuple(tex.shape), cube_mode)


def _tex_setup(fn, tex, uv, uv_da, mip_level_bias, mip_wrapper, mip_stack, filter_mode, boundary_mode):
    """Validation + padded level pointer table shared by forward and backward (csrc/torch/torch_texture.cpp:174-330).
    Returns the transient padded tensors too: they must stay referenced until the kernels are enqueued."""
    _require(0 <= filter_mode < 4, fn, "filter_mode unsupported")
    _require(0 <= boundary_mode < 4, fn, "boundary_mode unsupported")
    enable_mip = filter_mode in (2, 3)
    has_stack = len(mip_stack) > 0
    has_uv_da = uv_da is not None and uv_da.numel() > 0
    has_bias = mip_level_bias is not None and mip_level_bias.numel() > 0
    if enable_mip:
        _require(has_uv_da or has_bias, fn, "mipmapping filter mode requires uv_da and/or mip_level_bias input")
        _require(has_stack or (mip_wrapper is not None and mip_wrapper.mip is not None), fn, "mipmapping filter mode requires mip wrapper or mip stack input")
    dev = _check_inputs(fn, f32=[("tex", tex), ("uv", uv)])
    cube = (boundary_mode == 0)
    if cube:
        _require(tex.dim() == 5 and min(tex.shape) > 0 and tex.shape[1] == 6, fn, "tex must have shape[>0, 6, >0, >0, >0] in cube map mode")
        _require(uv.dim() == 4 and min(uv.shape[:3]) > 0 and uv.shape[3] == 3, fn, "uv must have shape [>0, >0, >0, 3] in cube map mode")
        _require(tex.shape[2] == tex.shape[3], fn, "texture shape must be square in cube map mode")
        D, Ht, Wt, C = tex.shape[0], tex.shape[2], tex.shape[3], tex.shape[4]
    else:
        _require(tex.dim() == 4 and min(tex.shape) > 0, fn, "tex must have shape[>0, >0, >0, >0]")
        _require(uv.dim() == 4 and min(uv.shape[:3]) > 0 and uv.shape[3] == 2, fn, "uv must have shape [>0, >0, >0, 2]")
        D, Ht, Wt, C = tex.shape
    Cp = (C + 3) & ~3
    lvl_shape = (lambda h, w: (D, 6, h, w, C)) if cube else (lambda h, w: (D, h, w, C))
    N, H, W = uv.shape[0], uv.shape[1], uv.shape[2]
    _require(D == 1 or D == N, fn, "minibatch size mismatch between inputs tex, uv")
    _require(Wt <= (1 << _TEX_MAX_MIP_LEVEL) and Ht <= (1 << _TEX_MAX_MIP_LEVEL), fn, "texture size too large")
    level_max = 0
    ptrs = (ctypes.c_void_p * (_TEX_MAX_MIP_LEVEL + 1))()
    with torch.cuda.device(dev):
        keep = [_tex_pad(tex)]
    if enable_mip:
        if has_uv_da:
            _check_inputs(fn, f32=[("tex", tex), ("uv_da", uv_da)])
            if cube:
                _require(tuple(uv_da.shape) == (N, H, W, 6), fn, "uv_da must have shape [minibatch_size, height, width, 6] in cube map mode")
            else:
                _require(tuple(uv_da.shape) == (N, H, W, 4), fn, "uv_da must have shape [minibatch_size, height, width, 4]")
        if has_bias:
            _check_inputs(fn, f32=[("tex", tex), ("mip_level_bias", mip_level_bias)])
            _require(tuple(mip_level_bias.shape) == (N, H, W), fn, "mip_level_bias must have shape [minibatch_size, height, width]")
        if has_stack:
            for t in mip_stack:
                if not t.is_cuda or t.device != dev:
                    _fail(fn, "Mip stack inputs must reside on the correct GPU device")
                if not t.is_contiguous():
                    _fail(fn, "Mip stack inputs must be contiguous tensors")
                if t.dtype != torch.float32:
                    _fail(fn, "Mip stack inputs must be float32 tensors")
            level_max = len(mip_stack)
            _require(level_max <= _TEX_MAX_MIP_LEVEL, fn, "too many mip levels")
            for i in range(1, level_max + 1):
                t = mip_stack[i - 1]
                sw, sh = max(Wt >> i, 1), max(Ht >> i, 1)
                _require(tuple(t.shape) == lvl_shape(sh, sw), fn, "mip level size mismatch in custom mip stack")
                if sw == 1 and sh == 1:
                    _require(i == level_max, fn, "mip level size mismatch in mip stack")
                with torch.cuda.device(dev):
                    keep.append(_tex_pad(t.detach()))
                ptrs[i] = keep[-1].data_ptr()
        else:
            total, offs, level_max = _mip_info(fn, (6 * D if cube else D, Ht, Wt, Cp), mip_wrapper.max_mip_level)
            _require(tuple(tex.shape) == tuple(mip_wrapper.texture_size) and cube == bool(mip_wrapper.cube_mode), fn, "mip does not match texture size")
            _check_inputs(fn, f32=[("tex", tex), ("mip", mip_wrapper.mip)])
            _require(mip_wrapper.mip.dim() == 1 and mip_wrapper.mip.shape[0] == total, fn, "wrapped mip tensor size mismatch")
            base = mip_wrapper.mip.data_ptr()
            for i in range(1, level_max + 1):
                ptrs[i] = base + 4 * offs[i]
    return dev, (N, H, W, D, Ht, Wt, C), enable_mip, has_uv_da, has_bias, level_max, ptrs, keep


def _texture_fwd(tex, uv, uv_da, mip_level_bias, mip_wrapper, mip_stack, filter_mode, boundary_mode):
    fn = "texture_fwd_mip" if filter_mode in (2, 3) else "texture_fwd"
    lib = _lib.load()
    dev, (N, H, W, D, Ht, Wt, C), enable_mip, has_uv_da, has_bias, level_max, ptrs, keep = \
        _tex_setup(fn, tex, uv, uv_da, mip_level_bias, mip_wrapper, mip_stack, filter_mode, boundary_mode)
    with torch.cuda.device(dev):
        out = torch.empty((N, H, W, C), dtype=torch.float32, device=dev)
        _lib.call("b200dr_texture_fwd", _ptr(keep[0]), ptrs, level_max, _ptr(uv), _ptr(uv_da) if (enable_mip and has_uv_da) else None,
                                          _ptr(mip_level_bias) if (enable_mip and has_bias) else None,
                                          N, H, W, D, Ht, Wt, C, filter_mode, boundary_mode, _ptr(out), _stream(dev))
    return out


def _texture_grad(tex, uv, dy, uv_da, mip_level_bias, mip_wrapper, mip_stack, filter_mode, boundary_mode, need_uv=True, need_da=True, need_bias=True):
    """-> (g_tex, g_uv, g_uv_da, g_mip_level_bias, g_mip_stack) (csrc/torch/torch_texture.cpp:421-691).
    need_*: gradients autograd does not ask for are neither allocated nor written (the reference always writes them)."""
    fn = "texture_grad_linear_mipmap_linear"
    lib = _lib.load()
    dev, (N, H, W, D, Ht, Wt, C), enable_mip, has_uv_da, has_bias, level_max, ptrs, keep = \
        _tex_setup(fn, tex, uv, uv_da, mip_level_bias, mip_wrapper, mip_stack, filter_mode, boundary_mode)
    _check_inputs(fn, f32=[("tex", tex), ("dy", dy)], contiguous=False)
    _require(tuple(dy.shape) == (N, H, W, C), fn, "dy must have shape [minibatch_size, height, width, channels]")
    dy = dy.contiguous()
    has_stack = len(mip_stack) > 0
    with torch.cuda.device(dev):
        g_tex = torch.empty_like(tex)
        g_uv = torch.empty_like(uv) if (filter_mode != 0 and need_uv) else None
        g_uv_da = torch.empty_like(uv_da) if (filter_mode == 3 and has_uv_da and need_da) else None
        g_bias = torch.empty_like(mip_level_bias) if (filter_mode == 3 and has_bias and need_bias) else None
        g_ptrs = (ctypes.c_void_p * (_TEX_MAX_MIP_LEVEL + 1))()
        g_stack = []
        if enable_mip and has_stack:        # custom mip stack: every level receives its own gradient tensor
            for i in range(1, level_max + 1):
                g = torch.empty_like(mip_stack[i - 1])
                g_stack.append(g)
                g_ptrs[i] = g.data_ptr()
        nbytes = lib.b200dr_texture_grad_scratch_bytes(6 * D if boundary_mode == 0 else D, Ht, Wt, C, level_max if enable_mip else 0)
        scratch = torch.empty((nbytes,), dtype=torch.uint8, device=dev)
        _lib.call("b200dr_texture_grad", _ptr(keep[0]), ptrs, level_max, _ptr(uv), _ptr(uv_da) if (enable_mip and has_uv_da) else None,
                                           _ptr(mip_level_bias) if (enable_mip and has_bias) else None, _ptr(dy),
                                           N, H, W, D, Ht, Wt, C, filter_mode, boundary_mode,
                                           _ptr(g_tex), g_ptrs, int(enable_mip and not has_stack),
                                           _ptr(g_uv), _ptr(g_uv_da), _ptr(g_bias), _ptr(scratch), nbytes, _stream(dev))
    return g_tex, g_uv, g_uv_da, g_bias, g_stack


class _texture_func_mip(torch.autograd.Function):
    @staticmethod
    def forward(ctx, filter_mode, tex, uv, uv_da, mip_level_bias, mip_wrapper, filter_mode_enum, boundary_mode_enum, *mip_stack):
        out = _texture_fwd(tex, uv, uv_da, mip_level_bias, mip_wrapper, mip_stack, filter_mode_enum, boundary_mode_enum)
        ctx.has_uv_da = uv_da is not None
        ctx.has_bias = mip_level_bias is not None
        empty = torch.tensor([])
        ctx.save_for_backward(tex, uv, uv_da if uv_da is not None else empty, mip_level_bias if mip_level_bias is not None else empty, *mip_stack)
        ctx.saved_misc = filter_mode, mip_wrapper, filter_mode_enum, boundary_mode_enum
        return out

    @staticmethod
    def backward(ctx, dy):
        tex, uv, uv_da, mip_level_bias, *mip_stack = ctx.saved_tensors
        uv_da = uv_da if ctx.has_uv_da else None
        mip_level_bias = mip_level_bias if ctx.has_bias else None
        filter_mode, mip_wrapper, filter_mode_enum, boundary_mode_enum = ctx.saved_misc
        need = ctx.needs_input_grad
        g_tex, g_uv, g_uv_da, g_bias, g_stack = _texture_grad(tex, uv, dy, uv_da, mip_level_bias, mip_wrapper, mip_stack, filter_mode_enum, boundary_mode_enum,
                                                              need_uv=need[2], need_da=need[3], need_bias=need[4])
        if filter_mode == 'linear-mipmap-linear':
            return (None, g_tex, g_uv, g_uv_da, g_bias, None, None, None) + tuple(g_stack)
        return (None, g_tex, g_uv, None, None, None, None, None) + tuple(g_stack)


class _texture_func(torch.autograd.Function):
    @staticmethod
    def forward(ctx, filter_mode, tex, uv, filter_mode_enum, boundary_mode_enum):
        out = _texture_fwd(tex, uv, None, None, None, (), filter_mode_enum, boundary_mode_enum)
        ctx.save_for_backward(tex, uv)
        ctx.saved_misc = filter_mode, filter_mode_enum, boundary_mode_enum
        return out

    @staticmethod
    def backward(ctx, dy):
        tex, uv = ctx.saved_tensors
        filter_mode, filter_mode_enum, boundary_mode_enum = ctx.saved_misc
        g_tex, g_uv, _a, _b, _c = _texture_grad(tex, uv, dy, None, None, None, (), filter_mode_enum, boundary_mode_enum, need_uv=ctx.needs_input_grad[2])
        if filter_mode == 'linear':
            return None, g_tex, g_uv, None, None
        return None, g_tex, None, None, None


def _fusion_plan(tex, uv, uv_da, mip_level_bias, mip_stack, filter_mode, boundary_mode):
    """-> dict describing the fused backward, or None when the op-by-op path must be taken (see the comment at the top)."""
    if not (_fusion_enabled and torch.is_grad_enabled()) or boundary_mode == 'cube' or mip_level_bias is not None or mip_stack:
        return None
    if filter_mode not in ('linear', 'linear-mipmap-nearest', 'linear-mipmap-linear'):
        return None
    if not (tex.is_cuda and tex.dim() == 4 and tex.shape[-1] <= 4 and tex.dtype == torch.float32):
        return None
    prov = getattr(uv, "_b200dr_prov", None)
    if prov is None or prov.kind != "interpolate" or not _pristine(uv, prov, 0):
        return None
    mip = 'mipmap' in filter_mode
    attr, rast, attr_tri, rast_db = prov.inputs
    if mip:
        if uv_da is None or not prov.meta["diff_all"] or rast_db is None or not _pristine(uv_da, prov, 1):
            return None
    elif uv_da is not None:
        return None
    if attr.shape[-1] != 2 or attr.dtype != torch.float32 or not attr.is_contiguous():
        return None
    if not (uv.requires_grad or tex.requires_grad):
        return None
    plan = dict(attr=attr, attr_tri=attr_tri, rast=rast, rast_db=rast_db if mip else None, pos=None, pos_tri=None, grad_db=False,
                attr_per_image=bool(attr.dim() == 3 and attr.shape[0] > 1))
    rprov = getattr(rast, "_b200dr_prov", None)
    if (rprov is not None and rprov.kind == "rasterize" and _pristine(rast, rprov, 0) and rprov.inputs[0].dim() == 3
            and (rast_db is None or not mip or _pristine(rast_db, rprov, 1))):
        plan.update(pos=rprov.inputs[0], pos_tri=rprov.inputs[1], grad_db=bool(rprov.meta["grad_db"]))
    return plan


def _texture_chain_grad(tex, dy, mip_wrapper, filter_mode, boundary_mode, attr, attr_tri, rast, rast_db, pos, pos_tri, grad_db,
                        attr_per_image, need_tex, need_attr, need_pos, need_rast):
    """Host side of b200dr_texture_chain_grad (include/b200dr.h): -> (g_tex, g_attr, g_pos, g_rast, g_rast_db)."""
    fn = "texture_chain_grad"
    lib = _lib.load()
    dev = _check_inputs(fn, f32=[("tex", tex), ("attr", attr), ("rast", rast), ("rast_db", rast_db), ("pos", pos)], i32=[("tri", attr_tri), ("pos_tri", pos_tri)])
    _check_inputs(fn, f32=[("tex", tex), ("dy", dy)], contiguous=False)
    N, H, W = rast.shape[0], rast.shape[1], rast.shape[2]
    D, Ht, Wt, C = tex.shape
    _require(tuple(dy.shape) == (N, H, W, C), fn, "dy must have shape [minibatch_size, height, width, channels]")
    dy = dy.contiguous()
    enable_mip = filter_mode in (2, 3)
    Cp = (C + 3) & ~3
    level_max = 0
    ptrs = (ctypes.c_void_p * (_TEX_MAX_MIP_LEVEL + 1))()
    with torch.cuda.device(dev):
        keep = _tex_pad(tex)
        if enable_mip:
            total, offs, level_max = _mip_info(fn, (D, Ht, Wt, Cp), mip_wrapper.max_mip_level)
            base = mip_wrapper.mip.data_ptr()
            for i in range(1, level_max + 1):
                ptrs[i] = base + 4 * offs[i]
        Va = attr.shape[-2]
        use_pos = pos is not None and need_pos                  # carry the gradient through rasterize's backward as well
        g_tex = torch.empty_like(tex) if need_tex else None
        g_attr = torch.zeros_like(attr) if need_attr else None
        g_pos = torch.zeros_like(pos) if use_pos else None
        g_rast = torch.empty_like(rast) if (not use_pos and need_rast) else None
        g_rast_db = torch.empty_like(rast_db) if (g_rast is not None and rast_db is not None) else None
        nbytes = lib.b200dr_texture_grad_scratch_bytes(D, Ht, Wt, C, level_max)
        scratch = torch.empty((nbytes,), dtype=torch.uint8, device=dev)
        _lib.call("b200dr_texture_chain_grad", _ptr(keep), ptrs, level_max, _ptr(dy),
                  _ptr(attr), _ptr(attr_tri), Va, int(attr_per_image), _ptr(rast), _ptr(rast_db),
                  _ptr(pos) if use_pos else None, _ptr(pos_tri) if use_pos else None, pos.shape[1] if use_pos else 0, int(grad_db),
                  N, H, W, attr_tri.shape[0], D, Ht, Wt, C, filter_mode, boundary_mode,
                  _ptr(g_tex), _ptr(g_attr), _ptr(g_rast), _ptr(g_rast_db), _ptr(g_pos), _ptr(scratch), nbytes, _stream(dev))
    return g_tex, g_attr, g_pos, g_rast, g_rast_db


class _texture_chain_func(torch.autograd.Function):
    """texture() whose backward runs the fused chain kernel.  Differentiable inputs: tex, attr and either pos (rasterize's
    outputs proven pristine) or rast / rast_db; uv and uv_da enter as constants."""

    @staticmethod
    def forward(ctx, tex, attr, pos, rast, rast_db, uv, uv_da, mip_wrapper, filter_mode_enum, boundary_mode_enum, plan):
        out = _texture_fwd(tex, uv, uv_da, None, mip_wrapper, (), filter_mode_enum, boundary_mode_enum)
        empty = torch.tensor([])
        ctx.save_for_backward(tex, attr, plan["attr_tri"], plan["rast"], plan["rast_db"] if plan["rast_db"] is not None else empty,
                              plan["pos"] if plan["pos"] is not None else empty, plan["pos_tri"] if plan["pos_tri"] is not None else empty)
        ctx.saved_misc = mip_wrapper, filter_mode_enum, boundary_mode_enum, plan["grad_db"], plan["attr_per_image"], plan["rast_db"] is not None, plan["pos"] is not None
        return out

    @staticmethod
    def backward(ctx, dy):
        tex, attr, attr_tri, rast, rast_db, pos, pos_tri = ctx.saved_tensors
        mip_wrapper, fm, bm, grad_db, attr_per_image, has_db, has_pos = ctx.saved_misc
        need = ctx.needs_input_grad
        g_tex, g_attr, g_pos, g_rast, g_rast_db = _texture_chain_grad(
            tex, dy, mip_wrapper, fm, bm, attr, attr_tri, rast, rast_db if has_db else None, pos if has_pos else None, pos_tri if has_pos else None,
            grad_db, attr_per_image, need[0], need[1], need[2], need[3] or need[4])
        return (g_tex, g_attr, g_pos, g_rast if need[3] else None, g_rast_db if need[4] else None, None, None, None, None, None, None)


def texture(tex, uv, uv_da=None, mip_level_bias=None, mip=None, filter_mode='auto', boundary_mode='wrap', max_mip_level=None):
    """Perform texture sampling.

    Args:
        tex: float32 texture [minibatch_size, tex_height, tex_width, tex_channels] (minibatch 1 broadcasts);
             [minibatch_size, 6, tex_height, tex_width, tex_channels] in cube map mode.
        uv: per-pixel texture coordinates [minibatch_size, height, width, 2]; 3-D direction vectors
            [minibatch_size, height, width, 3] in cube map mode.
        uv_da: (optional) image-space derivatives of `uv`, [minibatch_size, height, width, 4] (6 in cube map mode).
        mip_level_bias: (optional) per-pixel mip level bias [minibatch_size, height, width].
        mip: (optional) result of `texture_construct_mip()` or a list of tensors forming a custom mip stack.
        filter_mode: 'auto', 'nearest', 'linear', 'linear-mipmap-nearest', 'linear-mipmap-linear'.
        boundary_mode: 'wrap', 'clamp', 'zero', 'cube'.
        max_mip_level: limits the number of mip levels built / used.

    Returns:
        [minibatch_size, height, width, tex_channels].
    """
    for name, t in (("tex", tex), ("uv", uv)):
        assert isinstance(t, torch.Tensor), "%s must be a tensor" % name
    max_mip_level = _mip_limit(max_mip_level)
    has_footprint = uv_da is not None or mip_level_bias is not None
    if filter_mode == 'auto':                                   # mip-mapped exactly when a footprint was given
        filter_mode = 'linear-mipmap-linear' if has_footprint else 'linear'
    if 'mipmap' in filter_mode:
        assert isinstance(uv_da, torch.Tensor) or isinstance(mip_level_bias, torch.Tensor), "mip-mapped filtering needs uv_da or mip_level_bias"
        if max_mip_level == 0 and filter_mode.startswith('linear-'):
            filter_mode = 'linear'                              # only the base level may be used
    filter_mode_enum, boundary_mode_enum = _FILTER_MODES[filter_mode], _BOUNDARY_MODES[boundary_mode]

    mip_wrapper, mip_stack = None, []
    if 'mipmap' in filter_mode:
        if mip is None:
            mip_wrapper = _texture_construct_mip(tex, max_mip_level, boundary_mode == 'cube')
        elif isinstance(mip, TextureMipWrapper):
            mip_wrapper = mip
        else:
            assert isinstance(mip, list) and all(isinstance(x, torch.Tensor) for x in mip), "mip: texture_construct_mip() result or a list of tensors"
            mip_stack = mip
    plan = _fusion_plan(tex, uv, uv_da if 'mipmap' in filter_mode else None, mip_level_bias, mip_stack, filter_mode, boundary_mode)
    if plan is not None:
        # the intermediates enter as constants (detached): the fused node owns the whole path colour -> (tex, attr, pos)
        if plan["pos"] is not None:
            return _texture_chain_func.apply(tex, plan["attr"], plan["pos"], plan["rast"].detach(), None, uv.detach(),
                                             uv_da.detach() if 'mipmap' in filter_mode else None, mip_wrapper, filter_mode_enum, boundary_mode_enum, plan)
        return _texture_chain_func.apply(tex, plan["attr"], None, plan["rast"], plan["rast_db"], uv.detach(),
                                         uv_da.detach() if 'mipmap' in filter_mode else None, mip_wrapper, filter_mode_enum, boundary_mode_enum, plan)
    if 'mipmap' in filter_mode:
        return _texture_func_mip.apply(filter_mode, tex, uv, uv_da, mip_level_bias, mip_wrapper, filter_mode_enum, boundary_mode_enum, *mip_stack)
    return _texture_func.apply(filter_mode, tex, uv, filter_mode_enum, boundary_mode_enum)


def texture_construct_mip(tex, max_mip_level=None, cube_mode=False):
    """Construct a mipmap stack for a texture that stays constant (pass the result as `mip` to `texture()`)."""
    assert isinstance(tex, torch.Tensor), "tex must be a tensor"
    assert isinstance(cube_mode, bool)
    return _texture_construct_mip(tex, _mip_limit(max_mip_level), cube_mode)


# ----------------------------------------------------------------------------------------------
# Antialias (nvdiffrast/torch/ops.py:471-543, csrc/torch/torch_antialias.cpp).


class TopologyHashWrapper:
    """Opaque topology object returned by `antialias_construct_topology_hash()` (csrc/torch/torch_types.h:41-45).
    Holds a [num_triangles, 4] int32 table of the three opposite ('wing') vertices of every triangle."""

    def __init__(self, wings):
        self.wings = wings


def _antialias_construct_topology(tri):
    fn = "antialias_construct_topology_hash"
    lib = _lib.load()
    dev = _check_inputs(fn, i32=[("tri", tri)])
    _require(tri.dim() == 2 and tri.shape[0] > 0 and tri.shape[1] == 3, fn, "tri must have shape [>0, 3]")
    T = tri.shape[0]
    with torch.cuda.device(dev):
        wings = torch.empty((T, 4), dtype=torch.int32, device=dev)
        nbytes = lib.b200dr_topology_scratch_bytes(T)
        scratch = torch.empty((nbytes,), dtype=torch.uint8, device=dev)
        _lib.call("b200dr_antialias_construct_topology", _ptr(tri), T, 0x7fffffff, _ptr(wings), _ptr(scratch), nbytes, _stream(dev))
    return TopologyHashWrapper(wings)


def _aa_dims(fn, color, rast, pos, tri):
    instance_mode = pos.dim() > 2
    _require(color.dim() == 4 and min(color.shape) > 0, fn, "color must have shape[>0, >0, >0, >0]")
    _require(rast.dim() == 4 and min(rast.shape[:3]) > 0 and rast.shape[3] == 4, fn, "rast must have shape[>0, >0, >0, 4]")
    _require(tri.dim() == 2 and tri.shape[0] > 0 and tri.shape[1] == 3, fn, "tri must have shape [>0, 3]")
    _require(color.shape[1] == rast.shape[1] and color.shape[2] == rast.shape[2], fn, "color and rast inputs must have same spatial dimensions")
    if instance_mode:
        _require(pos.dim() == 3 and pos.shape[0] > 0 and pos.shape[1] > 0 and pos.shape[2] == 4, fn, "pos must have shape [>0, >0, 4] or [>0, 4]")
        _require(rast.shape[0] == color.shape[0] and pos.shape[0] == color.shape[0], fn, "minibatch size mismatch between inputs color, rast, pos")
    else:
        _require(pos.dim() == 2 and pos.shape[0] > 0 and pos.shape[1] == 4, fn, "pos must have shape [>0, >0, 4] or [>0, 4]")
        _require(rast.shape[0] == color.shape[0], fn, "minibatch size mismatch between inputs color, rast")
    N, H, W, C = color.shape
    V = pos.shape[1] if instance_mode else pos.shape[0]
    return N, H, W, C, V, int(instance_mode)


def _antialias_fwd(color, rast, pos, tri, topology_hash):
    fn = "antialias_fwd"
    lib = _lib.load()
    wings = topology_hash.wings
    dev = _check_inputs(fn, f32=[("color", color), ("rast", rast), ("pos", pos)], i32=[("tri", tri), ("topology_hash", wings)])
    N, H, W, C, V, inst = _aa_dims(fn, color, rast, pos, tri)
    _require(wings.dim() == 2 and wings.shape[0] == tri.shape[0] and wings.shape[1] == 4, fn, "topology hash does not match the triangle tensor")
    with torch.cuda.device(dev):
        out = torch.empty_like(color)
        work_items = torch.empty((2 * N * H * W, 4), dtype=torch.int32, device=dev)
        work_counters = torch.empty((4,), dtype=torch.int32, device=dev)
        _lib.call("b200dr_antialias_fwd", _ptr(color), _ptr(rast), _ptr(pos), _ptr(tri), _ptr(wings), N, H, W, C, V, tri.shape[0], inst,
                                            _ptr(out), _ptr(work_items), _ptr(work_counters), _stream(dev))
    return out, (work_items, work_counters)


def _antialias_grad(color, rast, pos, tri, dy, work_buffer):
    fn = "antialias_grad"
    lib = _lib.load()
    work_items, work_counters = work_buffer
    dev = _check_inputs(fn, f32=[("color", color), ("rast", rast), ("pos", pos)], i32=[("tri", tri)])
    _check_inputs(fn, f32=[("color", color), ("dy", dy)], contiguous=False)
    N, H, W, C, V, inst = _aa_dims(fn, color, rast, pos, tri)
    _require(tuple(dy.shape) == tuple(color.shape), fn, "dy must have same shape as color")
    dy = dy.contiguous()
    with torch.cuda.device(dev):
        g_color = torch.empty_like(color)
        g_pos = torch.zeros_like(pos)
        _lib.call("b200dr_antialias_grad", _ptr(color), _ptr(rast), _ptr(pos), _ptr(tri), _ptr(dy), _ptr(work_items), _ptr(work_counters),
                                             N, H, W, C, V, tri.shape[0], inst, _ptr(g_color), _ptr(g_pos), _stream(dev))
    return g_color, g_pos


class _antialias_func(torch.autograd.Function):
    @staticmethod
    def forward(ctx, color, rast, pos, tri, topology_hash, pos_gradient_boost):
        out, work_buffer = _antialias_fwd(color, rast, pos, tri, topology_hash)
        ctx.save_for_backward(color, rast, pos, tri)
        ctx.saved_misc = pos_gradient_boost, work_buffer
        return out

    @staticmethod
    def backward(ctx, dy):
        color, rast, pos, tri = ctx.saved_tensors
        pos_gradient_boost, work_buffer = ctx.saved_misc
        g_color, g_pos = _antialias_grad(color, rast, pos, tri, dy, work_buffer)
        if pos_gradient_boost != 1.0:
            g_pos = g_pos * pos_gradient_boost
        return g_color, None, g_pos, None, None, None


def antialias(color, rast, pos, tri, topology_hash=None, pos_gradient_boost=1.0):
    """Perform antialiasing.

    Silhouette detection relies on vertex indices: a vertex shared by several triangles must be referenced by the same
    index everywhere (un-welded meshes make every edge a silhouette).

    Args:
        color: image to antialias, [minibatch_size, height, width, num_channels].
        rast: main output of `rasterize()`.
        pos, tri: the tensors used in the rasterization.
        topology_hash: (optional) result of `antialias_construct_topology_hash(tri)`.
        pos_gradient_boost: (optional) multiplier for gradients propagated to `pos`.

    Returns:
        The antialiased image, same shape as `color`.
    """
    assert all(isinstance(x, torch.Tensor) for x in (color, rast, pos, tri))
    if topology_hash is not None:
        assert isinstance(topology_hash, TopologyHashWrapper)
    else:
        topology_hash = _antialias_construct_topology(tri)
    return _antialias_func.apply(color, rast, pos, tri, topology_hash, pos_gradient_boost)


def antialias_construct_topology_hash(tri):
    """Precompute the topology object for a constant triangle tensor [num_triangles, 3] (int32, GPU)."""
    assert isinstance(tri, torch.Tensor)
    return _antialias_construct_topology(tri)


# ----------------------------------------------------------------------------------------------
# Legacy OpenGL context stub (nvdiffrast/torch/ops.py:550-559).


class RasterizeGLContext(RasterizeCudaContext):
    def __init__(self, output_db=True, mode='automatic', device=None):
        warnings.warn("RasterizeGLContext has been deprecated and uses RasterizeCudaContext internally", DeprecationWarning, stacklevel=2)
        super().__init__(device=device)

    def set_context(self):
        pass

    def release_context(self):
        pass
