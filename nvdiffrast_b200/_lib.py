"""ctypes binding of libb200dr.so (C ABI declared in include/b200dr.h).

There is no CPU fallback: if the CUDA library cannot be loaded (or built), importing the ops fails loudly.
"""
import ctypes
import os

from . import build as _build

_c = ctypes
_vp = _c.c_void_p
_int = _c.c_int
_sz = _c.c_size_t

# name -> (restype, argtypes); mirrors include/b200dr.h line by line.
SIGNATURES = {
    "b200dr_version": (_int, []),
    "b200dr_last_error": (_c.c_char_p, []),
    "b200dr_get_log_level": (_int, []),
    "b200dr_set_log_level": (None, [_int]),
    "b200dr_launch_count": (_c.c_ulonglong, []),
    "b200dr_rasterize_scratch_bytes": (_sz, [_int, _int, _int, _int, _int]),
    "b200dr_rasterize_pick_path": (_int, [_int, _int, _int, _int]),
    "b200dr_rasterize_fwd": (_int, [_vp, _vp, _vp, _int, _int, _int, _int, _int, _int, _int, _vp, _vp, _vp, _sz, _int, _vp, _vp, _vp]),
    "b200dr_rasterize_grad": (_int, [_vp, _vp, _vp, _vp, _vp, _int, _int, _int, _int, _int, _int, _vp, _vp]),
    "b200dr_interpolate_fwd": (_int, [_vp, _vp, _vp, _vp] + [_int] * 9 + [_vp, _int, _vp, _vp, _vp]),
    "b200dr_interpolate_grad": (_int, [_vp, _vp, _vp, _vp, _vp, _vp] + [_int] * 9 + [_vp, _int, _vp, _vp, _vp, _vp]),
    "b200dr_topology_scratch_bytes": (_sz, [_int]),
    "b200dr_antialias_construct_topology": (_int, [_vp, _int, _int, _vp, _vp, _sz, _vp]),
    "b200dr_antialias_fwd": (_int, [_vp, _vp, _vp, _vp, _vp] + [_int] * 7 + [_vp, _vp, _vp, _vp]),
    "b200dr_antialias_grad": (_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp] + [_int] * 7 + [_vp, _vp, _vp]),
    "b200dr_texture_mip_info": (_c.c_longlong, [_int, _int, _int, _int, _int, _c.POINTER(_int), _c.POINTER(_int)]),
    "b200dr_texture_pad": (_int, [_vp, _vp, _c.c_longlong, _int, _vp]),
    "b200dr_texture_construct_mip": (_int, [_vp, _vp, _int, _int, _int, _int, _c.POINTER(_int), _int, _vp]),
    "b200dr_texture_fwd": (_int, [_vp, _c.POINTER(_vp), _int, _vp, _vp, _vp] + [_int] * 9 + [_vp, _vp]),
    "b200dr_texture_grad_scratch_bytes": (_sz, [_int, _int, _int, _int, _int]),
    "b200dr_texture_chain_grad": (_int, [_vp, _c.POINTER(_vp), _int, _vp, _vp, _vp, _int, _int, _vp, _vp, _vp, _vp, _int, _int] + [_int] * 10 +
                                  [_vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "b200dr_interpolate_chain_grad": (_int, [_vp] * 8 + [_int] * 11 + [_vp, _vp, _vp]),
    "b200dr_texture_grad": (_int, [_vp, _c.POINTER(_vp), _int, _vp, _vp, _vp, _vp] + [_int] * 9 + [_vp, _c.POINTER(_vp), _int, _vp, _vp, _vp, _vp, _sz, _vp]),
}

_lib = None


def so_path():
    return _build.SO_PATH


def load():
    """Load (building first if the sources are newer and nvcc is available) the CUDA library."""
    global _lib
    if _lib is not None:
        return _lib
    path = _build.SO_PATH
    if os.environ.get("B200DR_LIB"):        # developer switch: A/B a tuning build made by build.build_variant()
        path = os.environ["B200DR_LIB"]
    elif _build.needs_build():
        try:
            _build.build()
        except FileNotFoundError as e:  # no nvcc on this box: a prebuilt library (it travels with the tree) is acceptable
            if not os.path.exists(path):
                raise RuntimeError("nvdiffrast_b200: libb200dr.so is missing and nvcc is not available (%s). "
                                   "There is no CPU fallback." % e)
            import warnings
            warnings.warn("nvdiffrast_b200: sources are newer than libb200dr.so but nvcc is not available; using the prebuilt library")
        # any other failure (a compile error after a source edit) propagates: never run stale kernels silently
    lib = ctypes.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError here = header/library mismatch: fail loudly
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(ret):
    if ret != 0:
        raise RuntimeError(load().b200dr_last_error().decode())


# Optional per-op timing (bench.py): when a list is installed, every C-ABI op call is bracketed by CUDA events recorded
# on the stream the op's kernels are launched on; entries are (name, start_event, end_event).
_trace = None


def set_trace(lst):
    global _trace
    _trace = lst


def call(name, *args):
    """Invoke one int-returning C-ABI entry point; raise RuntimeError with the library's message on failure."""
    fn = getattr(load(), name)
    if _trace is None:
        ret = fn(*args)
    else:
        import torch
        stream = torch.cuda.ExternalStream(args[-1].value) if args[-1].value else torch.cuda.default_stream()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        ret = fn(*args)
        e1.record(stream)
        _trace.append((name, e0, e1))
    check(ret)
