"""CPU-only checks of the drop-in boundary: the C-ABI library builds for sm_100a, loads without a GPU, and exports
every symbol include/b200dr.h declares (no compute calls here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "b200dr.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b200dr_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_reference_op_surface():
    syms = declared_symbols()
    # one entry point per native op of the reference's pybind module (csrc/torch/torch_bindings.cpp:43-71)
    for needed in ["b200dr_rasterize_fwd", "b200dr_rasterize_grad", "b200dr_interpolate_fwd", "b200dr_interpolate_grad",
                   "b200dr_antialias_construct_topology", "b200dr_antialias_fwd", "b200dr_antialias_grad",
                   "b200dr_texture_construct_mip", "b200dr_texture_fwd", "b200dr_texture_grad",
                   "b200dr_get_log_level", "b200dr_set_log_level", "b200dr_last_error"]:
        assert needed in syms


def test_library_loads_and_exports_every_declared_symbol():
    from nvdiffrast_b200 import _lib
    lib = _lib.load()
    raw = ctypes.CDLL(_lib.so_path())
    for name in declared_symbols():
        assert hasattr(raw, name), "libb200dr.so does not export %s" % name
    assert set(_lib.SIGNATURES) == set(declared_symbols()), "ctypes signature table out of sync with include/b200dr.h"
    assert lib.b200dr_version() == 1


def test_host_side_helpers_without_gpu():
    from nvdiffrast_b200 import _lib
    lib = _lib.load()
    lib.b200dr_set_log_level(2)
    assert lib.b200dr_get_log_level() == 2
    lib.b200dr_set_log_level(1)
    # path heuristic and scratch sizing are pure host code
    assert lib.b200dr_rasterize_pick_path(16, 5856, 2048, 2048) == 1
    assert lib.b200dr_rasterize_pick_path(8, 1000000, 2048, 2048) == 2
    assert lib.b200dr_rasterize_scratch_bytes(8, 1000000, 2048, 2048, 2) == 8 * 2048 * 2048 * 8
    assert lib.b200dr_rasterize_scratch_bytes(1, 12, 64, 64, 1) >= 7 * 12 * 40
    # mip layout matches the oracle's restatement of calculateMipInfo
    import oracle
    for shape, limit in [((1, 1024, 1024, 3), -1), ((2, 16, 64, 4), -1), ((1, 32, 32, 1), 3), ((1, 1, 8, 2), -1)]:
        offs = (ctypes.c_int * 17)(); ml = ctypes.c_int(0)
        tot = lib.b200dr_texture_mip_info(shape[2], shape[1], shape[0], shape[3], limit, offs, ctypes.byref(ml))
        otot, ooffs, oml = oracle.mip_info(shape, limit)
        assert (tot, ml.value) == (otot, oml) and list(offs)[:oml + 1] == list(ooffs)[:oml + 1]
    offs = (ctypes.c_int * 17)(); ml = ctypes.c_int(0)
    assert lib.b200dr_texture_mip_info(24, 20, 1, 3, -1, offs, ctypes.byref(ml)) == -1      # 20 -> 10 -> 5: odd extent


def test_ops_refuse_cpu_tensors():
    import torch
    import nvdiffrast_b200.torch as dr
    with pytest.raises(RuntimeError):
        dr.interpolate(torch.zeros(3, 2), torch.zeros(1, 4, 4, 4), torch.zeros(1, 3, dtype=torch.int32))
    with pytest.raises(RuntimeError):
        dr.antialias_construct_topology_hash(torch.zeros(4, 3, dtype=torch.int32))
