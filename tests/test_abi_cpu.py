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
    # texture-gradient scratch: channel-padded pyramid (levels 0..max) + the 256-byte bookkeeping tail (strip counter of the fused backward)
    assert lib.b200dr_texture_grad_scratch_bytes(1, 4, 8, 3, 0) == 4 * 8 * 4 * 4 + 256
    assert lib.b200dr_texture_grad_scratch_bytes(2, 4, 4, 5, 2) == 2 * 8 * 4 * (16 + 4 + 1) + 256
    assert lib.b200dr_texture_grad_scratch_bytes(1, 1024, 1024, 3, 9) % 16 == 0
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


def test_c_abi_rejects_bad_arguments_before_touching_the_device():
    """Argument validation of the C entry points runs on the host before any launch: error code + message, no GPU needed.
    (Mirrors the NVDR_CHECK preconditions of csrc/torch/torch_*.cpp.)"""
    from nvdiffrast_b200 import _lib
    lib = _lib.load()
    vp = ctypes.c_void_p
    err = lambda: lib.b200dr_last_error().decode()
    ptrs = (vp * 17)()
    # texture: unsupported modes, empty shapes, mipmapping without uv_da / bias, cube restrictions
    base = dict(tex=vp(256), uv=vp(256), N=1, H=4, W=4, D=1, Ht=8, Wt=8, C=3)
    def tex_fwd(filter_mode, boundary_mode, uv_da=None, bias=None, **kw):
        a = dict(base); a.update(kw)
        return lib.b200dr_texture_fwd(a["tex"], ptrs, 0, a["uv"], uv_da, bias, a["N"], a["H"], a["W"], a["D"], a["Ht"], a["Wt"], a["C"],
                                      filter_mode, boundary_mode, vp(256), None)
    assert tex_fwd(7, 1) != 0 and "filter_mode unsupported" in err()
    assert tex_fwd(1, 9) != 0 and "boundary_mode unsupported" in err()
    assert tex_fwd(1, 1, N=0) != 0 and "uv must have shape" in err()
    assert tex_fwd(1, 1, C=0) != 0 and "tex must have shape" in err()
    assert tex_fwd(1, 1, D=2, N=3) != 0 and "minibatch size mismatch" in err()
    assert tex_fwd(3, 1) != 0 and "requires uv_da and/or mip_level_bias" in err()
    # fused chain backward (extension): cube maps / wide textures / missing pixel differentials are refused on the host
    def chain(filter_mode=3, boundary_mode=1, C=3, rast_db=vp(256), pos=None):
        return lib.b200dr_texture_chain_grad(vp(256), ptrs, 0, vp(256), vp(256), vp(256), 8, 0, vp(256), rast_db, pos, None, 0, 1,
                                             1, 4, 4, 2, 1, 8, 8, C, filter_mode, boundary_mode, vp(256), vp(256), vp(256), vp(256), None, vp(256), 1 << 20, None)
    assert chain(boundary_mode=0) != 0 and "at most 4 channels" in err()
    assert chain(C=5) != 0 and "at most 4 channels" in err()
    assert chain(filter_mode=0) != 0 and "filter_mode unsupported" in err()
    assert chain(rast_db=None) != 0 and "needs rast_db" in err()
    assert chain(pos=vp(256)) != 0 and "pos / tri missing" in err()
    assert tex_fwd(1, 0, Ht=8, Wt=16) != 0 and "square in cube map mode" in err()
    assert tex_fwd(1, 1, uv=vp(260)) != 0 and "not aligned" in err()
    # rasterize
    def rast_fwd(**kw):
        a = dict(pos=vp(256), tri=vp(256), ranges=None, N=1, V=3, T=1, H=16, W=16, inst=1, max_tris=1, scratch=vp(256), sbytes=1 << 20, path=1,
                 rast=vp(256), db=vp(256)); a.update(kw)
        return lib.b200dr_rasterize_fwd(a["pos"], a["tri"], a["ranges"], a["N"], a["V"], a["T"], a["H"], a["W"], a["inst"], a["max_tris"], None, None,
                                        a["scratch"], a["sbytes"], a["path"], a["rast"], a["db"], None)
    assert rast_fwd(T=0) != 0 and "non-empty" in err()
    assert rast_fwd(H=0) != 0 and "resolution" in err()
    assert rast_fwd(inst=0) != 0 and "range mode requires ranges" in err()
    assert rast_fwd(pos=vp(260)) != 0 and "not aligned to float4" in err()
    assert rast_fwd(path=5) != 0 and "invalid path" in err()
    assert rast_fwd(sbytes=8) != 0 and "scratch buffer too small" in err()
    assert rast_fwd(N=70000, sbytes=1 << 40) != 0
    # antialias topology
    assert lib.b200dr_antialias_construct_topology(vp(256), 0, 3, vp(256), vp(256), 1 << 20, None) != 0 and "tri must have shape" in err()
    assert lib.b200dr_antialias_construct_topology(vp(256), 4, 3, vp(256), vp(256), 8, None) != 0 and "scratch buffer too small" in err()


def test_public_api_signatures_match_the_reference_front_end():
    """Drop-in surface: every public function / class of nvdiffrast/torch/ops.py exists in nvdiffrast_b200.torch with the same
    positional parameter names and defaults (fixture: tests/golden/api_signatures.json, made by make_api_signatures.py)."""
    import inspect
    import json
    import nvdiffrast_b200.torch as dr
    ref = json.load(open(os.path.join(ROOT, "tests", "golden", "api_signatures.json")))

    def check(name, fn, want):
        params = list(inspect.signature(fn).parameters.values())
        got = [(p.name, None if p.default is inspect.Parameter.empty else repr(p.default)) for p in params
               if p.kind in (p.POSITIONAL_ONLY, p.POSITIONAL_OR_KEYWORD)]
        exp = [(w["name"], w["default"]) for w in want]
        if name.endswith("__exit__"):
            return                                              # *args catch-all on both sides
        assert got == exp, "%s: %s != reference %s" % (name, got, exp)

    for name, spec in ref.items():
        assert hasattr(dr, name), "missing public name %s" % name
        obj = getattr(dr, name)
        if isinstance(spec, dict):
            if name == "RasterizeGLContext":
                continue                                        # OpenGL context: deliberately a stub that raises (no GL interop)
            for m, want in spec["methods"].items():
                assert hasattr(obj, m), "%s.%s missing" % (name, m)
                check("%s.%s" % (name, m), getattr(obj, m), want)
        else:
            check(name, obj, spec)


def test_header_is_plain_c_and_the_c_example_links():
    """The boundary is a C ABI: include/b200dr.h compiles as C99 and examples/c_abi_triangle.c links against the library
    with nothing but the CUDA runtime (no torch, no C++)."""
    import shutil
    import subprocess
    import tempfile
    from nvdiffrast_b200 import _lib
    _lib.load()
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc")
    hdr = os.path.join(ROOT, "include", "b200dr.h")
    subprocess.check_call([gcc, "-std=c99", "-fsyntax-only", "-x", "c", hdr])
    cuda = "/usr/local/cuda"
    if not os.path.exists(os.path.join(cuda, "include", "cuda_runtime_api.h")):
        pytest.skip("CUDA toolkit headers not found")
    with tempfile.TemporaryDirectory() as tmp:
        subprocess.check_call([gcc, "-std=c99", "-Wall", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(cuda, "include"),
                               os.path.join(ROOT, "examples", "c_abi_triangle.c"), "-o", os.path.join(tmp, "t"),
                               "-L", os.path.dirname(_lib.so_path()), "-lb200dr", "-L", os.path.join(cuda, "lib64"), "-lcudart"])


def test_concurrent_builds_are_serialised():
    """Two processes that find the sources newer than the library (the ranks of one node at start-up) must not write the same object
    files at once: the build takes an inter-process lock and installs the library by rename."""
    import subprocess, sys, time
    from nvdiffrast_b200 import build as b
    if b._nvcc() == "nvcc" and not any(os.access(os.path.join(d, "nvcc"), os.X_OK) for d in os.environ.get("PATH", "").split(os.pathsep)):
        pytest.skip("nvcc not available")
    b.build()
    src = os.path.join(os.path.dirname(b.SO_PATH), "csrc", "api.cu")
    now = time.time() + 2
    os.utime(src, (now, now))                                   # sources newer than the library
    code = "import sys; sys.path.insert(0, %r); from nvdiffrast_b200 import _lib; l = _lib.load(); print(len(_lib.SIGNATURES))" % ROOT
    procs = [subprocess.Popen([sys.executable, "-c", code], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True) for _ in range(2)]
    outs = [p.communicate(timeout=900) for p in procs]
    for p, (o, e) in zip(procs, outs):
        assert p.returncode == 0, e[-2000:]
        assert int(o.strip().splitlines()[-1]) > 10
    assert not b.needs_build()
