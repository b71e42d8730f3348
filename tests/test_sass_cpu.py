"""Build facts of libb200dr.so that DESIGN.md section 4 states, checked on the shipped binary with cuobjdump (no GPU needed):
sm_100a cubins only; TMA bulk copies in the rasterizer; 128-bit vector reductions; the fused backward's L2 prefetches and its
device-side work counter; no local-memory spills worth mentioning in the fused kernel that dominates the step."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CUOBJDUMP = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
pytestmark = pytest.mark.skipif(not os.path.exists(CUOBJDUMP), reason="cuobjdump not available")


@pytest.fixture(scope="module")
def sass():
    from nvdiffrast_b200 import _lib
    _lib.load()                                             # builds the library if the sources are newer
    r = subprocess.run([CUOBJDUMP, "-sass", _lib.so_path()], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-500:]
    return r.stdout


def _functions(sass):
    """{mangled name: text of its SASS}"""
    parts = re.split(r"\n\s*Function : ", sass)
    return {p.split("\n", 1)[0].strip(): p for p in parts[1:]}


def test_only_sm100a_code(sass):
    archs = set(re.findall(r"arch = (sm_\w+)", sass))
    assert archs == {"sm_100a"}, archs


def test_blackwell_specific_instructions_are_present(sass):
    fns = _functions(sass)
    bins = [t for n, t in fns.items() if "raster_bins_kernel" in n]
    assert len(bins) == 3 and all("UBLKCP" in t and "SYNCS" in t for t in bins)          # cp.async.bulk + mbarrier per variant
    assert sass.count("REDG.E.ADD.F32x4") > 500 and sass.count("REDG.E.ADD.F32x2") > 100   # vector reductions, not scalar atomics
    assert "UTMALDG" not in sass and "HMMA" not in sass and "UTCHMMA" not in sass             # no tensor maps / tensor cores on this path


def test_fused_chain_backward_kernel(sass):
    fns = _functions(sass)
    # c3's instance: C = 3, trilinear, wrap, through rasterize, STRIPS schedule
    name = [n for n in fns if "texture_chain_grad_fastILi3ELi3ELi1ELb1ELb1E" in n]
    assert len(name) == 1, [n for n in fns if "texture_chain_grad_fast" in n][:4]
    t = fns[name[0]]
    assert t.count("ATOMG.E.ADD") == 2                       # the strip counter: current strip + the one after it
    assert t.count("CCTL.E.PF2") >= 3                        # L2 prefetch of rast, dy, rast_db of the next row
    assert t.count("REDG.E.ADD.F32x4") >= 10                 # texel taps + position gradients
    assert len(re.findall(r"\bSTL\b", t)) <= 4 and len(re.findall(r"\bLDL\b", t)) <= 4      # 8 bytes of spills at 96 registers
    # the grid-stride instance exists for widths that are not a multiple of 32
    assert any("texture_chain_grad_fastILi3ELi3ELi1ELb1ELb0E" in n for n in fns)
