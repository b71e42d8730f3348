"""The fused chain backward has two work schedules (csrc/texture.cu, TCG_STRIPS): 32 x 8 pixel strips taken from a work counter
when the image width is a multiple of 32 and the height a multiple of 8, and the grid-stride loop over the flat pixel array
otherwise (its 32-pixel groups straddle rows and images).  Every other chain test renders at sizes that take the strip schedule;
this file holds the grid-stride one to the same fused == op-by-op bar (tests/test_fusion_gpu.py: forward 1e-6, accumulated
gradients 2e-5 |g|_inf)."""
import pytest

import workloads as wl
from test_fusion_gpu import _compare

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dr():
    import nvdiffrast_b200.torch as dr
    return dr


@pytest.fixture(scope="module")
def ctx(dr):
    return dr.RasterizeCudaContext()


@pytest.mark.parametrize("res", [100, 72])
def test_fused_chain_at_widths_that_are_not_a_multiple_of_32(dr, ctx, res):
    _compare(dr, ctx, wl.config_c3(n=3, res=res, tex_res=64), 3)
