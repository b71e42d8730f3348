"""End-to-end use of the op surface the way the reference's samples use it (samples/torch/earth.py, cube.py in miniature,
tests/fit_scenes.py): Adam loops that learn a texture, and vertex positions + colours, from renders.  The GPU loop must start
from the same loss as the CPU oracle's loop (same forward math: rel 1e-3) and converge like it (bounds = the oracle's measured
result with a 3-10x margin; trajectories are not compared step by step, Adam amplifies last-bit differences)."""
import numpy as np
import pytest
import torch

import fit_scenes as fs

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dr():
    import nvdiffrast_b200.torch as dr
    return dr


@pytest.fixture(scope="module")
def oracle_texture_fit():
    s = fs.texture_scene()
    return (s,) + tuple(fs.fit_texture_oracle(s))


@pytest.mark.parametrize("fusion", [True, False])
def test_learns_texture(dr, oracle_texture_fit, fusion):
    s, o_losses, o_tex = oracle_texture_fit
    prev = dr.get_fusion()
    dr.set_fusion(fusion)
    try:
        losses, tex = fs.fit_texture_b200(s, dr, torch.device("cuda", 0))
    finally:
        dr.set_fusion(prev)
    assert abs(losses[0] - o_losses[0]) <= 1e-3 * o_losses[0], (losses[0], o_losses[0])
    assert losses[-1] < 5e-3 * losses[0], (losses[0], losses[-1], o_losses[-1])
    assert np.abs(tex - s["tex_target"]).mean() < 0.03


def test_learns_cube_geometry_and_colours(dr):
    s = fs.cube_scene()
    o_losses, o_verts, o_col = fs.fit_cube_oracle(s)
    losses, verts, col = fs.fit_cube_b200(s, dr, torch.device("cuda", 0))
    assert abs(losses[0] - o_losses[0]) <= 1e-3 * o_losses[0], (losses[0], o_losses[0])
    assert losses[-1] < 1e-3 * losses[0], (losses[0], losses[-1], o_losses[-1])
    assert fs.geometric_error(verts, s["verts_target"]) < 0.1 * fs.geometric_error(s["verts_init"], s["verts_target"])
    assert np.abs(col - s["col_target"]).mean() < 0.01
