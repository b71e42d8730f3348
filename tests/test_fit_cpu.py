"""The miniature inverse-rendering loops of tests/fit_scenes.py on the CPU oracle: pins, without a GPU, what the loops are
expected to do (the GPU run in tests/test_fit_gpu.py must converge the same way).  The reference's samples (earth.py, cube.py)
are the model: a texture learned from renders; vertex positions and colours learned from renders through the antialias and
rasterize gradients."""
import numpy as np

import fit_scenes as fs


def test_oracle_learns_texture():
    s = fs.texture_scene()
    losses, tex = fs.fit_texture_oracle(s)
    assert losses[-1] < 5e-3 * losses[0], (losses[0], losses[-1])                # measured 1.0e-3
    assert np.abs(tex - s["tex_target"]).mean() < 0.03                             # measured 0.011 (initial 0.295)


def test_oracle_learns_cube_geometry_and_colours():
    s = fs.cube_scene()
    losses, verts, col = fs.fit_cube_oracle(s)
    assert losses[-1] < 1e-3 * losses[0], (losses[0], losses[-1])                # measured 2.7e-5
    assert fs.geometric_error(verts, s["verts_target"]) < 0.1 * fs.geometric_error(s["verts_init"], s["verts_target"])   # measured 0.002 of the initial error
    assert np.abs(col - s["col_target"]).mean() < 0.01                              # measured 5.6e-4 (initial 0.33)
