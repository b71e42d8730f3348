"""nvdiffrast_b200 -- B200-native (sm_100a) differentiable rasterizer behind the nvdiffrast.torch op surface.

    import nvdiffrast_b200.torch as dr      # same names / arguments / return shapes as nvdiffrast.torch

`install_as_nvdiffrast()` registers this package under the name `nvdiffrast` so that unmodified pipelines
(`import nvdiffrast.torch as dr`) pick it up.
"""
__version__ = "0.1.0"


def install_as_nvdiffrast():
    import sys
    from . import torch as _t
    sys.modules.setdefault("nvdiffrast", sys.modules[__name__])
    sys.modules.setdefault("nvdiffrast.torch", _t)
