"""Same public names as nvdiffrast.torch (nvdiffrast/torch/__init__.py:9-10)."""
from .ops import (RasterizeCudaContext, get_log_level, set_log_level, rasterize, DepthPeeler, interpolate, texture,
                  texture_construct_mip, antialias, antialias_construct_topology_hash, RasterizeGLContext)

__all__ = ["RasterizeCudaContext", "get_log_level", "set_log_level", "rasterize", "DepthPeeler", "interpolate", "texture",
           "texture_construct_mip", "antialias", "antialias_construct_topology_hash", "RasterizeGLContext"]
