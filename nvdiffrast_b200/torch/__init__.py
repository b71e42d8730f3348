"""Same public names as nvdiffrast.torch (nvdiffrast/torch/__init__.py:9-10), plus the fusion switch of this implementation."""
from .ops import (RasterizeCudaContext, get_log_level, set_log_level, rasterize, DepthPeeler, interpolate, texture,
                  texture_construct_mip, antialias, antialias_construct_topology_hash, RasterizeGLContext, set_fusion, get_fusion)

__all__ = ["RasterizeCudaContext", "get_log_level", "set_log_level", "rasterize", "DepthPeeler", "interpolate", "texture",
           "texture_construct_mip", "antialias", "antialias_construct_topology_hash", "RasterizeGLContext", "set_fusion", "get_fusion"]
