"""B200-native drop-in for the part of the ``gsplat`` (gsplat-rade fork) namespace that collab-splats imports.

collab-splats reaches the rasterizer at exactly four import sites:
``gsplat.rendering.rasterization``, ``gsplat.cuda._wrapper.fully_fused_projection``,
``gsplat.cuda._wrapper.spherical_harmonics`` and ``gsplat.strategy.DefaultStrategy``.  Put the directory that
contains this package (``collab-splats_b200/``) on ``sys.path`` ahead of any other ``gsplat`` and
``collab_splats/models/*`` run unchanged on the sm_100a kernels in ``librade_b200.so``.
"""

from .cuda._wrapper import (fully_fused_projection, isect_offset_encode, isect_tiles, rasterize_to_pixels,
                            spherical_harmonics)
from .rendering import rasterization
from .strategy import DefaultStrategy, MCMCStrategy

__version__ = "1.5.0+rade.b200"

__all__ = ["rasterization", "fully_fused_projection", "spherical_harmonics", "isect_tiles", "isect_offset_encode",
           "rasterize_to_pixels", "DefaultStrategy", "MCMCStrategy"]
