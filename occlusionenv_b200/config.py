"""One frozen record of the reference's raster / reward constants (``createRenderers``,
``environment.py:234-284``; step constants ``environment.py:219,386-392``)."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Tuple

import numpy as np


def fov_proj_scale(fov_deg: float = 60.0, znear: float = 1.0) -> float:
    """``K[0,0]`` of ``FoVPerspectiveCameras()`` (pytorch3d renderer/cameras.py), evaluated in fp32 in
    the library's order: fov*(pi/180), tan(fov/2), max_y = tan*znear, 2*znear/(max_x - min_x)."""
    fov = np.float32(fov_deg) * np.float32(np.pi / 180.0)
    t = np.tan(np.float32(fov / np.float32(2.0)), dtype=np.float32)
    max_y = np.float32(t * np.float32(znear))
    return float(np.float32(np.float32(2.0 * znear) / np.float32(max_y - (-max_y))))


@dataclass(frozen=True)
class RasterConfig:
    image_size: int = 512                        # environment.py:202
    sigma: float = 1e-4                          # BlendParams(sigma=1e-4, gamma=1e-4)   :242
    faces_per_pixel: int = 100                   # :252
    cull_backfaces: bool = True                  # :253, :271
    fov: float = 60.0                            # FoVPerspectiveCameras defaults        :238
    znear: float = 1.0
    light: Tuple[float, float, float] = (2.0, 2.0, -2.0)   # PointLights            :275
    step_size: float = 0.05                      # :219
    done_threshold: float = 0.1                  # :386
    reward_done: float = 5.0                     # :390
    reward_step: float = 0.2                     # :392
    norm_with_object_size: bool = False          # :208
    tile_w: int = 0                              # CTA tile (0 = automatic)
    tile_h: int = 0
    debug_exact: bool = False                    # every pair through the reference-order arithmetic (tests)
    obs_planes: int = 4                          # 4: (N,4,S,S) RGB + depth (reference); 2: grey + depth (compact transport)
    ws_budget_mb: int = 0                        # cap of the rasteriser's per-face scratch in MiB (0 = 8192): envs are
                                                 # rasterised in chunks that fit it, see occl_b200.h

    @property
    def blur_radius(self) -> float:
        # np.log(1. / 1e-4 - 1.) * blend_params.sigma                                   :251
        return float(np.float32(np.log(1.0 / 1e-4 - 1.0) * self.sigma))

    @property
    def proj_scale(self) -> float:
        return fov_proj_scale(self.fov, self.znear)

    @property
    def z_clip(self) -> float:
        # MeshRasterizer: z_clip_value = znear / 2 for perspective cameras
        return self.znear / 2.0
