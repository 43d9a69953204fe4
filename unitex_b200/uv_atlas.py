"""UV layout for meshes that come without one.

The reference unwraps such meshes with a chain of CPU third-party tools (`geometry/uv/uv_atlas.py:131-175`: open3d manifold
clean-up, quadric decimation / loop subdivision, Laplacian smoothing, Microsoft UVAtlas through `compute_uvatlas`) that is out of
scope of the B200 hot path and absent from this image.  So that a UV-less mesh still goes through the drop-in call, this module
lays out the SIMPLEST valid atlas: every triangle gets its own right-isosceles slot, two slots per square cell of a regular grid,
separated by gutters.  It is not the reference's atlas (charts, low stretch, shared seams): texel density per face is uniform
instead of proportional to area, and every edge is a seam -- the bake does not care (it works texel by texel in 3-D: visibility,
nearest-neighbour fill and pull-push are all defined per texel), the exported texture is simply less economical.
"""
from __future__ import annotations

import math
from typing import Tuple

import numpy as np


def per_triangle_atlas(n_faces: int, resolution: int = 2048, gutter: int = 1) -> Tuple[np.ndarray, np.ndarray]:
    """-> (UV [3 n_faces, 2] float32 in [0, 1], F_uv [n_faces, 3] int32).  Face f owns UV vertices 3f .. 3f + 2 (no sharing).
    Cells are `s` texels wide, s = resolution // ceil(sqrt(ceil(n_faces / 2))); the two triangles of a cell keep `gutter`
    texels to the cell border and 4 * gutter / sqrt(2) texels between their hypotenuses, so no two triangles touch a common
    texel centre.  Raises when the cells would be smaller than 6 texels (decimate the mesh first: the reference caps UV-less
    meshes at 200 000 faces, uv_atlas.py:153-156, which still fits a 2048^2 atlas here)."""
    if n_faces <= 0:
        return np.zeros((0, 2), np.float32), np.zeros((0, 3), np.int32)
    n_cells = (n_faces + 1) // 2
    side = int(math.ceil(math.sqrt(n_cells)))
    s = resolution // side
    g = float(gutter)
    if s < 6 * gutter:
        raise ValueError(f"per_triangle_atlas: {n_faces} faces need {side} x {side} cells, i.e. {s} texels per cell at "
                         f"{resolution}^2 -- decimate the mesh (the reference caps UV-less meshes at 200 000 faces)")
    f = np.arange(n_faces)
    c = f // 2
    x0 = (c % side).astype(np.float64) * s
    y0 = (c // side).astype(np.float64) * s
    upper = (f % 2 == 1)
    lo = np.stack([np.stack([x0 + g, y0 + g], -1), np.stack([x0 + s - 3 * g, y0 + g], -1), np.stack([x0 + g, y0 + s - 3 * g], -1)], 1)
    up = np.stack([np.stack([x0 + s - g, y0 + s - g], -1), np.stack([x0 + 3 * g, y0 + s - g], -1), np.stack([x0 + s - g, y0 + 3 * g], -1)], 1)
    tex = np.where(upper[:, None, None], up, lo)                         # [F, 3, 2] texel coordinates, counter-clockwise
    uv = (tex.reshape(-1, 2) / float(resolution)).astype(np.float32)
    return uv, np.arange(3 * n_faces, dtype=np.int32).reshape(n_faces, 3)
