"""Synthetic meshes with a UV atlas for the bake tests (no mesh library needed)."""
import numpy as np


def uv_sphere(rows=24, cols=48, radius=0.6, center=(0.0, 0.0, 0.0), uv_rect=(0.02, 0.02, 0.96, 0.96)):
    """Lat-long sphere; vertices duplicated along the seam so every 3-D vertex has exactly one UV (faces_2d == faces)."""
    th = np.linspace(0.02, np.pi - 0.02, rows + 1)
    ph = np.linspace(0, 2 * np.pi, cols + 1)
    T, P = np.meshgrid(th, ph, indexing="ij")
    v = np.stack([np.sin(T) * np.cos(P), np.cos(T), np.sin(T) * np.sin(P)], -1).reshape(-1, 3) * radius + np.asarray(center)
    u0, v0, du, dv = uv_rect
    uv = np.stack([u0 + du * (P / (2 * np.pi)), v0 + dv * (T / np.pi)], -1).reshape(-1, 2)
    f = []
    for r in range(rows):
        for c in range(cols):
            a, b = r * (cols + 1) + c, r * (cols + 1) + c + 1
            d, e = a + cols + 1, b + cols + 1
            f += [[a, b, d], [b, e, d]]        # outward-facing
    return v.astype(np.float32), np.asarray(f, np.int32), uv.astype(np.float32)


def two_spheres(rows=20, cols=40):
    """A big sphere and a smaller one partly hidden behind/inside its silhouette: gives occluded (covered, invisible)
    texels for the nearest-neighbour fill.  Atlas: left half / right half."""
    v1, f1, uv1 = uv_sphere(rows, cols, 0.55, (0.0, 0.0, 0.0), (0.02, 0.02, 0.46, 0.96))
    v2, f2, uv2 = uv_sphere(rows // 2, cols // 2, 0.25, (0.62, 0.1, 0.05), (0.52, 0.02, 0.46, 0.96))
    v = np.concatenate([v1, v2])
    f = np.concatenate([f1, f2 + len(v1)])
    uv = np.concatenate([uv1, uv2])
    return v, f, uv * 2.0 - 1.0, f.copy()        # uvs_2d in [-1, 1] like trimesh_to_pbr_mesh (structure_v2.py:287)


def analytic_color(p):
    """Smooth colour field of world position, in [0.1, 0.9]."""
    return 0.5 + 0.4 * np.stack([np.sin(3.1 * p[..., 0] + 0.3), np.cos(2.7 * p[..., 1] - 0.2), np.sin(2.3 * p[..., 2] + 1.1)], -1)


_TEASER = None


def teaser_robot_raw():
    """The reference's own test mesh (test_cases/teaser_robot/inputmesh.obj, BASELINE.json config 4) as parsed from the OBJ:
    (V [269026,3] f32, F [499981,3] i32, UV [268818,2] f32 in [0,1], F_uv [499981,3] i32).  Stored losslessly in
    tests/golden/teaser_robot.npz.xz (tests/golden/make_teaser_fixture.py)."""
    global _TEASER
    if _TEASER is None:
        import io
        import lzma
        from pathlib import Path
        z = np.load(io.BytesIO(lzma.decompress((Path(__file__).resolve().parent / "golden" / "teaser_robot.npz.xz").read_bytes())))
        F = np.cumsum(z["F_delta"].astype(np.int64)).reshape(-1, 3)
        Ft = F + z["Ft_minus_F"]
        _TEASER = (z["V"].view(np.float32), F.astype(np.int32), z["UV"].view(np.float32), Ft.astype(np.int32))
    return tuple(a.copy() for a in _TEASER)


def teaser_robot(scale=0.95):
    """teaser_robot in the frame the bake's cameras assume: bounding box centred, longest side = 2*scale (float64 like
    `preprocess_blank_mesh`, reference pipeline.py:170-179 -> geometry/uv/uv_atlas.py:131-194), uvs_2d = uv*2-1
    (mesh/structure_v2.py:287).  -> (V f32, F i32, uvs_2d f32, F_uv i32)."""
    V, F, UV, Ft = teaser_robot_raw()
    V = V.astype(np.float64)
    lo, hi = V.min(0), V.max(0)
    s = (hi - lo).max() / (2.0 * scale)
    V = (V / s - (lo + hi) / (2.0 * s)).astype(np.float32)
    return V, F, (UV * 2.0 - 1.0).astype(np.float32), Ft
