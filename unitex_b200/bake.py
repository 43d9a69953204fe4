"""Host side of the B200 bake path: cameras, the ray-tracer plug-in, rasterise/interpolate wrappers and
`NVDiffRendererInverse` -- same names and call surface as the reference's TextureTools modules
(texturetools/camera/conversion.py, camera/generator.py, raytracing/__init__.py,
render/nvdiffrast/renderer_inverse.py) with every kernel in libunitex_b200.so.  No nvdiffrast / slangtorch /
torch_kdtree / trimesh.
"""
from __future__ import annotations

import ctypes as C
import math
import os
from typing import Optional, Tuple, Union

import numpy as np
import torch

from . import _lib
from .ops import _p, _stream


# ------------------------------------------------------------------------------------------------ cameras (b2)
def intr_to_proj(intr_mtx: torch.Tensor, near=0.01, far=1000.0, perspective=True) -> torch.Tensor:
    """camera/conversion.py:8-28 (GL projection, y row negated for the rasteriser's row-0-at-y=-1 convention)."""
    proj = torch.zeros((*intr_mtx.shape[:-2], 4, 4), dtype=intr_mtx.dtype, device=intr_mtx.device)
    if perspective:
        proj[..., 0, 0] = 2 * intr_mtx[..., 0, 0]
        proj[..., 1, 1] = 2 * intr_mtx[..., 1, 1]
        proj[..., 2, 2] = -(far + near) / (far - near)
        proj[..., 0, 2] = 2 * intr_mtx[..., 0, 2] - 1
        proj[..., 1, 2] = 2 * intr_mtx[..., 1, 2] - 1
        proj[..., 3, 2] = -1.0
        proj[..., 2, 3] = -2.0 * far * near / (far - near)
    else:
        proj[..., 0, 0] = intr_mtx[..., 0, 0]
        proj[..., 1, 1] = intr_mtx[..., 1, 1]
        proj[..., 2, 2] = -2.0 / (far - near)
        proj[..., 3, 3] = 1.0
        proj[..., 0, 3] = -(2 * intr_mtx[..., 0, 2] - 1)
        proj[..., 1, 3] = -(2 * intr_mtx[..., 1, 2] - 1)
        proj[..., 2, 3] = -(far + near) / (far - near)
    proj[..., 1, :] = -proj[..., 1, :]
    return proj


def c2w_to_w2c(c2w: torch.Tensor) -> torch.Tensor:
    """camera/conversion.py:50-57."""
    w2c = torch.zeros((*c2w.shape[:-2], 4, 4), dtype=c2w.dtype, device=c2w.device)
    w2c[..., :3, :3] = c2w[..., :3, :3].transpose(-1, -2)
    w2c[..., :3, 3:] = -c2w[..., :3, :3].transpose(-1, -2) @ c2w[..., :3, 3:]
    w2c[..., 3, 3] = 1.0
    return w2c


def generate_intrinsics(f_x: float, f_y: float, fov=True, degree=False) -> torch.Tensor:
    """camera/generator.py:93-114."""
    if fov:
        if degree:
            f_x, f_y = math.radians(f_x), math.radians(f_y)
        f_x, f_y = 1 / (2 * math.tan(f_x / 2)), 1 / (2 * math.tan(f_y / 2))
    return torch.as_tensor([[f_x, 0.0, 0.5], [0.0, f_y, 0.5], [0.0, 0.0, 1.0]], dtype=torch.float32)


def generate_box_views_c2ws(radius=2.8) -> torch.Tensor:
    """camera/generator.py:153-185: front(+z), right(+x), back(-z), left(-x), top(+y), down(-y)."""
    r = radius
    return torch.tensor([
        [[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, r], [0, 0, 0, 1]],
        [[0, 0, 1, r], [0, 1, 0, 0], [-1, 0, 0, 0], [0, 0, 0, 1]],
        [[-1, 0, 0, 0], [0, 1, 0, 0], [0, 0, -1, -r], [0, 0, 0, 1]],
        [[0, 0, -1, -r], [0, 1, 0, 0], [1, 0, 0, 0], [0, 0, 0, 1]],
        [[1, 0, 0, 0], [0, 0, 1, r], [0, -1, 0, 0], [0, 0, 0, 1]],
        [[-1, 0, 0, 0], [0, 0, -1, -r], [0, -1, 0, 0], [0, 0, 0, 1]],
    ], dtype=torch.float32)


def generate_orbit_views_c2ws(num_views: int, radius: float = 1.0, height: float = 0.0, theta_0: float = 0.0, degree=False) -> torch.Tensor:
    """camera/generator.py:115-125 over lookat_to_matrix :8-41: cameras on a horizontal circle (world x forward, y right, z up, then
    re-labelled to the renderer's z forward, x right, y up), each looking at the origin; the angle grid includes both end points."""
    if degree:
        theta_0 = math.radians(theta_0)
    pr = math.sqrt(radius ** 2 - height ** 2)
    theta = torch.linspace(theta_0, 2.0 * math.pi + theta_0, num_views, dtype=torch.float32)
    eye = torch.stack([pr * torch.cos(theta), pr * torch.sin(theta), torch.full((num_views,), height, dtype=torch.float32)], dim=-1)
    z_axis = torch.nn.functional.normalize(eye, dim=-1)
    up = torch.tensor([0.0, 0.0, 1.0]).expand_as(z_axis)
    x_axis = torch.linalg.cross(up, z_axis, dim=-1)
    degenerate = (x_axis == 0).all(dim=-1, keepdim=True)                    # looking straight down / up: the reference hard-codes +y
    x_axis = torch.where(degenerate, torch.tensor([0.0, 1.0, 0.0]), x_axis)
    y_axis = torch.linalg.cross(z_axis, x_axis, dim=-1)
    c2w = torch.zeros(num_views, 4, 4)
    c2w[:, :3, 0], c2w[:, :3, 1], c2w[:, :3, 2], c2w[:, :3, 3] = x_axis, y_axis, z_axis, eye
    c2w[:, 3, 3] = 1.0
    relabel = torch.tensor([[0.0, 1.0, 0.0, 0.0], [0.0, 0.0, 1.0, 0.0], [1.0, 0.0, 0.0, 0.0], [0.0, 0.0, 0.0, 1.0]])
    return torch.matmul(relabel, c2w)


# ------------------------------------------------------------------------------------------------ raster wrappers
def _f32(t, device):
    return t.to(device=device, dtype=torch.float32).contiguous()


def rasterize(pos: torch.Tensor, tri: torch.Tensor, resolution: Tuple[int, int]) -> torch.Tensor:
    """dr.rasterize: pos [B,V,4] (or [1,V,4] shared), tri [F,3] int32 -> [B,H,W,4] (u, v, z/w, id+1)."""
    L = _lib.load()
    H, W = resolution
    pos = pos.contiguous()
    B, V = pos.shape[0], pos.shape[1]
    tri = tri.to(torch.int32).contiguous()
    out = torch.empty(B, H, W, 4, device=pos.device, dtype=torch.float32)
    ws = torch.empty(L.utx_rasterize_workspace_bytes(B, H, W, tri.shape[0]), device=pos.device, dtype=torch.uint8)
    _lib.check(L.utx_rasterize(_p(pos), 1, V, _p(tri), tri.shape[0], B, H, W, _p(out), _p(ws), _stream()), "utx_rasterize")
    return out


def interpolate(attr: torch.Tensor, rast: torch.Tensor, tri: torch.Tensor) -> torch.Tensor:
    """dr.interpolate: attr [V,C] or [B,V,C]; rast [B,H,W,4] -> [B,H,W,C]."""
    L = _lib.load()
    batched = attr.dim() == 3
    attr = attr.contiguous()
    V, Cn = attr.shape[-2], attr.shape[-1]
    B, H, W, _ = rast.shape
    tri = tri.to(torch.int32).contiguous()
    out = torch.empty(B, H, W, Cn, device=rast.device, dtype=torch.float32)
    _lib.check(L.utx_interpolate(_p(attr), int(batched), V, Cn, _p(rast.contiguous()), _p(tri), B, H, W, _p(out), _stream()),
               "utx_interpolate")
    return out


def transform_points(vertices: torch.Tensor, mats: torch.Tensor) -> torch.Tensor:
    """[V,3] x [n,4,4] -> clip [n,V,4]  (vertices_homo @ M^T, renderer_inverse.py:177-178)."""
    L = _lib.load()
    vertices, mats = vertices.contiguous(), mats.contiguous()
    n, V = mats.shape[0], vertices.shape[0]
    out = torch.empty(n, V, 4, device=vertices.device, dtype=torch.float32)
    _lib.check(L.utx_transform_points(_p(vertices), V, _p(mats), n, _p(out), _stream()), "utx_transform_points")
    return out


# ------------------------------------------------------------------------------------------------ ray tracer plug-in (b4)
class RayTracing:
    """Drop-in for texturetools.raytracing.RayTracing (raytracing/__init__.py:12-80); one backend: the B200 LBVH."""

    def __init__(self, vertices: torch.Tensor, faces: torch.Tensor, backend: Optional[str] = None, device="cuda"):
        self.device = torch.device(device)
        self.update_raw(vertices, faces)

    def update_raw(self, vertices: torch.Tensor, faces: torch.Tensor):
        L = _lib.load()
        self.vertices = _f32(vertices, self.device)
        self.faces = faces.to(device=self.device, dtype=torch.int32).contiguous()
        F = self.faces.shape[0]
        self.nodes = torch.empty(L.utx_bvh_nodes_bytes(F), device=self.device, dtype=torch.uint8)
        ws = torch.empty(L.utx_bvh_workspace_bytes(F), device=self.device, dtype=torch.uint8)
        _lib.check(L.utx_bvh_build(_p(self.vertices), self.vertices.shape[0], _p(self.faces), F, _p(self.nodes), _p(ws),
                                   ws.numel(), _stream()), "utx_bvh_build")

    def export(self):
        """(info [2F-1,3] int32, aabb [2F-1,6] fp32) in the reference's LBVHNode layout."""
        L = _lib.load()
        n = 2 * self.faces.shape[0] - 1
        info = torch.empty(n, 3, device=self.device, dtype=torch.int32)
        aabb = torch.empty(n, 6, device=self.device, dtype=torch.float32)
        _lib.check(L.utx_bvh_export(_p(self.nodes), self.faces.shape[0], _p(info), _p(aabb), _stream()), "utx_bvh_export")
        return info, aabb

    def intersects_closest(self, rays_o: torch.Tensor, rays_d: torch.Tensor):
        """-> (hit bool[...], front None, tri_idx int64[...] (-1 = miss), loc [...,3], uv [...,2])"""
        L = _lib.load()
        rays_o, rays_d = torch.broadcast_tensors(rays_o, rays_d)
        shape = rays_o.shape[:-1]
        o = _f32(rays_o, self.device).reshape(-1, 3)
        d = _f32(rays_d, self.device).reshape(-1, 3)
        N = o.shape[0]
        hit = torch.empty(N, device=self.device, dtype=torch.uint8)
        tid = torch.empty(N, device=self.device, dtype=torch.int32)
        loc = torch.empty(N, 3, device=self.device, dtype=torch.float32)
        uv = torch.empty(N, 2, device=self.device, dtype=torch.float32)
        _lib.check(L.utx_bvh_intersect(_p(self.nodes), _p(self.vertices), _p(self.faces), self.faces.shape[0], _p(o), _p(d), N, _p(hit), _p(tid),
                                       _p(loc), _p(uv), _stream()), "utx_bvh_intersect")
        return hit.bool().reshape(shape), None, tid.to(torch.int64).reshape(shape), loc.reshape(*shape, 3), uv.reshape(*shape, 2)


def knn(src: torch.Tensor, dst: torch.Tensor, k: int = 1, backend: Optional[str] = None, batch_size: Optional[int] = None,
        device="cuda"):
    """Drop-in for texturetools.pcd.knn (pcd/knn/__init__.py:104-114): -> (score [M,k] fp32 distance, index [M,k] int64),
    exact, rows ascending by (distance, index).  k = 1 is the reproject fill (renderer_inverse.py:611); k = 32 / 8+1 the
    kdtree bake (:385,:413,:427).  `backend` / `batch_size` are accepted and ignored (one backend, no batching needed)."""
    L = _lib.load()
    dev = torch.device(device)
    s, d = _f32(src, dev), _f32(dst, dev)
    n, M = s.shape[0], d.shape[0]
    if not 1 <= k <= 32:
        raise ValueError("knn: k must be in 1..32")
    if k > n:
        raise ValueError("knn: k exceeds the number of source points")
    index = torch.empty(M, k, device=dev, dtype=torch.int64)
    score = torch.empty(M, k, device=dev, dtype=torch.float32)
    nodes = torch.empty(max(L.utx_bvh_nodes_bytes(max(n, 2)), 64), device=dev, dtype=torch.uint8)
    ws = torch.empty(L.utx_bvh_workspace_bytes(max(n, 2)), device=dev, dtype=torch.uint8)
    _lib.check(L.utx_knn(_p(s), n, _p(d), M, k, _p(index), _p(score), _p(nodes), _p(ws), ws.numel(), _stream()), "utx_knn")
    return score, index


# ------------------------------------------------------------------------------------------------ lens-blur kernel (b9)
_LENS5 = [[4.892608, 1.685979, -22.356787, 85.91246], [4.71187, 4.998496, 35.918936, -28.875618],
          [4.052795, 8.244168, -13.212253, -1.578428], [2.929212, 11.900859, 0.507991, 1.816328],
          [1.512961, 16.116382, 0.138051, -0.01]]   # yehar.com 5-component parameters (image/lens_blur.py:45-50)


def lens_blur_kernel_2d(radius: float = 3.0, scale: float = 1.2) -> np.ndarray:
    """Effective real 2-D kernel of lens_blur_torch(radius=3, components=5) (image/lens_blur.py:82-93,109-121,172-195,
    260-280): K[i,j] = sum_c A_c Re(k_c[i] k_c[j]) + B_c Im(k_c[i] k_c[j]), all components normalised together."""
    size = int(math.ceil(radius)) * 2 + 1
    ax = np.linspace(-radius, radius, size, dtype=np.float32).astype(np.float64) * scale * (1 / radius)
    ks = [np.exp(-a * ax ** 2) * (np.cos(b * ax ** 2) + 1j * np.sin(b * ax ** 2)) for a, b, _, _ in _LENS5]
    total = 0.0
    for k, (_, _, A, B) in zip(ks, _LENS5):
        kk = np.outer(k, k)
        total += (A * kk.real + B * kk.imag).sum()
    ks = [k / math.sqrt(total) for k in ks]
    K = np.zeros((size, size))
    for k, (_, _, A, B) in zip(ks, _LENS5):
        kk = np.outer(k, k)          # [vertical tap, horizontal tap]
        K += A * kk.real + B * kk.imag
    return K.astype(np.float32)


def gaussian_kernel_2d(kernel_size: int = 5) -> torch.Tensor:
    """The kernel of image/gaussian_blur.py:42-44 -> torchvision gaussian_blur [ext, present in this image] with its default
    sigma = 0.3 ((k - 1) / 2 - 1) + 0.8, embedded in the 7x7 frame the seam-blur kernel reads (fp32 like torchvision's)."""
    sigma = 0.3 * ((kernel_size - 1) * 0.5 - 1) + 0.8
    half = (kernel_size - 1) * 0.5
    x = torch.linspace(-half, half, steps=kernel_size, dtype=torch.float32)
    pdf = torch.exp(-0.5 * (x / sigma).pow(2))
    k1 = pdf / pdf.sum()
    k2 = torch.mm(k1[:, None], k1[None, :])
    out = torch.zeros(7, 7, dtype=torch.float32)
    o = (7 - kernel_size) // 2
    out[o:o + kernel_size, o:o + kernel_size] = k2
    return out.contiguous()


def area_weighted_vertex_normals(vertices: torch.Tensor, faces: torch.Tensor) -> torch.Tensor:
    """PBRMesh.vertex_normals (mesh/structure_v2.py:49,63-71): each face's area vector is scattered into one slot per triangle
    corner, the three slots are averaged, the result normalised."""
    f = faces.long()
    areas = torch.linalg.cross(vertices[f[:, 1]] - vertices[f[:, 0]], vertices[f[:, 2]] - vertices[f[:, 0]], dim=-1)
    vn = torch.zeros(vertices.shape[0], 3, 3, dtype=vertices.dtype, device=vertices.device)
    vn.scatter_add_(0, f.unsqueeze(-1).expand(-1, -1, 3), areas.unsqueeze(1).expand(-1, 3, -1))
    return torch.nn.functional.normalize(vn.mean(dim=1), dim=-1)


# ------------------------------------------------------------------------------------------------ mesh container (b3)
class BakeMesh:
    """The fields of PBRMesh the bake reads (mesh/structure_v2.py:25-77): vertices [V,3], faces [F,3], uvs_2d [V2,2] in
    [-1,1] (uv*2-1, :287), faces_2d [F,3]; lazy LBVH like PBRMesh.optix."""

    def __init__(self, vertices, faces, uvs_2d, faces_2d, device="cuda"):
        self.device = torch.device(device)
        self.vertices = _f32(torch.as_tensor(vertices), self.device)
        self.faces = torch.as_tensor(faces).to(device=self.device, dtype=torch.int32).contiguous()
        self.uvs_2d = _f32(torch.as_tensor(uvs_2d), self.device)
        self.faces_2d = torch.as_tensor(faces_2d).to(device=self.device, dtype=torch.int32).contiguous()
        assert self.faces.shape == self.faces_2d.shape
        self._optix = None
        self._normals = None
        self._vertex_normals = None

    @property
    def areas(self) -> torch.Tensor:
        """Face area vectors, mesh/structure_v2.py:49."""
        f = self.faces.long()
        v = self.vertices
        return torch.linalg.cross(v[f[:, 1]] - v[f[:, 0]], v[f[:, 2]] - v[f[:, 0]], dim=-1)

    @property
    def normals(self) -> torch.Tensor:
        """Face normals, :50."""
        if self._normals is None:
            self._normals = torch.nn.functional.normalize(self.areas, dim=-1)
        return self._normals

    @property
    def vertex_normals(self) -> torch.Tensor:
        """Area-weighted vertex normals, :63-71 (only the gradient filter of mv_to_pcd reads them)."""
        if self._vertex_normals is None:
            # once per mesh, on the host: a sequential scatter is reproducible run to run (atomics on the device are not), and the
            # gradient filter thresholds quantities derived from these normals
            self._vertex_normals = area_weighted_vertex_normals(self.vertices.cpu(), self.faces.cpu()).to(self.device)
        return self._vertex_normals

    @property
    def optix(self) -> RayTracing:
        if self._optix is None:
            self._optix = RayTracing(self.vertices, self.faces, device=self.device)
        return self._optix


def _load_obj_slow(path: str):
    v, vt, f, ft = [], [], [], []
    with open(path) as fh:
        for line in fh:
            if line.startswith("v "):
                v.append([float(x) for x in line.split()[1:4]])
            elif line.startswith("vt "):
                vt.append([float(x) for x in line.split()[1:3]])
            elif line.startswith("f "):
                c = [p.split("/") for p in line.split()[1:]]
                vi = [int(p[0]) for p in c]
                ti = [int(p[1]) if len(p) > 1 and p[1] else 0 for p in c]
                for k in range(1, len(vi) - 1):
                    f.append([vi[0], vi[k], vi[k + 1]])
                    ft.append([ti[0], ti[k], ti[k + 1]])
    return (np.asarray(v, np.float32).reshape(-1, 3), np.asarray(vt, np.float32).reshape(-1, 2),
            np.asarray(f, np.int64).reshape(-1, 3), np.asarray(ft, np.int64).reshape(-1, 3))


def load_obj(path: str):
    """Minimal OBJ reader (v / vt / f with v[/vt[/vn]] corners, negative indices, polygons fan-triangulated)
    -> (V[,3], F[,3], UV[,2], F_uv[,3]).  Files whose faces are all triangles of one corner format (what MeshLab / open3d /
    trimesh write, e.g. the reference's 500 k-face teaser mesh) are parsed in bulk with numpy; anything else line by line."""
    with open(path, "rb") as fh:
        lines = fh.read().split(b"\n")
    vl = [ln for ln in lines if ln[:2] == b"v "]
    tl = [ln for ln in lines if ln[:3] == b"vt "]
    fl = [ln for ln in lines if ln[:2] == b"f "]
    V = UV = Fv = Ft = None
    try:
        if not vl or not fl:
            raise ValueError
        nv = len(vl[0].split()) - 1
        V = np.array(b" ".join(ln[2:] for ln in vl).split(), dtype=np.float64).reshape(len(vl), nv)[:, :3].astype(np.float32)
        if tl:
            nt = len(tl[0].split()) - 1
            UV = np.array(b" ".join(ln[3:] for ln in tl).split(), dtype=np.float64).reshape(len(tl), nt)[:, :2].astype(np.float32)
        else:
            UV = np.zeros((0, 2), np.float32)
        first = fl[0].split()[1:]
        per_corner = first[0].count(b"/") + 1
        if len(first) != 3 or b"//" in fl[0]:
            raise ValueError
        tok = np.array(b" ".join(ln[2:] for ln in fl).replace(b"/", b" ").split(), dtype=np.int64)
        if tok.size != len(fl) * 3 * per_corner:
            raise ValueError                                   # mixed polygons / corner formats
        tok = tok.reshape(len(fl), 3, per_corner)
        Fv = tok[:, :, 0]
        Ft = tok[:, :, 1] if per_corner > 1 else np.zeros_like(Fv)
    except ValueError:
        V, UV, Fv, Ft = _load_obj_slow(path)
    Fv = np.where(Fv < 0, Fv + len(V) + 1, Fv) - 1
    Ft = np.where(Ft < 0, Ft + len(UV) + 1, Ft) - 1
    return V, Fv.astype(np.int32), UV, Ft.astype(np.int32)


def load_glb(path: str):
    """Minimal binary glTF 2.0 reader for blank / textured input meshes (the reference's second test case ships `.glb`;
    trimesh / open3d load them there: io/mesh_loader.py:22-30, geometry/uv/uv_atlas.py:181): every triangle primitive of every
    mesh node of the default scene, node transforms (matrix or TRS, nested) applied, primitives concatenated.
    -> (V[,3], F[,3], UV[,2], F_uv[,3]) with UV in the OBJ convention (v up: glTF's v axis points down, so v -> 1 - v);
    UV is empty when a primitive has no TEXCOORD_0."""
    import json
    import struct
    with open(path, "rb") as fh:
        blob = fh.read()
    magic, version, _ = struct.unpack_from("<III", blob, 0)
    if magic != 0x46546C67 or version != 2:
        raise ValueError(f"{path}: not a binary glTF 2.0 file")
    off, gltf, binc = 12, None, b""
    while off < len(blob):
        clen, ctype = struct.unpack_from("<II", blob, off)
        data = blob[off + 8: off + 8 + clen]
        if ctype == 0x4E4F534A:
            gltf = json.loads(data)
        elif ctype == 0x004E4942:
            binc = data
        off += 8 + clen
    comp = {5120: np.int8, 5121: np.uint8, 5122: np.int16, 5123: np.uint16, 5125: np.uint32, 5126: np.float32}
    width = {"SCALAR": 1, "VEC2": 2, "VEC3": 3, "VEC4": 4, "MAT4": 16}

    def accessor(i):
        a = gltf["accessors"][i]
        bv = gltf["bufferViews"][a["bufferView"]]
        dt, w = np.dtype(comp[a["componentType"]]), width[a["type"]]
        start = bv.get("byteOffset", 0) + a.get("byteOffset", 0)
        stride = bv.get("byteStride", 0)
        if stride and stride != dt.itemsize * w:
            raw = np.frombuffer(binc, np.uint8, count=stride * (a["count"] - 1) + dt.itemsize * w, offset=start)
            rows = np.lib.stride_tricks.as_strided(raw, (a["count"], dt.itemsize * w), (stride, 1))
            return np.ascontiguousarray(rows).view(dt).reshape(a["count"], w)
        return np.frombuffer(binc, dt, count=a["count"] * w, offset=start).reshape(a["count"], w)

    def local_matrix(node):
        if "matrix" in node:
            return np.asarray(node["matrix"], np.float64).reshape(4, 4).T          # glTF stores column-major
        m = np.eye(4)
        x, y, z, w = node.get("rotation", [0.0, 0.0, 0.0, 1.0])
        r = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                      [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                      [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
        m[:3, :3] = r * np.asarray(node.get("scale", [1.0, 1.0, 1.0]))[None, :]
        m[:3, 3] = node.get("translation", [0.0, 0.0, 0.0])
        return m

    Vs, Fs, UVs, have_uv, base = [], [], [], True, 0

    def visit(ni, parent):
        nonlocal base, have_uv
        node = gltf["nodes"][ni]
        world = parent @ local_matrix(node)
        if "mesh" in node:
            for prim in gltf["meshes"][node["mesh"]]["primitives"]:
                if prim.get("mode", 4) != 4:
                    continue
                pos = accessor(prim["attributes"]["POSITION"]).astype(np.float64)
                pos = pos @ world[:3, :3].T + world[:3, 3]
                idx = accessor(prim["indices"]).reshape(-1) if "indices" in prim else np.arange(len(pos))
                Vs.append(pos.astype(np.float32))
                Fs.append(idx.astype(np.int64).reshape(-1, 3) + base)
                if "TEXCOORD_0" in prim["attributes"]:
                    uv = accessor(prim["attributes"]["TEXCOORD_0"]).astype(np.float32).copy()
                    uv[:, 1] = 1.0 - uv[:, 1]
                    UVs.append(uv)
                else:
                    have_uv = False
                base += len(pos)
        for c in node.get("children", []):
            visit(c, world)

    scene = gltf["scenes"][gltf.get("scene", 0)]
    for ni in scene["nodes"]:
        visit(ni, np.eye(4))
    if not Vs:
        raise ValueError(f"{path}: no triangle primitives")
    V, F = np.concatenate(Vs), np.concatenate(Fs).astype(np.int32)
    if have_uv:
        return V, F, np.concatenate(UVs), F.copy()
    return V, F, np.zeros((0, 2), np.float32), np.zeros((0, 3), np.int32)


_MESH_CACHE: dict = {}


def load_mesh(path: str):
    """OBJ or GLB by extension -> (V, F, UV, F_uv).  The last few files are kept parsed, keyed by (path, mtime, size): the
    pipeline stages hand each other file paths (reference pipeline.py:568-575), so the same mesh is opened four times per asset."""
    ext = os.path.splitext(path)[1].lower()
    if ext not in (".obj", ".glb"):
        raise NotImplementedError(f"mesh format {ext!r}: .obj and .glb are read")
    st = os.stat(path)
    key = (os.path.abspath(path), st.st_mtime_ns, st.st_size)
    hit = _MESH_CACHE.get(key)
    if hit is None:
        hit = load_obj(path) if ext == ".obj" else load_glb(path)
        if len(_MESH_CACHE) >= 4:
            _MESH_CACHE.pop(next(iter(_MESH_CACHE)))
        _MESH_CACHE[key] = hit
    return tuple(a.copy() for a in hit)       # callers own their arrays (a few MB; the parse is what costs seconds)


# ------------------------------------------------------------------------------------------------ NVDiffRendererInverse (b5-b7)
class NVDiffRendererInverse:
    """Drop-in for render/nvdiffrast/renderer_inverse.py::NVDiffRendererInverse: `infer` with method 'reproject' (lens blur) and
    'kdtree' ('order_mean' | 'mean'), orthographic or perspective views, RGB or 9-channel PBR attributes, the gradient filter of
    mv_to_pcd, and the `*_inpainting` query-field hook."""

    def __init__(self, device="cuda", pbr_mesh: Optional[BakeMesh] = None):
        self.device = torch.device(device)
        self.pbr_mesh = pbr_mesh
        self.index = [0, 3, 4, 1, 2, 5]      # frtbld ==> fblrtd   (renderer_inverse.py:44)
        self.query_field_function = None
        self._k2d = torch.from_numpy(lens_blur_kernel_2d()).to(self.device).contiguous()
        self.last_nn_index = None

    def clear(self):
        self.pbr_mesh = None
        self.query_field_function = None

    def update_from_file(self, path: str):
        V, F, UV, Ft = load_mesh(path)
        self.pbr_mesh = BakeMesh(V, F, UV * 2.0 - 1.0, Ft, device=self.device)
        return self

    def register_query_field(self, fn):
        self.query_field_function = fn

    def _view_mats(self, c2ws, intrinsics, perspective):
        return torch.matmul(intr_to_proj(intrinsics.float().cpu(), perspective=perspective), c2w_to_w2c(c2ws.float().cpu()))

    def mv_to_pcd(self, c2ws, intrinsics, render_size, image_attrs=None, perspective=True, grad_norm_threhold=0.20,
                  ray_normal_angle_threhold=115.0, filt_gradient_points=False):
        """:159-241, the raster masks of the views.  filt_gradient_points=False (what pipeline.py:343-347 passes): the visible
        alpha is the coverage mask.  True (the default of the reference's `infer`): a pixel also has to face its ray and lie
        in a stretch where the screen-space gradient of (position, vertex normal) stays below `grad_norm_threhold` (:188-214)."""
        H, W = (render_size, render_size) if isinstance(render_size, int) else render_size
        m = self.pbr_mesh
        n = c2ws.shape[0]
        mats = self._view_mats(c2ws, intrinsics, perspective).to(self.device)
        clip = transform_points(m.vertices, mats)
        rast = rasterize(clip, m.faces, (H, W))
        mask = rast[..., 3:4] > 0
        mask_vis = mask
        if filt_gradient_points:
            # one kernel (bake_filter.cu): torch.gradient of the interpolated attributes, ray / face-normal cosine and the
            # reference's MaxPool2d(31, 1, 15) erosion, which on its [n,H,W,1] tensor runs along the image x axis only (:204-205)
            attrs = interpolate(torch.cat([m.vertices, m.vertex_normals], dim=-1).contiguous(), rast, m.faces)
            c2 = c2ws.to(self.device, torch.float32)
            view_dirs = (c2[:, :3, 3] if perspective else c2[:, :3, 2].neg()).contiguous()
            vis = torch.empty(n, H, W, 1, device=self.device, dtype=torch.uint8)
            face_n = m.normals.contiguous().float()
            _lib.check(_lib.load().utx_mv_visibility_filter(
                _p(attrs), _p(rast), _p(face_n), _p(view_dirs), int(bool(perspective)), n, H, W,
                float(grad_norm_threhold), math.cos(math.radians(ray_normal_angle_threhold)), _p(vis), _stream()),
                "utx_mv_visibility_filter")
            mask_vis = vis.bool()
        return {"mask": mask, "alpha": mask.float(), "mask_visiable": mask_vis, "alpha_visiable": mask_vis.float(), "rast": rast}

    def query_field(self, vertices_visiable, colors_visiable, vertices_invisiable):
        """:139-154."""
        if self.query_field_function is None:
            raise NotImplementedError("using register_query_field before query")
        return self.query_field_function(vertices_visiable, colors_visiable, vertices_invisiable)

    def infer(self, blank_mesh, c2ws: torch.Tensor, intrinsics: torch.Tensor, image_attrs: torch.Tensor, H=512, W=512,
              H2D=2048, W2D=2048, perspective=True, grad_norm_threhold=0.20, ray_normal_angle_threhold=115.0,
              grid_interpolate_mode="torch", method="reproject", kdtree_n_neighbors=32, kdtree_n_neighbors_visiable=1,
              kdtree_n_neighbors_invisiable=32, kdtree_method="order_mean", kdtree_inpainting=False,
              reproject_method="lens", reproject_kernel_size_boundary=3, reproject_kernel_size_boundary_blur=3,
              reproject_kernel_size_blur=5, reproject_inpainting=False, return_mv_reproject_uv=False,
              filt_gradient_points=True):
        """:635-726.  Returns (textured_mesh, mask_2d_visiable [n,H2D,W2D,1] bool, mask_2d [1,H2D,W2D,1] bool,
        color_2d [1,H2D,W2D,3] fp32).  textured_mesh is left to the caller's exporter (io layer).
        method='reproject' (reproject_method='lens') is the path CustomRGBTextureFullPipeline uses (pipeline.py:335-348);
        method='kdtree' (kdtree_method 'order_mean' | 'mean') is bake_mv_to_uv_kdtree (:367-433); `*_inpainting=True` routes the
        uncovered texels through the registered query field (:387-389, :427-432, :609-614)."""
        assert method in ("kdtree", "reproject")
        assert image_attrs.shape[-1] in (3, 9)
        if image_attrs.shape[-1] == 9:
            # PBR attributes (albedo | metallic-roughness | bump, :711-719): visibility, ownership, seams and neighbour indices do
            # not depend on the colours and every colour stage is per channel, so the 9-channel bake is three RGB bakes.
            if kdtree_inpainting or reproject_inpainting:
                raise NotImplementedError("9-channel attributes with a query field: the field would see 3-channel colours")
            kw = dict(H=H, W=W, H2D=H2D, W2D=W2D, perspective=perspective, grad_norm_threhold=grad_norm_threhold,
                      ray_normal_angle_threhold=ray_normal_angle_threhold, grid_interpolate_mode=grid_interpolate_mode, method=method,
                      kdtree_n_neighbors=kdtree_n_neighbors, kdtree_n_neighbors_visiable=kdtree_n_neighbors_visiable,
                      kdtree_n_neighbors_invisiable=kdtree_n_neighbors_invisiable, kdtree_method=kdtree_method,
                      reproject_method=reproject_method, reproject_kernel_size_boundary=reproject_kernel_size_boundary,
                      reproject_kernel_size_boundary_blur=reproject_kernel_size_boundary_blur,
                      reproject_kernel_size_blur=reproject_kernel_size_blur, filt_gradient_points=filt_gradient_points)
            parts = [self.infer(blank_mesh, c2ws, intrinsics, image_attrs[..., 3 * i:3 * i + 3], **kw) for i in range(3)]
            return None, parts[0][1], parts[0][2], torch.cat([p[3] for p in parts], dim=-1)
        if method == "reproject":
            assert reproject_method in ("gaussian", "lens")
            if (reproject_kernel_size_boundary, reproject_kernel_size_boundary_blur) != (3, 3):
                raise NotImplementedError("reproject bake: the seam mask is built for the 3x3 boundary kernels (the defaults, :649-650)")
            if reproject_method == "gaussian" and reproject_kernel_size_blur not in (3, 5, 7):
                raise NotImplementedError("reproject bake: gaussian kernel sizes 3, 5, 7")
        if method == "kdtree":
            assert kdtree_method in ("mean", "mvpaint", "order_mean")
        if isinstance(blank_mesh, str):
            self.update_from_file(blank_mesh)
        elif isinstance(blank_mesh, BakeMesh):
            self.pbr_mesh = blank_mesh
        m = self.pbr_mesh
        L = _lib.load()
        n = c2ws.shape[0]
        assert n == len(self.index), "the reference bake is hard-wired to 6 views (renderer_inverse.py:171,256,589)"
        mv = self.mv_to_pcd(c2ws, intrinsics, (H, W), perspective=perspective, grad_norm_threhold=grad_norm_threhold,
                            ray_normal_angle_threhold=ray_normal_angle_threhold, filt_gradient_points=filt_gradient_points)
        rgba = torch.cat([_f32(image_attrs, self.device), mv["alpha_visiable"]], dim=-1).contiguous()
        uv_clip = torch.cat([m.uvs_2d, torch.zeros_like(m.uvs_2d[:, :1]), torch.ones_like(m.uvs_2d[:, :1])], dim=-1)[None]
        rast2d = rasterize(uv_clip, m.faces_2d, (H2D, W2D))
        mats = self._view_mats(c2ws, intrinsics, perspective).contiguous()
        # orthographic: the common ray direction -c2w[:3, 2]; perspective: the camera position the rays leave from (:279-284)
        dirs = (c2ws[:, :3, 3] if perspective else -c2ws[:, :3, 2]).float().cpu().contiguous()
        prio = (C.c_int32 * n)(*self.index)
        T = H2D * W2D
        mask2d = torch.empty(T, device=self.device, dtype=torch.uint8)
        mask_vis = torch.empty(n, T, device=self.device, dtype=torch.uint8)
        color = torch.empty(T, 3, device=self.device, dtype=torch.float32)
        nn_index = torch.empty(T, device=self.device, dtype=torch.int32)
        ws = torch.empty(L.utx_uv_bake_workspace_bytes(H2D, W2D), device=self.device, dtype=torch.uint8)
        cos_t = float(np.float32(math.cos(math.radians(ray_normal_angle_threhold))))
        self.last_rast2d = rast2d

        def out():      # built AFTER the launches: .bool() copies
            return (None, mask_vis.bool().reshape(n, H2D, W2D, 1), mask2d.bool().reshape(1, H2D, W2D, 1),
                    color.reshape(1, H2D, W2D, 3))

        vis_args = (_p(m.vertices), m.vertices.shape[0], _p(m.faces), m.faces.shape[0], _p(m.optix.nodes), _p(rast2d), H2D, W2D,
                    n, mats.numpy().ctypes.data_as(_lib.fp), dirs.numpy().ctypes.data_as(_lib.fp), int(bool(perspective)), prio, _p(rgba), H, W, cos_t)
        if method == "reproject" and reproject_method == "lens" and not reproject_inpainting:
            lo_arr = (C.c_float * 3)(0.0, 0.0, 0.0)
            _lib.check(L.utx_uv_bake(*vis_args, _p(self._k2d), 5.0, lo_arr, 1.0, _p(mask2d), _p(mask_vis), _p(color),
                                     _p(nn_index), _p(ws), ws.numel(), _stream()), "utx_uv_bake")
            self.last_nn_index = nn_index
            return out()
        # staged form: visibility -> [per-view k-NN colours] -> fill (k-NN or the caller's query field) -> finish
        _lib.check(L.utx_uv_bake_visibility(*vis_args, _p(mask2d), _p(mask_vis), _p(ws), ws.numel(), _stream()),
                   "utx_uv_bake_visibility")
        self.last_nn_index = None
        skip_fill = False
        if method == "kdtree":
            if kdtree_method == "mvpaint":
                self._mvpaint_fill(ws, mask2d, rgba, mv["rast"], rast2d, H2D, W2D, kdtree_n_neighbors)
            elif kdtree_method == "mean" and kdtree_inpainting:
                self._field_fill(ws, mask2d, rgba, mv["rast"], H2D, W2D, union_cloud=True)
            else:
                merge = int(kdtree_method == "mean")
                k = kdtree_n_neighbors if merge else kdtree_n_neighbors_visiable
                pix_pos = interpolate(m.vertices, mv["rast"], m.faces)
                scratch = torch.empty(L.utx_uv_bake_views_workspace_bytes(n, H, W), device=self.device, dtype=torch.uint8)
                _lib.check(L.utx_uv_bake_views_knn(_p(pix_pos), _p(rgba), n, H, W, k, merge, _p(mask2d), H2D, W2D, _p(ws),
                                                   ws.numel(), _p(scratch), scratch.numel(), _stream()), "utx_uv_bake_views_knn")
            skip_fill = kdtree_method in ("mean", "mvpaint")
            inpaint, k_fill, blur = kdtree_inpainting, kdtree_n_neighbors_invisiable, 0
        else:
            inpaint, k_fill, blur = reproject_inpainting, 1, (1 if reproject_method == "lens" else 2)
        k2d, gamma = (self._k2d, 5.0) if blur != 2 else (gaussian_kernel_2d(reproject_kernel_size_blur).to(self.device), 1.0)
        if not skip_fill:
            if inpaint:
                self._field_fill(ws, mask2d, rgba, mv["rast"], H2D, W2D, union_cloud=False)
            else:
                _lib.check(L.utx_uv_bake_fill(_p(mask2d), H2D, W2D, k_fill, _p(nn_index), _p(ws), ws.numel(), _stream()),
                           "utx_uv_bake_fill")
                self.last_nn_index = nn_index
        _lib.check(L.utx_uv_bake_finish(_p(mask2d), H2D, W2D, blur, _p(k2d), gamma, _p(color), _p(ws), ws.numel(), _stream()),
                   "utx_uv_bake_finish")
        return out()

    def _mvpaint_fill(self, ws, mask2d, rgba, rast_mv, rast2d, H2D, W2D, k: int):
        """kdtree_method='mvpaint' (:390-399; MVPaint, arXiv 2411.02336 sec. 3.2): every covered texel takes its k nearest points of
        the union pixel cloud, weighted by normalised inverse distance x cosine between the face normals of point and texel.  The
        neighbour search is `utx_knn`, the weighting of the [M, k] table `utx_mvpaint_blend`, written into the staged bake's colour
        plane.  (`score` is the Euclidean distance here, the reference's scipy convention -- see INTEGRATION.md on torch_kdtree.)"""
        L = _lib.load()
        off = [C.c_size_t() for _ in range(4)]
        _lib.check(L.utx_uv_bake_layout(H2D, W2D, *[C.byref(o) for o in off]), "utx_uv_bake_layout")
        T = H2D * W2D
        pos = ws[off[1].value:off[1].value + T * 12].view(torch.float32).reshape(T, 3)
        col = ws[off[2].value:off[2].value + T * 12].view(torch.float32).reshape(T, 3)
        m = self.pbr_mesh
        covered = mask2d.bool()
        sel = rgba[..., 3] > 0.5
        cloud_p = interpolate(m.vertices, rast_mv, m.faces)[sel]
        cloud_c = rgba[..., :3][sel]
        cloud_n = m.normals[(rast_mv[..., 3].to(torch.int64) - 1)[sel]]
        tex_n = m.normals[(rast2d[0, ..., 3].to(torch.int64) - 1).reshape(-1)[covered]]
        score, index = knn(cloud_p, pos[covered], k=k, device=self.device)
        cloud_c, cloud_n, tex_n = cloud_c.contiguous().float(), cloud_n.contiguous().float(), tex_n.contiguous().float()
        out = torch.empty(index.shape[0], 3, device=self.device, dtype=torch.float32)
        _lib.check(L.utx_mvpaint_blend(_p(score), _p(index), index.shape[0], k, _p(cloud_c), _p(cloud_n), _p(tex_n), _p(out),
                                       _stream()), "utx_mvpaint_blend")
        col[covered] = out

    def _field_fill(self, ws, mask2d, rgba, rast_mv, H2D, W2D, union_cloud: bool):
        """The `*_inpainting=True` branches: the registered query field colours the texels the views do not own
        (:427-432, :609-614), or every covered texel from the union pixel cloud (`mean`, :387-389).  Reads and writes the
        staged bake's owner / position / colour planes in place (utx_uv_bake_layout)."""
        L = _lib.load()
        off = [C.c_size_t() for _ in range(4)]
        _lib.check(L.utx_uv_bake_layout(H2D, W2D, *[C.byref(o) for o in off]), "utx_uv_bake_layout")
        T = H2D * W2D
        owner = ws[off[0].value:off[0].value + T].view(torch.int8)
        pos = ws[off[1].value:off[1].value + T * 12].view(torch.float32).reshape(T, 3)
        col = ws[off[2].value:off[2].value + T * 12].view(torch.float32).reshape(T, 3)
        covered = mask2d.bool()
        if union_cloud:
            m = self.pbr_mesh
            sel = rgba[..., 3] > 0.5
            pix_pos = interpolate(m.vertices, rast_mv, m.faces)
            col[covered] = self.query_field(pix_pos[sel], rgba[..., :3][sel], pos[covered]).to(torch.float32)
            return
        seen = covered & (owner >= 0)
        unseen = covered & (owner < 0)
        if bool(unseen.any()) and bool(seen.any()):
            col[unseen] = self.query_field(pos[seen], col[seen], pos[unseen]).to(torch.float32)
