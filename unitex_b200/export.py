"""Forward G-buffer render (`VideoExporter.export_condition`) and the mesh/texture wire formats around the bake
(SURVEY 8f rows 2-3): reference TextureTools/texturetools/video/export_nvdiffrast_video.py:900-999,
render/nvdiffrast/renderer_base.py:101-200, mesh/structure.py:290-304 (scale_to_bbox), io/link_pbr_to_mesh.py:9-31.

Rasterise + interpolate run in libunitex_b200.so; the per-pixel post-processing of six 512^2 images (normalise, lerp to
background, uint8) is left to torch as plumbing.  Vertex normals are the area-weighted ones of PBRMesh
(mesh/structure_v2.py:63-71); the reference's exporter takes trimesh's [ext] -- a documented deviation.
"""
from __future__ import annotations

import json
import struct
from typing import Dict, Optional

import numpy as np
import torch

from . import bake as ub


def parse_color(c):
    """utils/parse_color.py:6-18: names resolve through PIL's colour table to 8-bit RGB / 255 ('grey' is #808080 = 128/255, not
    0.5 -- pinned by tests/golden/ref_glue.npz); a float broadcasts; three floats are taken as given."""
    if c is None:
        return None
    if isinstance(c, str):
        from PIL import ImageColor
        if c not in ImageColor.colormap:
            raise NotImplementedError(f"unknown colour name {c!r}")
        return torch.tensor(ImageColor.getrgb(c)[:3], dtype=torch.float32).div(255.0)   # (PIL caches tuples in `colormap`)
    if isinstance(c, float):
        return torch.tensor([c], dtype=torch.float32)
    if isinstance(c, (tuple, list)) and len(c) == 3 and all(isinstance(x, float) for x in c):
        return torch.tensor(c, dtype=torch.float32)
    raise NotImplementedError(f"colour {c!r}")


def vertex_normals(vertices: torch.Tensor, faces: torch.Tensor) -> torch.Tensor:
    """mesh/structure_v2.py:49-50,63-71: area vectors scattered to the three corners, normalised."""
    f = faces.long()
    areas = torch.linalg.cross(vertices[f[:, 1]] - vertices[f[:, 0]], vertices[f[:, 2]] - vertices[f[:, 0]], dim=-1)
    vn = torch.zeros_like(vertices)
    for k in range(3):
        vn.index_add_(0, f[:, k], areas)
    return torch.nn.functional.normalize(vn, dim=-1)


def scale_to_bbox(vertices: torch.Tensor, scale: float = 1.0) -> torch.Tensor:
    """mesh/structure.py:290-304 + apply_transform: centre the bbox, largest side -> 2*scale."""
    lo, hi = vertices.min(0).values, vertices.max(0).values
    ccc = (lo + hi) / 2
    sss = ((hi - lo) / (2.0 * scale)).max()
    return (vertices - ccc) / sss


class VideoExporter:
    """export_condition (box or orbit views, orthographic or perspective); the orbit VIDEO / CAD exporters are out of scope."""

    def __init__(self, device="cuda"):
        self.device = torch.device(device)

    @torch.no_grad()
    def export_condition(self, mesh_path, geometry_scale=1.0, n_views=4, n_rows=2, n_cols=2, H=512, W=512, scale=0.85,
                         fov_deg=49.1, perspective=False, orbit=False, background=None, return_info=False,
                         return_image=True, return_mesh=False, return_camera=False) -> Dict:
        from PIL import Image
        assert n_views == n_rows * n_cols, f"Value Error: (n_views, n_rows, n_cols)={(n_views, n_rows, n_cols)}"
        if isinstance(mesh_path, str):
            V, F, _, _ = ub.load_mesh(mesh_path)
        else:
            V, F = mesh_path
        v = scale_to_bbox(torch.as_tensor(V, dtype=torch.float32, device=self.device), geometry_scale)
        f = torch.as_tensor(F, device=self.device).to(torch.int32).contiguous()
        # once per mesh, on the host: the device's index_add_ accumulates with atomics in an order that changes run to run, and
        # a last-bit change of a normal can move a uint8 of the normal map -- the control image of the FLUX call -- so the same
        # asset would not reproduce across processes / ranks (tests/test_gpu_batch.py holds 2 GPUs to the 1-GPU bytes)
        vn = vertex_normals(v.cpu(), f.cpu()).to(self.device)
        if orbit:                                                              # :922-923
            c2ws = ub.generate_orbit_views_c2ws(n_views + 1, radius=2.8, height=0.0, theta_0=0.0, degree=True)[:n_views]
        else:
            assert n_views in (1, 2, 4, 6), f"Value Error: n_views={n_views}"
            sel = {1: [0], 2: [0, 2], 4: [0, 1, 2, 3], 6: [0, 1, 4, 2, 3, 5] if (n_rows, n_cols) == (2, 3) else [0, 1, 2, 3, 4, 5]}[n_views]
            c2ws = ub.generate_box_views_c2ws(radius=2.8)[sel]
        intr = ub.generate_intrinsics(fov_deg, fov_deg, fov=True, degree=True) if perspective else ub.generate_intrinsics(scale, scale, fov=False)
        mats = torch.matmul(ub.intr_to_proj(intr, perspective=bool(perspective)), ub.c2w_to_w2c(c2ws)).to(self.device)
        rast = ub.rasterize(ub.transform_points(v.contiguous(), mats), f, (H, W))
        alpha = (rast[..., 3:4] > 0).float()
        attrs = ub.interpolate(torch.cat([v, vn], -1).contiguous(), rast, f)
        pos = torch.lerp(torch.full_like(attrs[..., :3], -1.0), attrs[..., :3], alpha)
        nrm = torch.lerp(torch.full_like(attrs[..., 3:], -1.0), torch.nn.functional.normalize(attrs[..., 3:], dim=-1), alpha)
        bg = parse_color(background)
        ccm, normal = pos * 0.5 + 0.5, nrm * 0.5 + 0.5
        if bg is not None:
            bg = bg.to(self.device)
            ccm, normal = ccm * alpha + bg * (1 - alpha), normal * alpha + bg * (1 - alpha)
        cam = {"c2ws": c2ws.to(self.device), "intrinsics": intr.to(self.device), "perspective": perspective}
        if not return_image:
            out = {"alpha": alpha.cpu().numpy(), "ccm": ccm.cpu().numpy(), "normal": normal.cpu().numpy()}
        else:
            def grid(t, c):
                a = t.clamp(0, 1).mul(255.0).cpu().numpy().astype(np.uint8)
                a = a.reshape(n_rows, n_cols, H, W, c).transpose(0, 2, 1, 3, 4).reshape(n_rows * H, n_cols * W, c)
                return Image.fromarray(a[..., 0], mode="L") if c == 1 else Image.fromarray(a, mode="RGB")
            out = {"alpha": grid(alpha, 1), "ccm": grid(ccm, 3), "normal": grid(normal, 3)}
        if return_camera:
            out.update(cam)
        return out


# ------------------------------------------------------------------------------------------------ grid re-ordering (a9)
def control_grid_to_strip(normal_grid: np.ndarray, ccm_grid: np.ndarray) -> np.ndarray:
    """pipeline.py:239-244: 0.5*normal + 0.5*ccm in uint8, tile 5 rotated 180 deg, 2x3 grid -> 1x6 strip in order [0,4,1,3,2,5]."""
    a = np.asarray(normal_grid).reshape(2, 512, 3, 512, -1)
    b = np.asarray(ccm_grid).reshape(2, 512, 3, 512, -1)
    t = (0.5 * a + 0.5 * b).astype(np.uint8)
    t[1, :, 2] = t[1, ::-1, 2, ::-1]
    return t.transpose(0, 2, 1, 3, 4).reshape(6, 512, 512, -1)[[0, 4, 1, 3, 2, 5]].transpose(1, 0, 2, 3).reshape(512, 6 * 512, -1)


def strip_to_view_grid(strip: np.ndarray) -> np.ndarray:
    """pipeline.py:280-282: un-rotate tile 5, order [0,2,4,3,1,5] back to the 2x3 f,r,t,b,l,d grid."""
    t = np.array(strip).reshape(512, 6, 512, -1)
    t[:, 5] = t[::-1, 5, ::-1]
    return t.transpose(1, 0, 2, 3)[[0, 2, 4, 3, 1, 5]].reshape(2, 3, 512, 512, -1).transpose(0, 2, 1, 3, 4).reshape(2 * 512, 3 * 512, -1)


# ------------------------------------------------------------------------------------------------ OBJ / GLB
def save_obj(path: str, V, F, UV=None, F_uv=None):
    """Wavefront OBJ (v / vt / f a/b corners, 1-based), written in bulk: the 500 k-face reference mesh takes ~1 s."""
    V = np.asarray(V, np.float64).reshape(-1, 3)
    F = np.asarray(F, np.int64).reshape(-1, 3) + 1
    with open(path, "w") as fh:
        fh.write("\n".join("v %.8f %.8f %.8f" % (a, b, c) for a, b, c in V.tolist()))
        fh.write("\n")
        if UV is not None:
            UV = np.asarray(UV, np.float64).reshape(-1, 2)
            Fu = np.asarray(F_uv, np.int64).reshape(-1, 3) + 1
            fh.write("\n".join("vt %.8f %.8f" % (a, b) for a, b in UV.tolist()))
            fh.write("\n")
            fh.write("\n".join("f %d/%d %d/%d %d/%d" % (r[0], r[3], r[1], r[4], r[2], r[5]) for r in np.concatenate([F, Fu], 1).tolist()))
        else:
            fh.write("\n".join("f %d %d %d" % (a, b, c) for a, b, c in F.tolist()))
        fh.write("\n")


def save_glb(path: str, V, F, UV, F_uv, texture_rgb: np.ndarray):
    """io/link_pbr_to_mesh.py:9-31 (link_rgb_to_mesh): baseColor texture = atlas flipped vertically, PBR material, one
    primitive.  Corners are un-welded to unique (vertex, uv) pairs as glTF needs one index buffer."""
    import io
    from PIL import Image
    key = np.stack([np.asarray(F).reshape(-1), np.asarray(F_uv).reshape(-1)], -1)
    uniq, inv = np.unique(key, axis=0, return_inverse=True)
    pos = np.asarray(V, np.float32)[uniq[:, 0]]
    uv = np.asarray(UV, np.float32)[uniq[:, 1]].copy()
    uv[:, 1] = 1.0 - uv[:, 1]                                   # glTF's v axis points down
    idx = inv.reshape(-1).astype(np.uint32)
    buf = io.BytesIO()
    Image.fromarray(np.asarray(texture_rgb)[::-1].copy()).save(buf, format="PNG")
    png = buf.getvalue()
    chunks, views, off = [], [], 0
    for data, target in ((pos.tobytes(), 34962), (uv.tobytes(), 34962), (idx.tobytes(), 34963), (png, None)):
        pad = (-len(data)) % 4
        v = {"buffer": 0, "byteOffset": off, "byteLength": len(data)}
        if target:
            v["target"] = target
        views.append(v)
        chunks.append(data + b"\x00" * pad)
        off += len(data) + pad
    gltf = {
        "asset": {"version": "2.0", "generator": "unitex-b200"}, "scene": 0, "scenes": [{"nodes": [0]}], "nodes": [{"mesh": 0}],
        "meshes": [{"primitives": [{"attributes": {"POSITION": 0, "TEXCOORD_0": 1}, "indices": 2, "material": 0}]}],
        "materials": [{"pbrMetallicRoughness": {"baseColorTexture": {"index": 0}, "metallicFactor": 0.0, "roughnessFactor": 1.0}}],
        "textures": [{"source": 0}], "images": [{"bufferView": 3, "mimeType": "image/png"}],
        "accessors": [
            {"bufferView": 0, "componentType": 5126, "count": len(pos), "type": "VEC3", "min": pos.min(0).tolist(), "max": pos.max(0).tolist()},
            {"bufferView": 1, "componentType": 5126, "count": len(uv), "type": "VEC2"},
            {"bufferView": 2, "componentType": 5125, "count": len(idx), "type": "SCALAR"}],
        "bufferViews": views, "buffers": [{"byteLength": off}],
    }
    js = json.dumps(gltf, separators=(",", ":")).encode()
    js += b" " * ((-len(js)) % 4)
    binc = b"".join(chunks)
    with open(path, "wb") as fh:
        fh.write(struct.pack("<III", 0x46546C67, 2, 12 + 8 + len(js) + 8 + len(binc)))
        fh.write(struct.pack("<II", len(js), 0x4E4F534A) + js)
        fh.write(struct.pack("<II", len(binc), 0x004E4942) + binc)
