"""Oracle for the bake path (TEST INFRASTRUCTURE, see oracle/__init__.py).

Geometric kernels (rasterise / interpolate / LBVH / intersect) are the C restatement in oracle/bake_ref.c.  Everything
after them restates the reference's torch code with the same torch calls, on the CPU:
  uv_to_pcd                        TextureTools/texturetools/render/nvdiffrast/renderer_inverse.py:243-365
  get_boundary_mask                :435-444
  bake_mv_to_uv_reproject_blur     :574-633   (k=1 nearest neighbour restated as exact brute force, lowest index on ties;
                                               torch_kdtree@86961f7d [ext] leaves ties unspecified)
  bake_mv_to_uv_kdtree             :367-433   (`order_mean` and `mean`; k nearest neighbours = exact brute force, ascending by
                                               (distance, index)), query_field hook :93-103,:139-154
  lens_blur_torch                  texturetools/image/lens_blur.py:82-93,109-121,172-195,260-280
  pull_push                        texturetools/texture/stitching/mip.py:9-96
PARITY UNPINNED (no golden vectors in the reference); pinned here by analytic tests in tests/test_oracle_bake.py.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Dict, Optional

import numpy as np
import torch
import torch.nn.functional as F

from . import build_c

_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(str(build_c.build()))
    return _lib


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def rasterize(pos: np.ndarray, tri: np.ndarray, H: int, W: int) -> np.ndarray:
    pos = np.ascontiguousarray(pos, np.float32)
    tri = np.ascontiguousarray(tri, np.int32)
    B, V = pos.shape[0], pos.shape[1]
    out = np.zeros((B, H, W, 4), np.float32)
    lib().ora_rasterize(_ptr(pos), 1, V, _ptr(tri), tri.shape[0], B, H, W, _ptr(out))
    return out


def interpolate(attr: np.ndarray, rast: np.ndarray, tri: np.ndarray) -> np.ndarray:
    attr = np.ascontiguousarray(attr, np.float32)
    rast = np.ascontiguousarray(rast, np.float32)
    tri = np.ascontiguousarray(tri, np.int32)
    batched = attr.ndim == 3
    V, Cn = attr.shape[-2], attr.shape[-1]
    B, H, W, _ = rast.shape
    out = np.zeros((B, H, W, Cn), np.float32)
    lib().ora_interpolate(_ptr(attr), int(batched), V, Cn, _ptr(rast), _ptr(tri), B, H, W, _ptr(out))
    return out


def lbvh_build(vert: np.ndarray, tri: np.ndarray):
    vert = np.ascontiguousarray(vert, np.float32)
    tri = np.ascontiguousarray(tri, np.int32)
    Fn = tri.shape[0]
    info = np.zeros((2 * Fn - 1, 3), np.int32)
    aabb = np.zeros((2 * Fn - 1, 6), np.float32)
    srt = np.zeros((Fn, 2), np.int32)
    lib().ora_lbvh_build(_ptr(vert), vert.shape[0], _ptr(tri), Fn, _ptr(info), _ptr(aabb), _ptr(srt))
    return info, aabb, srt


def intersect(vert, tri, info, aabb, rays_o, rays_d):
    vert = np.ascontiguousarray(vert, np.float32)
    tri = np.ascontiguousarray(tri, np.int32)
    o = np.ascontiguousarray(rays_o, np.float32).reshape(-1, 3)
    d = np.ascontiguousarray(np.broadcast_to(rays_d, rays_o.shape), np.float32).reshape(-1, 3)
    N = o.shape[0]
    hit = np.zeros(N, np.uint8)
    tid = np.zeros(N, np.int32)
    pos = np.zeros((N, 3), np.float32)
    uv = np.zeros((N, 2), np.float32)
    lib().ora_intersect(_ptr(vert), _ptr(tri), _ptr(info), _ptr(aabb), _ptr(o), _ptr(d), C.c_int64(N), _ptr(hit), _ptr(tid),
                        _ptr(pos), _ptr(uv))
    return hit.astype(bool), tid, pos, uv


def intersect_leafscan(vert, tri, info, aabb, rays_o, rays_d):
    """`intersect` without the hierarchy: leaves scanned in descending Morton position, each leaf's own box tested against the
    running closest t (oracle/bake_ref.c: ora_intersect_leafscan).  Must equal `intersect` exactly."""
    vert = np.ascontiguousarray(vert, np.float32)
    tri = np.ascontiguousarray(tri, np.int32)
    o = np.ascontiguousarray(rays_o, np.float32).reshape(-1, 3)
    d = np.ascontiguousarray(np.broadcast_to(rays_d, rays_o.shape), np.float32).reshape(-1, 3)
    N = o.shape[0]
    hit = np.zeros(N, np.uint8)
    tid = np.zeros(N, np.int32)
    pos = np.zeros((N, 3), np.float32)
    uv = np.zeros((N, 2), np.float32)
    lib().ora_intersect_leafscan(_ptr(vert), _ptr(tri), _ptr(info), _ptr(aabb), tri.shape[0], _ptr(o), _ptr(d), C.c_int64(N), _ptr(hit),
                                 _ptr(tid), _ptr(pos), _ptr(uv))
    return hit.astype(bool), tid, pos, uv


# ------------------------------------------------------------------------------------------------ cameras
def intr_to_proj_ortho(intr: torch.Tensor, near=0.01, far=1000.0) -> torch.Tensor:
    """camera/conversion.py:19-27, perspective=False branch + the y-row negation."""
    p = torch.zeros(4, 4)
    p[0, 0], p[1, 1] = intr[0, 0], intr[1, 1]
    p[2, 2] = -2.0 / (far - near)
    p[3, 3] = 1.0
    p[0, 3] = -(2 * intr[0, 2] - 1)
    p[1, 3] = -(2 * intr[1, 2] - 1)
    p[2, 3] = -(far + near) / (far - near)
    p[1, :] = -p[1, :]
    return p


def intr_to_proj_persp(intr: torch.Tensor, near=0.01, far=1000.0) -> torch.Tensor:
    """camera/conversion.py:11-18, perspective=True branch + the y-row negation."""
    p = torch.zeros(4, 4)
    p[0, 0], p[1, 1] = 2 * intr[0, 0], 2 * intr[1, 1]
    p[2, 2] = -(far + near) / (far - near)
    p[0, 2], p[1, 2] = 2 * intr[0, 2] - 1, 2 * intr[1, 2] - 1
    p[3, 2] = -1.0
    p[2, 3] = -2.0 * far * near / (far - near)
    p[1, :] = -p[1, :]
    return p


def c2w_to_w2c(c2w: torch.Tensor) -> torch.Tensor:
    w2c = torch.zeros_like(c2w)
    w2c[..., :3, :3] = c2w[..., :3, :3].transpose(-1, -2)
    w2c[..., :3, 3:] = -c2w[..., :3, :3].transpose(-1, -2) @ c2w[..., :3, 3:]
    w2c[..., 3, 3] = 1.0
    return w2c


# ------------------------------------------------------------------------------------------------ torch tail
def lens_blur_torch(img: torch.Tensor, radius: float = 3.0, exposure_gamma: float = 5.0) -> torch.Tensor:
    """image/lens_blur.py:260-280 with components=5 (parameter row 4, scale 1.2)."""
    params = [[4.892608, 1.685979, -22.356787, 85.91246], [4.71187, 4.998496, 35.918936, -28.875618],
              [4.052795, 8.244168, -13.212253, -1.578428], [2.929212, 11.900859, 0.507991, 1.816328],
              [1.512961, 16.116382, 0.138051, -0.01]]
    scale = 1.2
    size = int(math.ceil(radius)) * 2 + 1
    ax = torch.linspace(-radius, radius, size, dtype=torch.float32) * scale * (1 / radius)
    comps = []
    for a, b, _, _ in params:
        k = torch.zeros(size, dtype=torch.complex64)
        k.real = torch.exp(-a * ax ** 2) * torch.cos(b * ax ** 2)
        k.imag = torch.exp(-a * ax ** 2) * torch.sin(b * ax ** 2)
        comps.append(k.reshape(1, size))
    total = 0.0                                     # :109-121 -- the reference's sequential fp32 accumulation order, kept
    for k, (_, _, A, B) in zip(comps, params):
        for i in range(size):
            for j in range(size):
                total += A * (k[0, i].real * k[0, j].real - k[0, i].imag * k[0, j].imag) + \
                         B * (k[0, i].real * k[0, j].imag + k[0, i].imag * k[0, j].real)
    total = total.sqrt()
    comps = [k / total for k in comps]
    img = torch.pow(img, exposure_gamma)
    Cn = img.shape[1]
    acc = 0.0
    for k, (_, _, A, B) in zip(comps, params):
        kr = k.real[None, None].repeat(Cn, 1, 1, 1)
        ki = k.imag[None, None].repeat(Cn, 1, 1, 1)
        pad = [0, size // 2]
        ir = F.conv2d(img, kr, padding=pad, groups=Cn)
        ii = F.conv2d(img, ki, padding=pad, groups=Cn)
        krt, kit, padt = kr.transpose(-1, -2), ki.transpose(-1, -2), pad[::-1]
        f1 = F.conv2d(ir, krt, padding=padt, groups=Cn)
        f2 = F.conv2d(ir, kit, padding=padt, groups=Cn)
        f3 = F.conv2d(ii, krt, padding=padt, groups=Cn)
        f4 = F.conv2d(ii, kit, padding=padt, groups=Cn)
        acc = acc + ((f1 - f4) * A + (f2 + f3) * B)
    out = torch.clamp(acc, 0, None)
    out = torch.pow(out, 1.0 / exposure_gamma)
    return torch.clamp(out, 0, 1)


def pull_push(map_Kd: torch.Tensor, map_mask: torch.Tensor):
    """texture/stitching/mip.py:51-96 ([N,C,H,W], [N,1,H,W] bool)."""
    B, Cn, H, W = map_Kd.shape
    n_level = max(min(int(math.log2(H)), int(math.log2(W))) - 2, 0)
    if n_level == 0:
        return map_Kd, map_mask
    map_Kd = torch.where(map_mask, map_Kd, torch.zeros((), dtype=map_Kd.dtype))

    def mip(Kd, mask):
        alpha = F.avg_pool2d(mask.float(), 2, 2, 0)
        Kdm = F.avg_pool2d(Kd, 2, 2, 0)
        bnd = (alpha > 0) & (alpha < 1)
        Kdm = torch.where(bnd, Kdm / torch.where(bnd, alpha, torch.ones_like(alpha)), Kdm)
        return Kdm, alpha > 0

    def fill(Kd, mask, Kdm):
        k = torch.tensor([[[0.5625, 0.1875], [0.1875, 0.0625]], [[0.1875, 0.5625], [0.0625, 0.1875]],
                          [[0.1875, 0.0625], [0.5625, 0.1875]], [[0.0625, 0.1875], [0.1875, 0.5625]]], dtype=Kd.dtype)[:, None]
        Hm, Wm = Kdm.shape[-2:]
        pad = F.pad(Kdm, [1, 1, 1, 1], mode="replicate")
        conv = F.conv2d(pad, k.repeat(Cn, 1, 1, 1), None, 1, 0, 1, Cn)
        conv = conv.reshape(B, Cn, 2, 2, Hm + 1, Wm + 1).permute(0, 1, 4, 2, 5, 3).reshape(B, Cn, (Hm + 1) * 2, (Wm + 1) * 2)
        return torch.where(mask, Kd, conv[:, :, 1:-1, 1:-1])

    Ks, Ms = [], []
    K, M = map_Kd, map_mask
    for _ in range(n_level):
        K, M = mip(K, M)
        Ks.append(K)
        Ms.append(M)
    K = Ks[-1]
    for lvl in range(n_level - 1, 0, -1):
        K = fill(Ks[lvl - 1], Ms[lvl - 1], K)
    return fill(map_Kd, map_mask, K), map_mask


def boundary_mask(mask: torch.Tensor, k: int = 3) -> torch.Tensor:
    """:435-444 on [N,H,W,1] bool."""
    a = mask.float().permute(0, 3, 1, 2)
    inner = (a - (1.0 - F.max_pool2d(1.0 - a, 2 * (k // 2) + 1, 1, k // 2))) > 0
    outer = (F.max_pool2d(a, 2 * (k // 2) + 1, 1, k // 2) - a) > 0
    return (inner | outer).permute(0, 2, 3, 1)


def nearest_index(src: torch.Tensor, dst: torch.Tensor, chunk: int = 2048) -> torch.Tensor:
    """Exact 1-NN of each dst point among src ([N,3], [M,3] fp32): fp32 squared distance ((dx^2+dy^2)+dz^2), lowest index
    on ties."""
    out = torch.empty(dst.shape[0], dtype=torch.int64)
    for i in range(0, dst.shape[0], chunk):
        d = src[None, :, :] - dst[i:i + chunk, None, :]
        d2 = (d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1]) + d[..., 2] * d[..., 2]
        out[i:i + chunk] = torch.argmin(d2, dim=1)      # first minimum = lowest index
    return out


def nearest_k(src: torch.Tensor, dst: torch.Tensor, k: int, chunk: int = 512):
    """Exact k nearest src points of each dst point, ascending by (fp32 squared distance ((dx^2+dy^2)+dz^2), index):
    -> (distance [M,k] fp32, index [M,k] int64), the `knn(src, dst, k)` contract of pcd/knn/__init__.py:104-114."""
    M = dst.shape[0]
    idx = torch.empty(M, k, dtype=torch.int64)
    dist = torch.empty(M, k, dtype=torch.float32)
    for i in range(0, M, chunk):
        d = src[None, :, :] - dst[i:i + chunk, None, :]
        d2 = (d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1]) + d[..., 2] * d[..., 2]
        v, j = torch.sort(d2, dim=1, stable=True)            # stable: equal distances keep ascending index order
        idx[i:i + chunk] = j[:, :k]
        dist[i:i + chunk] = torch.from_numpy(np.sqrt(v[:, :k].contiguous().numpy()))   # IEEE sqrt; torch's vectorised CPU sqrt is not correctly rounded
    return dist, idx


@torch.no_grad()
def infer_reproject(*a, **kw):
    return infer(*a, **kw)


@torch.no_grad()
def infer(vert, tri, uv, tri_uv, c2ws: torch.Tensor, intrinsics: torch.Tensor, image_attrs: torch.Tensor,
                    H: int, W: int, H2: int, W2: int, angle_deg: float = 100.0, index=(0, 3, 4, 1, 2, 5), method: str = "reproject",
                    kdtree_method: str = "order_mean", k_all: int = 32, k_vis: int = 1, k_invis: int = 32,
                    query_field=None, filt_gradient_points: bool = False, grad_norm_threhold: float = 0.20,
                    perspective: bool = False, reproject_method: str = "lens", kernel_size_blur: int = 5,
                    nn_index_given: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
    """NVDiffRendererInverse.infer(perspective=False, filt_gradient_points=False) (:635-726) for method='reproject'
    (reproject_method='lens') and method='kdtree' (kdtree_method 'order_mean' | 'mean'); `query_field` set = the
    *_inpainting=True branches (:387-389, :427-432, :609-614).
    `nn_index_given` [H2*W2] int64 (flat texel index of the source, -1 elsewhere): at atlas sizes where the brute-force 1-NN of
    the fill is out of reach (2048^2: 0.7 M x 2.1 M pairs) the caller supplies the neighbour table -- after checking a sample of
    it against `nearest_index` -- and everything else (visibility, composite, seams, blur, pull-push) is still computed here."""
    vert = np.ascontiguousarray(vert, np.float32)
    tri = np.ascontiguousarray(tri, np.int32)
    tri_uv = np.ascontiguousarray(tri_uv, np.int32)
    n = c2ws.shape[0]
    V = torch.from_numpy(vert)
    mats = torch.matmul(intr_to_proj_persp(intrinsics) if perspective else intr_to_proj_ortho(intrinsics), c2w_to_w2c(c2ws))   # [n,4,4]
    vh = torch.cat([V, torch.ones_like(V[:, :1])], -1)
    clip = torch.matmul(vh, mats.permute(0, 2, 1))                                           # :263  [n,V,4]
    # mv_to_pcd :183-214
    rast_mv = torch.from_numpy(rasterize(clip.numpy(), tri, H, W))
    alpha_vis = (rast_mv[..., 3:4] > 0).float()
    if filt_gradient_points:                                                                 # mv_to_pcd :188-214
        mask_mv = rast_mv[..., 3:4] > 0
        Fl = torch.from_numpy(tri.astype(np.int64))
        ar = torch.linalg.cross(V[Fl[:, 1]] - V[Fl[:, 0]], V[Fl[:, 2]] - V[Fl[:, 0]], dim=-1)
        vn = torch.zeros(V.shape[0], 3, 3)
        vn.scatter_add_(0, Fl.unsqueeze(-1).expand(-1, -1, 3), ar.unsqueeze(1).expand(-1, 3, -1))     # structure_v2.py:63-71
        vn = F.normalize(vn.mean(dim=1), dim=-1)
        attrs = torch.from_numpy(interpolate(torch.cat([V, vn], -1).numpy(), rast_mv.numpy(), tri))
        a_dy, a_dx = torch.gradient(attrs, dim=(1, 2))
        gnorm = (a_dx.square() + a_dy.square()).sum(dim=-1, keepdim=True).sqrt()
        tid_mv = rast_mv[..., 3:4].to(torch.int64).sub(1)
        fn_mv = F.normalize(ar, dim=-1).gather(0, torch.where(mask_mv, tid_mv, 0).reshape(-1, 1).repeat(1, 3)).reshape(n, H, W, 3)
        rd = attrs[..., 0:3] - c2ws[:, :3, 3][:, None, None, :] if perspective else (-c2ws[:, :3, 2])[:, None, None, :]    # :191-196
        rd = F.normalize(rd, dim=-1)
        cos_mv = F.cosine_similarity(torch.broadcast_tensors(rd, fn_mv)[0], fn_mv, dim=-1).unsqueeze(-1)
        ok_grad = gnorm < grad_norm_threhold
        # nn.MaxPool2d(31, 1, 15) applied to the [n,H,W,1] tensor as it stands (:204-205): torch reads it as [N, C=H, H=W, W=1],
        # so the erosion runs along the image x axis only -- kept
        eroded = (1.0 - F.max_pool2d(1.0 - ok_grad.float(), kernel_size=31, stride=1, padding=15)).bool()
        alpha_vis = (mask_mv & (cos_mv < math.cos(math.radians(angle_deg))) & eroded).float()
    # uv_to_pcd
    uvc = np.concatenate([uv, np.zeros_like(uv[:, :1]), np.ones_like(uv[:, :1])], -1)[None].astype(np.float32)
    rast2 = rasterize(uvc, tri_uv, H2, W2)
    mask_2d = torch.from_numpy(rast2[..., 3:4] > 0)
    tid_2d = torch.from_numpy(rast2[..., 3]).long() - 1                                      # [1,H2,W2]
    pos_2d = torch.from_numpy(interpolate(vert, rast2, tri))                                 # [1,H2,W2,3]
    Ft = torch.from_numpy(tri.astype(np.int64))
    areas = torch.linalg.cross(V[Ft[:, 1]] - V[Ft[:, 0]], V[Ft[:, 2]] - V[Ft[:, 0]], dim=-1)
    normals = F.normalize(areas, dim=-1)
    fn_2d = normals[torch.where(mask_2d[..., 0], tid_2d, torch.zeros_like(tid_2d))]          # [1,H2,W2,3]
    if perspective:                                                                          # :279-284
        rays_o = c2ws[:, :3, 3][:, None, None, :]
        rays_d = pos_2d - rays_o
    else:
        rays_d = (-c2ws[:, :3, 2])[:, None, None, :]
        rays_o = pos_2d - (2.0 * math.sqrt(3.0)) * rays_d
    rays_d = F.normalize(rays_d, dim=-1)
    rays_o, rays_d = torch.broadcast_tensors(rays_o, rays_d)
    ndc_v = clip[..., :2] / clip[..., 3:4]
    ndc_2d = torch.from_numpy(interpolate(ndc_v.numpy(), np.repeat(rast2, n, 0), tri))       # :288
    img = torch.cat([image_attrs, alpha_vis], -1)
    samp = F.grid_sample(img.permute(0, 3, 1, 2), ndc_2d, mode="bilinear", align_corners=False).permute(0, 2, 3, 1)
    Cn = image_attrs.shape[-1]                                                               # 3 (RGB) or 9 (PBR, :711-719)
    col_2d, alpha_2d = samp[..., :Cn], samp[..., Cn:Cn + 1]
    m = mask_2d[0, ..., 0]
    info, aabb, _ = lbvh_build(vert, tri)
    ro, rd = rays_o[:, m], rays_d[:, m]                                                      # [n,Nv,3]
    _, rt, _, _ = intersect(vert, tri, info, aabb, ro.numpy(), rd.numpy())
    rt = torch.from_numpy(rt.astype(np.int64)).reshape(n, -1)
    tv = tid_2d[0][m]
    ok = (rt == tv[None]) & (rt != -1)
    cosv = F.cosine_similarity(rd, fn_2d[0][m][None].expand(n, -1, -1), dim=-1)
    ok = ok & (cosv < math.cos(math.radians(angle_deg)))
    vis = torch.zeros(n, H2, W2, 1, dtype=torch.bool)
    vis[:, m, 0] = ok
    raw_vis = vis.clone()
    for k in (3, 5):                                                                          # :329-339 with kernel_mode=7
        ker = F.pad(torch.full((1, 1, k - 2, k - 2), -1.0), (1, 1, 1, 1), value=float(k * k))
        conv = F.conv2d(vis.float().permute(0, 3, 1, 2), ker, stride=1, padding=k // 2).permute(0, 2, 3, 1)
        vis = vis | (conv >= ((k - 1) ** 2 - 1) * ((k - 2) ** 2))
    vis = vis & mask_2d & (alpha_2d > 0.999)
    if method == "kdtree":
        assert kdtree_method in ("mean", "mvpaint", "order_mean")
        # mv_to_pcd point clouds (:188, :227-231): positions interpolated at the pixels, one cloud per view
        attrs_mv = torch.from_numpy(interpolate(vert, rast_mv.numpy(), tri))                 # [n,H,W,3]
        mmv = alpha_vis[..., 0] > 0                                                          # mask_visiable (:233-239): the raster mask unless filtered
        clouds = [(attrs_mv[i][mmv[i]], image_attrs[i][mmv[i]]) for i in range(n)]
        P2 = pos_2d[0][m]                                                                    # point_cloud_2d.vertices
        wc = torch.zeros(P2.shape[0], Cn)
        cur = torch.zeros(1, H2, W2, 1, dtype=torch.bool)
        if kdtree_method == "mean":                                                          # :385-389
            allp, allc = torch.cat([c[0] for c in clouds]), torch.cat([c[1] for c in clouds])
            wc = query_field(allp, allc, P2) if query_field is not None else allc[nearest_k(allp, P2, k_all)[1]].mean(dim=-2)
        elif kdtree_method == "mvpaint":                                                     # :390-399 (MVPaint, arXiv 2411.02336 sec. 3.2)
            allp, allc = torch.cat([c[0] for c in clouds]), torch.cat([c[1] for c in clouds])
            tid_mv = rast_mv[..., 3].to(torch.int64) - 1
            alln = torch.cat([normals[tid_mv[i][mmv[i]]] for i in range(n)])                  # point_cloud_visiable.normals: face normals (:236)
            score, idx = nearest_k(allp, P2, k_all)
            weight = F.normalize(score.reciprocal().nan_to_num(nan=0.0), p=1, dim=-1) * \
                F.cosine_similarity(alln[idx], fn_2d[0][m].unsqueeze(-2), dim=-1)
            weight = weight.unsqueeze(-1)
            wc = (allc[idx] * weight).sum(dim=-2) / weight.sum(dim=-2)
            wc = torch.nan_to_num(wc, nan=0.0, posinf=0.0, neginf=0.0)
        else:                                                                                # order_mean :406-432
            for i in index:
                extra = (~cur) & vis[i:i + 1]
                sel = extra[0, ..., 0][m]
                if sel.any():
                    _, idx = nearest_k(clouds[i][0], P2[sel], k_vis)
                    wc[sel] = clouds[i][1][idx].mean(dim=-2)
                cur = cur | extra
            vsel = cur[0, ..., 0][m]
            if (~vsel).any():
                if query_field is not None:
                    wc[~vsel] = query_field(P2[vsel], wc[vsel], P2[~vsel])
                else:
                    kk = min(k_invis, int(vsel.sum()))
                    _, idx = nearest_k(P2[vsel], P2[~vsel], kk)
                    wc[~vsel] = wc[vsel][idx].mean(dim=-2)
        color = torch.zeros(1, H2, W2, Cn)
        color[0][m] = wc
        color_2d = pull_push(color.permute(0, 3, 1, 2), mask_2d.permute(0, 3, 1, 2))[0].permute(0, 2, 3, 1)
        return {"mask_2d": mask_2d, "mask_2d_visiable": vis, "color_2d": color_2d, "pre_pull_push": color,
                "rast_2d": torch.from_numpy(rast2)}
    # bake_mv_to_uv_reproject_blur
    color = torch.zeros(1, H2, W2, Cn)
    cur = torch.zeros(1, H2, W2, 1, dtype=torch.bool)
    bnd = torch.zeros(1, H2, W2, 1, dtype=torch.bool)
    owner = torch.full((H2, W2), -1, dtype=torch.int64)
    for i in index:
        extra = (~cur) & vis[i:i + 1]
        color = torch.where(extra, col_2d[i:i + 1], color)
        owner[extra[0, ..., 0]] = i
        cur = cur | extra
        bnd = bnd | boundary_mask(extra, 3)
    bnd = F.max_pool2d(bnd.float().permute(0, 3, 1, 2), 3, 1, 1).permute(0, 2, 3, 1) > 0
    bnd = ((1.0 - F.max_pool2d(1.0 - mask_2d.float().permute(0, 3, 1, 2), 7, 1, 3).permute(0, 2, 3, 1)) > 0) & bnd
    vis_m = cur[0, ..., 0] & m
    inv_m = (~cur[0, ..., 0]) & m
    nn_index = torch.full((H2 * W2,), -1, dtype=torch.int64)
    if inv_m.any() and vis_m.any():
        if query_field is not None:                                                          # inpainting=True (:612-613)
            color[0][inv_m] = query_field(pos_2d[0][vis_m], color[0][vis_m], pos_2d[0][inv_m])
        elif nn_index_given is not None:
            nn_index = nn_index_given.reshape(-1).to(torch.int64).clone()
            color[0][inv_m] = color[0].reshape(-1, Cn)[nn_index[inv_m.reshape(-1)]]
        else:
            src_idx = torch.nonzero(vis_m.reshape(-1))[:, 0]
            idx = nearest_index(pos_2d[0][vis_m], pos_2d[0][inv_m])
            color[0][inv_m] = color[0][vis_m][idx]
            nn_index[inv_m.reshape(-1)] = src_idx[idx]
    pre_blur = color.clone()
    if reproject_method == "gaussian":                                                       # :619-620 -> torchvision gaussian_blur [ext, present]
        from torchvision.transforms.functional import gaussian_blur
        blur = gaussian_blur(color.permute(0, 3, 1, 2), kernel_size=(kernel_size_blur, kernel_size_blur), sigma=None).permute(0, 2, 3, 1)
    else:
        blur = lens_blur_torch(color.permute(0, 3, 1, 2)).permute(0, 2, 3, 1)
    color = torch.where(bnd, blur, color)
    color_2d = pull_push(color.permute(0, 3, 1, 2), mask_2d.permute(0, 3, 1, 2))[0].permute(0, 2, 3, 1)
    return {"mask_2d": mask_2d, "mask_2d_visiable": vis, "raw_visible": raw_vis, "color_2d": color_2d, "tid_2d": tid_2d,
            "owner": owner, "seam": bnd, "nn_index": nn_index, "pre_blur": pre_blur, "rast_2d": torch.from_numpy(rast2),
            "alpha_mv": alpha_vis, "rays_tid": rt, "pos_2d": pos_2d, "rast_mv": rast_mv}


# ------------------------------------------------------------------------------------------------ forward G-buffers (b1)
@torch.no_grad()
def export_condition(vert, tri, vertex_normals, geometry_scale=1.0, n_views=6, n_rows=2, n_cols=3, H=512, W=512, scale=1.0,
                     background=128.0 / 255.0, perspective=False, fov_deg=49.1, orbit=False):
    """VideoExporter.export_condition (video/export_nvdiffrast_video.py:900-999, orthographic box views) over
    NVDiffRendererBase.simple_rendering(render_world_normal, render_world_position, enable_antialis=False)
    (render/nvdiffrast/renderer_base.py:101-200) and Mesh.scale_to_bbox / apply_transform (mesh/structure.py:190-202,
    :290-304).  -> uint8 grids alpha [n_rows H, n_cols W], ccm / normal [.., 3], c2ws, intrinsics.  Pinned against the
    reference's own run by tests/golden/ref_glue.npz."""
    from unitex_b200.bake import generate_box_views_c2ws, generate_intrinsics, generate_orbit_views_c2ws   # pinned themselves against camera/generator.py
    v = torch.from_numpy(np.ascontiguousarray(vert, np.float32))
    bbox = torch.stack([v.min(0).values, v.max(0).values])
    ccc = bbox.mean(dim=0)
    sss = ((bbox[1] - bbox[0]) / (2.0 * geometry_scale)).max()
    T = torch.eye(4)
    T[[0, 1, 2], [0, 1, 2]] = 1 / sss
    T[:3, 3] = -ccc / sss
    v = torch.matmul(torch.cat([v, torch.ones_like(v[:, :1])], -1), T.T)[:, :3].contiguous()
    vn = F.normalize(torch.matmul(torch.from_numpy(np.ascontiguousarray(vertex_normals, np.float32)), T[:3, :3].T), dim=-1).contiguous()
    sel = {1: [0], 2: [0, 2], 4: [0, 1, 2, 3], 6: [0, 1, 4, 2, 3, 5] if (n_rows, n_cols) == (2, 3) else list(range(6))}.get(n_views)
    c2ws = generate_orbit_views_c2ws(n_views + 1, radius=2.8, height=0.0, theta_0=0.0, degree=True)[:n_views] if orbit else generate_box_views_c2ws(radius=2.8)[sel]
    intr = generate_intrinsics(fov_deg, fov_deg, fov=True, degree=True) if perspective else generate_intrinsics(scale, scale, fov=False, degree=False)
    mvp = torch.matmul(intr_to_proj_persp(intr) if perspective else intr_to_proj_ortho(intr), c2w_to_w2c(c2ws))
    clip = torch.matmul(torch.cat([v, torch.ones_like(v[:, :1])], -1), mvp.permute(0, 2, 1))
    tri = np.ascontiguousarray(tri, np.int32)
    rast = rasterize(clip.numpy(), tri, H, W)
    mask = torch.from_numpy(rast[..., 3:4] > 0)
    alpha = mask.float()
    nrm = F.normalize(torch.from_numpy(interpolate(vn.numpy(), rast, tri)), dim=-1)
    nrm = torch.lerp(torch.full_like(nrm, -1.0), nrm, alpha)
    pos = torch.from_numpy(interpolate(v.numpy(), rast, tri))
    pos = torch.lerp(torch.full_like(pos, -1.0), pos, mask.float())
    ccm, normal = pos.mul(0.5).add(0.5), nrm.mul(0.5).add(0.5)
    if background is not None:
        bg = torch.full((3,), float(background)).float()   # 'grey' = PIL #808080 = 128/255 (utils/parse_color.py:6)
        ccm = ccm * alpha + bg * (1.0 - alpha)
        normal = normal * alpha + bg * (1.0 - alpha)

    def grid(t, c):
        a = t.clamp(0.0, 1.0).mul(255.0).numpy().astype(np.uint8)
        return a.reshape(n_rows, n_cols, H, W, c).transpose(0, 2, 1, 3, 4).reshape(n_rows * H, n_cols * W, c)
    return {"alpha": grid(alpha, 1)[..., 0], "ccm": grid(ccm, 3), "normal": grid(normal, 3), "c2ws": c2ws, "intrinsics": intr}
