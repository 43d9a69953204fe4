"""GPU, BASELINE.json config 4 geometry: the reference's own test mesh (test_cases/teaser_robot: V 269 026, F 499 981,
31-level LBVH, duplicate Morton codes, large-triangle raster path) from the committed lossless fixture
tests/golden/teaser_robot.npz.xz.  Everything integer is held BIT-EXACT against the C oracle (oracle/bake_ref.c) at full
scale: the 6 x 512^2 view rasters, the 2048^2 UV raster, the LBVH node arrays, closest-hit ids / positions / barycentrics of
1 M rays, the six 2048^2 visibility masks; the 1-NN fill is compared with brute force on a sample of the 0.7 M queries, and the
final colours with the oracle's tail evaluated on that (sample-verified) neighbour table.
Reference: TextureTools/texturetools/render/nvdiffrast/renderer_inverse.py:243-365 (uv_to_pcd), :574-633 (bake)."""
import numpy as np
import pytest
import torch

from tests.bake_meshes import teaser_robot

pytestmark = pytest.mark.gpu

COLOR_ATOL = 2e-4


def _views():
    from unitex_b200.bake import generate_box_views_c2ws, generate_intrinsics
    return generate_box_views_c2ws(2.8)[[0, 1, 4, 2, 3, 5]], generate_intrinsics(1.0, 1.0, fov=False)


@pytest.fixture(scope="module")
def teaser():
    v, f, uv, fuv = teaser_robot()
    assert v.shape == (269026, 3) and f.shape == (499981, 3)
    return v, f, uv, fuv


def test_teaser_rasters_bit_exact(lib, teaser):
    from oracle import bake as ob
    from unitex_b200 import bake as ub
    v, f, uv, fuv = teaser
    c2ws, intr = _views()
    mats = torch.matmul(ub.intr_to_proj(intr, perspective=False), ub.c2w_to_w2c(c2ws))
    clip = ub.transform_points(torch.from_numpy(v).cuda(), mats.cuda())
    rast = ub.rasterize(clip, torch.from_numpy(f).cuda(), (512, 512))
    attr = ub.interpolate(torch.from_numpy(v).cuda(), rast, torch.from_numpy(f).cuda())
    torch.cuda.synchronize()
    ref = ob.rasterize(clip.cpu().numpy(), f, 512, 512)
    assert (ref[..., 3] > 0).mean() > 0.3
    assert np.array_equal(rast.cpu().numpy(), ref)
    assert np.array_equal(attr.cpu().numpy(), ob.interpolate(v, ref, f))
    uvc = torch.from_numpy(np.concatenate([uv, np.zeros_like(uv[:, :1]), np.ones_like(uv[:, :1])], -1)[None]).cuda()
    r2 = ub.rasterize(uvc, torch.from_numpy(fuv).cuda(), (2048, 2048))
    ref2 = ob.rasterize(uvc.cpu().numpy(), fuv, 2048, 2048)
    assert (ref2[..., 3] > 0).sum() > 2_000_000
    assert np.array_equal(r2.cpu().numpy(), ref2)


def test_teaser_lbvh_and_rays_bit_exact(lib, teaser, parity_log):
    from oracle import bake as ob
    from unitex_b200.bake import RayTracing
    v, f, _, _ = teaser
    rt = RayTracing(torch.from_numpy(v), torch.from_numpy(f.astype(np.int64)))
    info, aabb = rt.export()
    rinfo, raabb, _ = ob.lbvh_build(v, f)
    assert np.array_equal(info.cpu().numpy(), rinfo) and np.array_equal(aabb.cpu().numpy(), raabb)
    # tree depth (the 4-wide walk's stack bound only bites on deep trees)
    depth, frontier = 0, np.array([0])                                    # nodes < F-1 are internal, the rest leaves
    while len(frontier):
        inner = frontier[frontier < len(f) - 1]
        frontier = rinfo[inner][:, :2].reshape(-1)
        depth += 1
    assert depth >= 24
    g = np.random.default_rng(1)
    N = 1_000_000
    # half: the bake's own ray family (parallel rays of the six box views towards surface points); half: random rays
    c2ws, _ = _views()
    dirs = -c2ws[:, :3, 2].numpy().astype(np.float32)
    tri = f[g.integers(0, len(f), N // 2)]
    w = g.dirichlet(np.ones(3), N // 2).astype(np.float32)
    p = (v[tri[:, 0]] * w[:, :1] + v[tri[:, 1]] * w[:, 1:2] + v[tri[:, 2]] * w[:, 2:3]).astype(np.float32)
    d0 = dirs[g.integers(0, 6, N // 2)]
    o0 = (p - np.float32(2.0 * np.sqrt(3.0)) * d0).astype(np.float32)
    o1 = (g.normal(size=(N // 2, 3)) * 1.5).astype(np.float32)
    d1 = (g.normal(size=(N // 2, 3)) * 0.3 - o1).astype(np.float32)
    o, d = np.concatenate([o0, o1]), np.concatenate([d0, d1])
    hit, _, tid, loc, uv = rt.intersects_closest(torch.from_numpy(o), torch.from_numpy(d))
    torch.cuda.synchronize()
    rh, rtid, rpos, ruv = ob.intersect(v, f, rinfo, raabb, o, d)
    assert 0.5 < rh.mean() < 0.999
    assert np.array_equal(hit.cpu().numpy(), rh) and np.array_equal(tid.cpu().numpy(), rtid.astype(np.int64))
    assert np.array_equal(loc.cpu().numpy(), rpos) and np.array_equal(uv.cpu().numpy(), ruv)
    parity_log(f"teaser_robot LBVH: {len(rinfo)} nodes, depth {depth}; {N} rays, hit rate {rh.mean():.3f}: ids/loc/uv bit-exact")


def test_teaser_uv_bake_masks_exact(lib, teaser, parity_log):
    from oracle import bake as ob
    from unitex_b200 import bake as ub
    v, f, uv, fuv = teaser
    c2ws, intr = _views()
    H = W = 512
    H2 = W2 = 2048
    mesh = ub.BakeMesh(v, f, uv, fuv)
    r = ub.NVDiffRendererInverse(pbr_mesh=mesh)
    mats = torch.matmul(ub.intr_to_proj(intr, perspective=False), ub.c2w_to_w2c(c2ws)).cuda()
    rast = ub.rasterize(ub.transform_points(mesh.vertices, mats), mesh.faces, (H, W))
    pos = ub.interpolate(mesh.vertices, rast, mesh.faces)
    img = ((0.5 + 0.4 * torch.sin(3.0 * pos + 0.3)) * (rast[..., 3:4] > 0)).cpu()
    _, vis, m2, col = r.infer(mesh, c2ws, intr, img, H=H, W=W, H2D=H2, W2D=W2, perspective=False,
                              ray_normal_angle_threhold=100.0, method="reproject", filt_gradient_points=False)
    torch.cuda.synchronize()
    nn = r.last_nn_index.cpu().long()
    ref = ob.infer(v, f, uv, fuv, c2ws, intr, img, H, W, H2, W2, nn_index_given=nn)
    assert torch.equal(r.last_rast2d.cpu(), ref["rast_2d"])
    assert torch.equal(m2.cpu(), ref["mask_2d"])
    assert torch.equal(vis.cpu(), ref["mask_2d_visiable"])                 # 6 x 2048^2 visibility bits, 12.7 M traced rays
    # the fill: exactly the covered texels no view owns, each from an owned texel ...
    owned = (ref["owner"] >= 0).reshape(-1)
    covered = ref["mask_2d"].reshape(-1)
    assert torch.equal(nn >= 0, covered & ~owned)
    assert owned[nn[nn >= 0]].all()
    # ... and the exact nearest one (lowest index on ties) on a sample of the queries, by brute force over all 2 M+ sources
    P = ref["pos_2d"][0].reshape(-1, 3)
    src_idx = torch.nonzero(owned)[:, 0]
    q = torch.nonzero(nn >= 0)[:, 0]
    g = torch.Generator().manual_seed(0)
    qs = q[torch.randperm(len(q), generator=g)[:768]]
    want = src_idx[ob.nearest_index(P[src_idx], P[qs], chunk=16)]
    assert torch.equal(nn[qs], want)
    err = (col.cpu() - ref["color_2d"]).abs().max().item()
    assert err < COLOR_ATOL, err
    parity_log(f"teaser_robot bake 2048^2: covered {int(covered.sum())}, owned {int(owned.sum())}, filled {len(q)}; rast/vis masks "
               f"bit-exact, 1-NN exact on {len(qs)} sampled queries, colour max|d| {err:.2e}")
