"""GPU parity of the bake path against the C/torch oracle: bit-exact for triangle ids, barycentrics, tree topology,
visibility masks and nearest-neighbour indices; fp32 tolerance for colours (north_star: "bit-exact UV indices")."""
import numpy as np
import pytest
import torch

from tests.bake_meshes import analytic_color, two_spheres, uv_sphere

pytestmark = pytest.mark.gpu

COLOR_ATOL = 2e-4     # fp32 colours: different summation order in the 7x7 blur / bilinear fetch


def _views():
    from unitex_b200.bake import generate_box_views_c2ws, generate_intrinsics
    return generate_box_views_c2ws(2.8)[[0, 1, 4, 2, 3, 5]], generate_intrinsics(1.0, 1.0, fov=False)


@pytest.mark.parametrize("res", [(64, 64), (200, 136)])
def test_rasterize_interpolate_bit_exact(lib, res):
    from oracle import bake as ob
    from unitex_b200 import bake as ub
    v, f, uv, fuv = two_spheres(14, 28)
    c2ws, intr = _views()
    mats = torch.matmul(ub.intr_to_proj(intr, perspective=False), ub.c2w_to_w2c(c2ws))
    clip = ub.transform_points(torch.from_numpy(v).cuda(), mats.cuda())
    H, W = res
    rast = ub.rasterize(clip, torch.from_numpy(f).cuda(), (H, W))
    attr = ub.interpolate(torch.from_numpy(v).cuda(), rast, torch.from_numpy(f).cuda())
    torch.cuda.synchronize()
    ref = ob.rasterize(clip.cpu().numpy(), f, H, W)
    assert (ref[..., 3] > 0).sum() > 0.1 * ref[..., 3].size
    assert np.array_equal(rast.cpu().numpy(), ref)                       # ids, u, v, z/w: bit for bit
    assert np.array_equal(attr.cpu().numpy(), ob.interpolate(v, ref, f))
    # UV-space pass (shared geometry, all z = 0)
    uvc = torch.from_numpy(np.concatenate([uv, np.zeros_like(uv[:, :1]), np.ones_like(uv[:, :1])], -1)[None]).cuda()
    r2 = ub.rasterize(uvc, torch.from_numpy(fuv).cuda(), (128, 128))
    assert np.array_equal(r2.cpu().numpy(), ob.rasterize(uvc.cpu().numpy(), fuv, 128, 128))


def test_lbvh_and_intersect_bit_exact(lib):
    from oracle import bake as ob
    from unitex_b200.bake import RayTracing
    v, f, _, _ = two_spheres(18, 36)
    rt = RayTracing(torch.from_numpy(v), torch.from_numpy(f.astype(np.int64)))
    info, aabb = rt.export()
    rinfo, raabb, _ = ob.lbvh_build(v, f)
    assert np.array_equal(info.cpu().numpy(), rinfo) and np.array_equal(aabb.cpu().numpy(), raabb)
    g = np.random.default_rng(0)
    N = 20000
    o = g.normal(size=(N, 3)).astype(np.float32) * 1.5
    d = (g.normal(size=(N, 3)) * 0.3 - o).astype(np.float32)            # roughly towards the scene; some miss
    d[:500] = np.array([0, 0, -1], np.float32)                          # axis-aligned rays hit the 1e-6 zero-direction path
    hit, _, tid, loc, uv = rt.intersects_closest(torch.from_numpy(o), torch.from_numpy(d))
    torch.cuda.synchronize()
    rh, rtid, rpos, ruv = ob.intersect(v, f, rinfo, raabb, o, d)
    assert 0.2 < rh.mean() < 0.99
    assert np.array_equal(hit.cpu().numpy(), rh) and np.array_equal(tid.cpu().numpy(), rtid.astype(np.int64))
    assert np.array_equal(loc.cpu().numpy(), rpos) and np.array_equal(uv.cpu().numpy(), ruv)
    assert tid.dtype == torch.int64 and (tid[~hit] == -1).all()


def test_uv_bake_matches_oracle(lib):
    from oracle import bake as ob
    from unitex_b200 import bake as ub
    v, f, uv, fuv = two_spheres(16, 32)
    c2ws, intr = _views()
    H = W = 96
    H2 = W2 = 128
    mats = torch.matmul(ob.intr_to_proj_ortho(intr), ob.c2w_to_w2c(c2ws))
    vh = torch.cat([torch.from_numpy(v), torch.ones(len(v), 1)], -1)
    rast = ob.rasterize(torch.matmul(vh, mats.permute(0, 2, 1)).numpy(), f, H, W)
    img = torch.from_numpy(analytic_color(ob.interpolate(v, rast, f)) * (rast[..., 3:4] > 0)).float()
    ref = ob.infer_reproject(v, f, uv, fuv, c2ws, intr, img, H, W, H2, W2)
    r = ub.NVDiffRendererInverse(pbr_mesh=ub.BakeMesh(v, f, uv, fuv))
    _, vis, m2, col = r.infer(r.pbr_mesh, c2ws, intr, img, H=H, W=W, H2D=H2, W2D=W2, perspective=False,
                              ray_normal_angle_threhold=100.0, method="reproject", filt_gradient_points=False)
    torch.cuda.synchronize()
    assert vis.shape == (6, H2, W2, 1) and m2.shape == (1, H2, W2, 1) and col.shape == (1, H2, W2, 3)
    assert torch.equal(m2.cpu(), ref["mask_2d"])
    assert torch.equal(r.last_rast2d.cpu(), ref["rast_2d"])                                  # bit-exact UV triangle ids + barycentrics
    assert torch.equal(vis.cpu(), ref["mask_2d_visiable"])                                   # bit-exact visibility masks
    nn = r.last_nn_index.cpu().long()
    assert (ref["nn_index"] >= 0).sum() > 100
    assert torch.equal(nn, ref["nn_index"])                                                  # exact 1-NN, same tie rule
    err = (col.cpu() - ref["color_2d"]).abs()
    assert err.max().item() < COLOR_ATOL, err.max().item()


def test_raytracing_plugin_surface(lib):
    """Same call surface as texturetools.raytracing.RayTracing (raytracing/__init__.py:12-80)."""
    from unitex_b200.bake import RayTracing
    v, f, _ = uv_sphere(10, 20)
    rt = RayTracing(torch.from_numpy(v), torch.from_numpy(f.astype(np.int64)), backend="aprmis")
    o = torch.tensor([[[0.0, 0.0, 3.0], [0.0, 2.0, 3.0]]])                # batch shape [1,2]
    d = torch.tensor([0.0, 0.0, -1.0])
    hit, front, tri_idx, loc, uvv = rt.intersects_closest(o, d)
    assert front is None and hit.shape == (1, 2) and loc.shape == (1, 2, 3) and uvv.shape == (1, 2, 2)
    assert hit.tolist() == [[True, False]] and tri_idx[0, 1] == -1
    rt.update_raw(torch.from_numpy(v * 2), torch.from_numpy(f.astype(np.int64)))
    hit2, _, _, loc2, _ = rt.intersects_closest(o, d)
    assert hit2.tolist() == [[True, False]] and abs(abs(loc2[0, 0, 2].item()) - 1.2) < 0.05


def test_uv_bake_matches_committed_golden(lib):
    """CUDA path vs tests/golden/bake_two_spheres.npz: triangle ids, visibility, ownership, 1-NN indices exact."""
    import os
    from unitex_b200 import bake as ub
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "bake_two_spheres.npz"))
    v, f, uv, fuv = two_spheres(10, 20)
    c2ws, intr = _views()
    r = ub.NVDiffRendererInverse(pbr_mesh=ub.BakeMesh(v, f, uv, fuv))
    _, vis, m2, col = r.infer(r.pbr_mesh, c2ws, intr, torch.from_numpy(z["image"]), H=48, W=48, H2D=64, W2D=64, perspective=False,
                              ray_normal_angle_threhold=100.0, method="reproject", filt_gradient_points=False)
    torch.cuda.synchronize()
    tid = r.last_rast2d[0, ..., 3].cpu().numpy().astype(np.int32) - 1
    assert np.array_equal(tid[None], z["tid_2d"])
    assert np.array_equal(np.packbits(vis.cpu().numpy()), z["mask_vis"])
    assert np.array_equal(r.last_nn_index.cpu().numpy().reshape(-1), z["nn_index"])
    assert np.abs(col.cpu().numpy() - z["color_2d"].astype(np.float32)).max() < 2e-3
    info, _ = r.pbr_mesh.optix.export()
    assert np.array_equal(info.cpu().numpy(), z["lbvh_info"])


def test_knn_plugin_exact(lib):
    """knn(src, dst, k=1) drop-in (pcd/knn/__init__.py:104-114) vs exact brute force with the same tie rule."""
    from oracle import bake as ob
    from unitex_b200.bake import knn
    g = torch.Generator().manual_seed(0)
    src = torch.rand(20000, 3, generator=g)
    src[100] = src[7]                                    # an exact duplicate: lowest index must win
    dst = torch.cat([torch.rand(3000, 3, generator=g), src[100:101]])
    score, index = knn(src, dst, k=1)
    torch.cuda.synchronize()
    ref = ob.nearest_index(src, dst)
    assert index.shape == (3001, 1) and index.dtype == torch.int64
    assert torch.equal(index[:, 0].cpu(), ref) and index[-1, 0] == 7
    assert torch.allclose(score[:, 0].cpu(), (src[ref] - dst).norm(dim=-1), atol=1e-6)


@pytest.mark.parametrize("k", [2, 9, 32])
def test_knn_k_exact(lib, k):
    """knn(src, dst, k) for the kdtree bake's k (renderer_inverse.py:385,413,427): rows ascending by (distance, index)."""
    from oracle import bake as ob
    from unitex_b200.bake import knn
    g = torch.Generator().manual_seed(k)
    src = torch.rand(6000, 3, generator=g)
    src[100] = src[7]
    src[4000:4004] = src[9]                              # a run of duplicates: must come out in index order
    dst = torch.cat([torch.rand(1500, 3, generator=g), src[100:101], src[4001:4002]])
    score, index = knn(src, dst, k=k)
    torch.cuda.synchronize()
    rd, ri = ob.nearest_k(src, dst, k)
    assert index.shape == (1502, k) and index.dtype == torch.int64 and score.dtype == torch.float32
    assert torch.equal(index.cpu(), ri)
    assert torch.equal(score.cpu(), rd)                   # sqrt of the same fp32 squared distance
    assert index[-2, :2].tolist() == [7, 100]
    with pytest.raises(ValueError):
        knn(src[:4], dst, k=5)


def _bake_case(rows=16):
    from oracle import bake as ob
    v, f, uv, fuv = two_spheres(rows, 2 * rows)
    c2ws, intr = _views()
    H = W = 96
    mats = torch.matmul(ob.intr_to_proj_ortho(intr), ob.c2w_to_w2c(c2ws))
    vh = torch.cat([torch.from_numpy(v), torch.ones(len(v), 1)], -1)
    rast = ob.rasterize(torch.matmul(vh, mats.permute(0, 2, 1)).numpy(), f, H, W)
    img = torch.from_numpy(analytic_color(ob.interpolate(v, rast, f)) * (rast[..., 3:4] > 0)).float()
    return v, f, uv, fuv, c2ws, intr, img, H, W


@pytest.mark.parametrize("kdtree_method,k_vis", [("order_mean", 1), ("order_mean", 9), ("mean", 32)])
def test_uv_bake_kdtree_matches_oracle(lib, kdtree_method, k_vis):
    """infer(method='kdtree') (bake_mv_to_uv_kdtree, renderer_inverse.py:367-433) vs the oracle restatement."""
    from oracle import bake as ob
    from unitex_b200 import bake as ub
    v, f, uv, fuv, c2ws, intr, img, H, W = _bake_case()
    H2 = W2 = 128
    ref = ob.infer(v, f, uv, fuv, c2ws, intr, img, H, W, H2, W2, method="kdtree", kdtree_method=kdtree_method,
                   k_all=32, k_vis=k_vis, k_invis=32)
    r = ub.NVDiffRendererInverse(pbr_mesh=ub.BakeMesh(v, f, uv, fuv))
    _, vis, m2, col = r.infer(r.pbr_mesh, c2ws, intr, img, H=H, W=W, H2D=H2, W2D=W2, perspective=False,
                              ray_normal_angle_threhold=100.0, method="kdtree", kdtree_method=kdtree_method,
                              kdtree_n_neighbors=32, kdtree_n_neighbors_visiable=k_vis, kdtree_n_neighbors_invisiable=32,
                              filt_gradient_points=False)
    torch.cuda.synchronize()
    assert torch.equal(m2.cpu(), ref["mask_2d"]) and torch.equal(vis.cpu(), ref["mask_2d_visiable"])
    err = (col.cpu() - ref["color_2d"]).abs()
    assert err.max().item() < COLOR_ATOL, err.max().item()
    # the bake reproduces the analytic colour field on the texels the views see (the reference's test_gt idea, :732-774)
    seen = ref["mask_2d_visiable"].any(dim=0)[..., 0]
    pos = torch.from_numpy(ob.interpolate(v, ref["rast_2d"].numpy(), f))[0]
    gt = torch.from_numpy(analytic_color(pos.numpy())).float()
    assert (col.cpu()[0][seen] - gt[seen]).abs().mean().item() < 0.03


@pytest.mark.parametrize("method,kdtree_method", [("reproject", "order_mean"), ("kdtree", "order_mean"), ("kdtree", "mean")])
def test_uv_bake_query_field_hook(lib, method, kdtree_method):
    """`*_inpainting=True`: the registered query field supplies the colours (register_query_field :93-103; call sites
    :387-389, :427-432, :609-614) -- same arguments, in the same order, as the oracle hands its field."""
    from oracle import bake as ob
    from unitex_b200 import bake as ub
    v, f, uv, fuv, c2ws, intr, img, H, W = _bake_case(12)
    H2 = W2 = 64
    calls = []

    def field(vv, cv, vi):
        calls.append((vv.detach().cpu(), cv.detach().cpu(), vi.detach().cpu()))
        return (0.25 + 0.5 * torch.sigmoid(vi * 3.0)).to(vi)      # a colour that depends only on the query position

    ref = ob.infer(v, f, uv, fuv, c2ws, intr, img, H, W, H2, W2, method=method, kdtree_method=kdtree_method, query_field=field)
    ref_call = calls.pop()
    r = ub.NVDiffRendererInverse(pbr_mesh=ub.BakeMesh(v, f, uv, fuv))
    with pytest.raises(NotImplementedError):
        r.infer(r.pbr_mesh, c2ws, intr, img, H=H, W=W, H2D=H2, W2D=W2, perspective=False, method=method,
                kdtree_method=kdtree_method, kdtree_inpainting=True, reproject_inpainting=True, filt_gradient_points=False)
    r.register_query_field(field)
    _, vis, m2, col = r.infer(r.pbr_mesh, c2ws, intr, img, H=H, W=W, H2D=H2, W2D=W2, perspective=False,
                              ray_normal_angle_threhold=100.0, method=method, kdtree_method=kdtree_method,
                              kdtree_inpainting=True, reproject_inpainting=True, filt_gradient_points=False)
    torch.cuda.synchronize()
    assert len(calls) == 1
    for got, want in zip(calls[0], ref_call):
        assert got.shape == want.shape and (got - want).abs().max().item() < COLOR_ATOL
    assert torch.equal(calls[0][2], ref_call[2])                  # query positions: bit-exact, same (row-major) order
    assert torch.equal(vis.cpu(), ref["mask_2d_visiable"])
    err = (col.cpu() - ref["color_2d"]).abs()
    assert err.max().item() < COLOR_ATOL, err.max().item()


def test_kdtree_bake_matches_committed_golden(lib):
    """CUDA path vs tests/golden/bake_kdtree.npz: k-NN indices and distances exact, kdtree-baked colours within fp16 storage."""
    import os
    from unitex_b200 import bake as ub
    G = os.path.join(os.path.dirname(__file__), "golden")
    z, zk = np.load(os.path.join(G, "bake_two_spheres.npz")), np.load(os.path.join(G, "bake_kdtree.npz"))
    score, index = ub.knn(torch.from_numpy(zk["knn_src"]), torch.from_numpy(zk["knn_dst"]), k=8)
    assert np.array_equal(index.cpu().numpy().astype(np.int32), zk["knn_index"])
    assert np.array_equal(score.cpu().numpy(), zk["knn_dist"])
    v, f, uv, fuv = two_spheres(10, 20)
    c2ws, intr = _views()
    r = ub.NVDiffRendererInverse(pbr_mesh=ub.BakeMesh(v, f, uv, fuv))
    _, _, _, col = r.infer(r.pbr_mesh, c2ws, intr, torch.from_numpy(z["image"]), H=48, W=48, H2D=64, W2D=64, perspective=False,
                           ray_normal_angle_threhold=100.0, method="kdtree", kdtree_n_neighbors_visiable=9,
                           kdtree_n_neighbors_invisiable=32, filt_gradient_points=False)
    torch.cuda.synchronize()
    assert np.abs(col.cpu().numpy() - zk["color_2d"].astype(np.float32)).max() < 2e-3


@pytest.mark.parametrize("perspective", [False, True])
@pytest.mark.parametrize("res", [(96, 96), (64, 300)])
def test_mv_visibility_filter_matches_torch_chain(lib, perspective, res):
    """mv_to_pcd(filt_gradient_points=True): the one-kernel filter against the reference's torch op chain (renderer_inverse.py:
    186-214) run on the same rasters -- including rows wider than one 256-pixel block segment and pinhole views."""
    import math
    from unitex_b200 import bake as ub
    F_ = torch.nn.functional
    v, f, uv, fuv = two_spheres(14, 28)
    if perspective:
        c2ws, intr = ub.generate_box_views_c2ws(2.8)[[0, 1, 4, 2, 3, 5]], ub.generate_intrinsics(49.1, 49.1, fov=True, degree=True)
    else:
        c2ws, intr = _views()
    r = ub.NVDiffRendererInverse(pbr_mesh=ub.BakeMesh(v, f, uv, fuv))
    H, W = res
    thr, ang = 0.2, 100.0
    mv = r.mv_to_pcd(c2ws, intr, (H, W), perspective=perspective, grad_norm_threhold=thr, ray_normal_angle_threhold=ang,
                     filt_gradient_points=True)
    m, rast = r.pbr_mesh, mv["rast"]
    n = c2ws.shape[0]
    mask = rast[..., 3:4] > 0
    attrs = ub.interpolate(torch.cat([m.vertices, m.vertex_normals], dim=-1).contiguous(), rast, m.faces)
    a_dy, a_dx = torch.gradient(attrs, dim=(1, 2))
    gnorm = (a_dx.square() + a_dy.square()).sum(dim=-1, keepdim=True).sqrt()
    tid = rast[..., 3:4].to(torch.int64).sub(1)
    fn = m.normals.gather(0, torch.where(mask, tid, 0).reshape(-1, 1).repeat(1, 3)).reshape(n, H, W, 3)
    c2 = c2ws.cuda().float()
    rays_d = attrs[..., 0:3] - c2[:, :3, 3][:, None, None] if perspective else c2[:, :3, 2].neg()[:, None, None]
    rays_d = torch.broadcast_tensors(F_.normalize(rays_d, dim=-1), fn)[0]
    cos = F_.cosine_similarity(rays_d, fn, dim=-1).unsqueeze(-1)
    eroded = (1.0 - F_.max_pool2d(1.0 - (gnorm < thr).float(), kernel_size=31, stride=1, padding=15)).bool()
    want = mask & (cos < math.cos(math.radians(ang))) & eroded
    got = mv["mask_visiable"]
    assert got.shape == want.shape and got.dtype == torch.bool
    assert 100 < int(want.sum()) < int(mask.sum())        # the filter removes something and keeps something
    assert int((got != want).sum()) == 0
    assert torch.equal(mv["alpha_visiable"], want.float())


@pytest.mark.parametrize("k", [1, 8, 32])
def test_mvpaint_blend_matches_torch_chain(lib, k):
    """utx_mvpaint_blend against the reference's weighting expression (renderer_inverse.py:390-399), with exact-zero distances
    (1 / 0 -> the largest finite value under nan_to_num) and opposed normals (negative weights) in the table."""
    import ctypes as C
    from unitex_b200 import _lib
    F_ = torch.nn.functional
    g = torch.Generator(device="cuda").manual_seed(k)
    N, M = 5000, 3000
    cloud_c = torch.rand(N, 3, device="cuda", generator=g)
    cloud_n = F_.normalize(torch.randn(N, 3, device="cuda", generator=g), dim=-1)
    tex_n = F_.normalize(torch.randn(M, 3, device="cuda", generator=g), dim=-1)
    index = torch.randint(0, N, (M, k), device="cuda", generator=g)
    score = torch.rand(M, k, device="cuda", generator=g).sort(dim=-1).values
    score[::7, 0] = 0.0
    out = torch.empty(M, 3, device="cuda")
    L = _lib.load()
    _lib.check(L.utx_mvpaint_blend(C.c_void_p(score.data_ptr()), C.c_void_p(index.data_ptr()), M, k, C.c_void_p(cloud_c.data_ptr()),
                                   C.c_void_p(cloud_n.data_ptr()), C.c_void_p(tex_n.data_ptr()), C.c_void_p(out.data_ptr()), None),
               "utx_mvpaint_blend")
    torch.cuda.synchronize()
    weight = F_.normalize(score.reciprocal().nan_to_num(nan=0.0), p=1, dim=-1) * \
        F_.cosine_similarity(cloud_n[index], tex_n.unsqueeze(-2), dim=-1)
    weight = weight.unsqueeze(-1)
    want = torch.nan_to_num((cloud_c[index] * weight).sum(dim=-2) / weight.sum(dim=-2), nan=0.0, posinf=0.0, neginf=0.0)
    # sum w can cancel (opposed normals): compare where the denominator is well conditioned, and finiteness everywhere
    cond = weight.sum(dim=-2).abs().squeeze(-1) > 1e-3 * weight.abs().sum(dim=-2).squeeze(-1)
    assert torch.isfinite(out).all() and cond.float().mean() > 0.9
    assert (out[cond] - want[cond]).abs().max().item() < 1e-3 * (1.0 + want[cond].abs().max().item())
