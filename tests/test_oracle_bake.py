"""CPU: analytic known-answer tests that pin the bake oracle (the reference has no golden vectors)."""
import math

import numpy as np
import torch

from oracle import bake as ob
from tests.bake_meshes import analytic_color, two_spheres, uv_sphere


def test_rasterizer_fullscreen_quad_covers_every_pixel_once():
    # two triangles sharing a diagonal: the top-left rule must give every pixel centre to exactly one of them
    pos = np.array([[[-1, -1, 0, 1], [1, -1, 0, 1], [1, 1, 0, 1], [-1, 1, 0, 1]]], np.float32)
    tri = np.array([[0, 1, 2], [0, 2, 3]], np.int32)
    r = ob.rasterize(pos, tri, 16, 16)
    assert (r[..., 3] > 0).all()
    ids = r[0, ..., 3].astype(int)
    assert set(np.unique(ids)) == {1, 2}
    # pixel (x=12, y=3) lies in triangle 0 (below the diagonal y=x in clip space); analytic barycentrics
    assert ids[3, 12] == 1
    px, py = (12 + 0.5) / 16 * 2 - 1, (3 + 0.5) / 16 * 2 - 1
    # P = u v0 + v v1 + w v2 with v0=(-1,-1), v1=(1,-1), v2=(1,1)  ->  w = (py+1)/2, u = (1-px)/2
    assert abs(r[0, 3, 12, 0] - (1 - px) / 2) < 1e-6 and abs(r[0, 3, 12, 1] - (1 - (1 - px) / 2 - (py + 1) / 2)) < 1e-6
    # reversing the winding changes nothing (no culling)
    r2 = ob.rasterize(pos, tri[:, ::-1].copy(), 16, 16)
    assert (r2[..., 3] == r[..., 3]).all()


def test_rasterizer_depth_and_tie_rules():
    # two overlapping full-screen triangles pairs at z = 0.5 and z = -0.5: nearer (smaller z/w) wins; equal depth -> lowest id
    q = [[-1, -1], [3, -1], [-1, 3]]
    pos = np.array([[[x, y, 0.5, 1] for x, y in q] + [[x, y, -0.5, 1] for x, y in q] + [[x, y, -0.5, 1] for x, y in q]], np.float32)
    tri = np.array([[0, 1, 2], [3, 4, 5], [6, 7, 8]], np.int32)
    r = ob.rasterize(pos, tri, 8, 8)
    assert (r[..., 3] == 2).all() and np.allclose(r[..., 2], -0.5)
    # row 0 is y_clip = -1 (nvdiffrast / GL convention)
    pos = np.array([[[-1, -1, 0, 1], [1, -1, 0, 1], [0, -0.5, 0, 1]]], np.float32)
    r = ob.rasterize(pos, np.array([[0, 1, 2]], np.int32), 8, 8)
    assert (r[0, 0, :, 3] > 0).any() and not (r[0, 4:, :, 3] > 0).any()


def test_interpolate_is_barycentric():
    v, f, uv = uv_sphere(8, 12)
    uvc = np.concatenate([uv * 2 - 1, np.zeros_like(uv[:, :1]), np.ones_like(uv[:, :1])], -1)[None]
    r = ob.rasterize(uvc, f, 64, 64)
    out = ob.interpolate(v, r, f)
    m = r[0, ..., 3] > 0
    assert m.sum() > 2000
    # interpolated positions lie (nearly) on the sphere of radius 0.6: inside the chord sag of the coarse mesh
    rad = np.linalg.norm(out[0][m], axis=-1)
    assert rad.max() <= 0.6 + 1e-5 and rad.min() > 0.6 * math.cos(math.pi / 8) - 1e-3
    assert np.abs(out[0][~m]).max() == 0


def test_lbvh_is_a_valid_tight_tree():
    v, f, _, _ = two_spheres(10, 16)
    info, aabb, srt = ob.lbvh_build(v, f)
    F = len(f)
    assert (np.diff(srt[:, 0].astype(np.int64)) >= 0).all()                     # sorted Morton codes
    assert sorted(srt[:, 1]) == list(range(F))                                   # permutation of elements
    leaves = info[F - 1:]
    assert (leaves[:, 0] == 0).all() and (leaves[:, 1] == 0).all() and sorted(leaves[:, 2]) == list(range(F))
    inner = info[:F - 1]
    children = np.concatenate([inner[:, 0], inner[:, 1]])
    assert sorted(children) == list(range(1, 2 * F - 1))                         # every node but the root has one parent
    assert np.allclose(aabb[0, :3], v[f].reshape(-1, 3).min(0)) and np.allclose(aabb[0, 3:], v[f].reshape(-1, 3).max(0))
    for n in range(F - 1):                                                       # exact unions
        a, b = inner[n, 0], inner[n, 1]
        assert (aabb[n, :3] == np.minimum(aabb[a, :3], aabb[b, :3])).all() and (aabb[n, 3:] == np.maximum(aabb[a, 3:], aabb[b, 3:])).all()


def test_intersect_known_answers_and_reference_quirks():
    # one big triangle in the z=0 plane + a far one; LBVH needs >= 2 triangles
    v = np.array([[-1, -1, 0], [1, -1, 0], [0, 1, 0], [-1, -1, -5], [1, -1, -5], [0, 1, -5]], np.float32)
    f = np.array([[0, 1, 2], [3, 4, 5]], np.int32)
    info, aabb, _ = ob.lbvh_build(v, f)
    o = np.array([[0, 0, 2], [5, 5, 2], [0.2, -0.3, 2]], np.float32)
    d = np.array([[0, 0, -3.0]] * 3, np.float32)                                # un-normalised: the tracer normalises
    hit, tid, pos, uv = ob.intersect(v, f, info, aabb, o, d)
    assert hit.tolist() == [True, False, True] and tid[1] == -1
    # quirk (b): both triangles are hit; the id reported is the LAST accepted leaf in traversal order, the position
    # is the closest one
    assert set(tid[[0, 2]].tolist()) <= {0, 1}
    assert np.allclose(pos[0], [0, 0, 0], atol=1e-6) and np.allclose(pos[2], [0.2, -0.3, 0], atol=1e-6)
    # Moller-Trumbore barycentrics of the reported triangle: P = (1-u-v) v0 + u v1 + v v2
    t = tid[2]
    p = (1 - uv[2, 0] - uv[2, 1]) * v[f[t, 0]] + uv[2, 0] * v[f[t, 1]] + uv[2, 1] * v[f[t, 2]]
    assert np.allclose(p[:2], [0.2, -0.3], atol=1e-6)
    # a box entirely behind the origin is culled by the slab test (t_min = 0) ...
    hit, _, _, _ = ob.intersect(v, f, info, aabb, np.array([[0, 0, -7.0]], np.float32), np.array([[0, 0, -1.0]], np.float32))
    assert not hit[0]
    # ... but quirk (a): when the origin sits inside a leaf box, a triangle BEHIND it is accepted (no t-range test)
    v2 = np.array([[-1, -1, -1], [1, -1, -1], [0, 1, 1], [-1, -1, -9], [1, -1, -9], [0, 1, -9]], np.float32)   # plane z = y
    info2, aabb2, _ = ob.lbvh_build(v2, f)
    hit, tid, pos, _ = ob.intersect(v2, f, info2, aabb2, np.array([[0, 0, 0.5]], np.float32), np.array([[0, 0, 1.0]], np.float32))
    assert hit[0] and tid[0] == 0 and abs(pos[0, 2]) < 1e-6          # t = -0.5


def test_tracer_equals_descending_leaf_scan():
    """What the product's leaf-grid ray kernel rests on: the reference's stack walk (push left, push right, pop right first; box
    test against the running closest t; last accepted leaf wins -- intersect_test2.slang:63-146) returns exactly what a scan of
    the LEAVES in descending Morton position with per-leaf box tests returns: same hit flags, ids, positions, barycentrics.
    Rays: the bake's own family (axis-parallel through surface points, where the 1e-6 zero-direction substitution is live),
    oblique rays, rays from inside the mesh, and misses."""
    from tests.bake_meshes import two_spheres
    from unitex_b200.bake import generate_box_views_c2ws
    v, f, _, _ = two_spheres(24, 48)
    info, aabb, _ = ob.lbvh_build(v, f)
    g = np.random.default_rng(5)
    n = 1500
    dirs = -generate_box_views_c2ws(2.8)[:, :3, 2].numpy().astype(np.float32)
    tri = f[g.integers(0, len(f), n)]
    w = g.dirichlet(np.ones(3), n).astype(np.float32)
    p = (v[tri[:, 0]] * w[:, :1] + v[tri[:, 1]] * w[:, 1:2] + v[tri[:, 2]] * w[:, 2:3]).astype(np.float32)
    d0 = dirs[g.integers(0, 6, n)]
    o0 = (p - np.float32(2.0 * np.sqrt(3.0)) * d0).astype(np.float32)
    o1 = (g.normal(size=(n, 3)) * 1.5).astype(np.float32)
    d1 = (g.normal(size=(n, 3)) * 0.3 - o1).astype(np.float32)
    o2 = (g.normal(size=(n // 2, 3)) * 0.2).astype(np.float32)                # origins inside the big sphere: hits behind and ahead
    d2 = g.normal(size=(n // 2, 3)).astype(np.float32)
    o, d = np.concatenate([o0, o1, o2]), np.concatenate([d0, d1, d2])
    a = ob.intersect(v, f, info, aabb, o, d)
    b = ob.intersect_leafscan(v, f, info, aabb, o, d)
    assert 0.3 < a[0].mean() < 0.99
    for x, y in zip(a, b):
        assert np.array_equal(x, y)


def test_pull_push_and_lens_blur_properties():
    H = 64
    img = torch.full((1, 3, H, H), 0.37)
    mask = torch.zeros(1, 1, H, H, dtype=torch.bool)
    mask[:, :, 10:30, 12:40] = True
    out, _ = ob.pull_push(img, mask)
    assert torch.equal(out[:, :, 10:30, 12:40], img[:, :, 10:30, 12:40])         # masked texels untouched
    filled = out[0, 0][~mask[0, 0]]
    assert (filled >= 0).all() and filled.max() <= 0.37 + 1e-6                  # fill by the constant, fading far away
    assert abs(out[0, 0, 9, 20].item() - 0.37) < 0.1
    b = ob.lens_blur_torch(torch.full((1, 3, 32, 32), 0.5))
    assert torch.allclose(b[:, :, 8:24, 8:24], torch.full((1, 3, 16, 16), 0.5), atol=2e-4)   # kernel sums to 1


def test_bake_roundtrip_recovers_analytic_colour():
    """Shape of the reference's own test_gt (renderer_inverse.py:732-774): render a known colour field from the 6 box
    views, bake, compare in the atlas."""
    from unitex_b200.bake import generate_box_views_c2ws, generate_intrinsics
    v, f, uv, fuv = two_spheres(16, 32)
    c2ws = generate_box_views_c2ws(2.8)[[0, 1, 4, 2, 3, 5]]                       # export_nvdiffrast_video.py:934-936
    intr = generate_intrinsics(1.0, 1.0, fov=False)
    H = W = 96
    mats = torch.matmul(ob.intr_to_proj_ortho(intr), ob.c2w_to_w2c(c2ws))
    vh = torch.cat([torch.from_numpy(v), torch.ones(len(v), 1)], -1)
    clip = torch.matmul(vh, mats.permute(0, 2, 1)).numpy()
    rast = ob.rasterize(clip, f, H, W)
    pos_img = ob.interpolate(v, rast, f)
    img = torch.from_numpy(analytic_color(pos_img) * (rast[..., 3:4] > 0)).float()
    out = ob.infer_reproject(v, f, uv, fuv, c2ws, intr, img, H, W, 128, 128)
    m2, vis = out["mask_2d"][0, ..., 0], out["mask_2d_visiable"]
    assert m2.sum() > 3000 and vis.any(dim=0)[..., 0].sum() > 0.5 * m2.sum()
    assert not (vis.any(dim=0)[..., 0] & ~m2).any()
    owned = (out["owner"] >= 0) & ~out["seam"][0, ..., 0]
    pos2d = ob.interpolate(v, out["rast_2d"].numpy(), f)[0]
    want = torch.from_numpy(analytic_color(pos2d)).float()
    err = (out["color_2d"][0] - want).abs().max(-1).values[owned]
    assert err.mean() < 0.02 and torch.quantile(err, 0.99) < 0.08               # bilinear resampling of a smooth field
    assert (out["color_2d"][0][~m2] >= 0).all()                                  # pull-push filled outside the charts
    assert ((out["nn_index"] >= 0).reshape(128, 128) == (m2 & (out["owner"] < 0))).all()


def test_nearest_k_known_answers():
    """knn contract (pcd/knn/__init__.py:104-114): k nearest, ascending distance; ties by index (our pinned rule)."""
    src = torch.tensor([[0.0, 0, 0], [1.0, 0, 0], [2.0, 0, 0], [1.0, 0, 0], [5.0, 0, 0]])
    dst = torch.tensor([[0.9, 0, 0], [10.0, 0, 0]])
    d, i = ob.nearest_k(src, dst, 3)
    assert i.tolist() == [[1, 3, 0], [4, 2, 1]]
    assert torch.allclose(d, torch.tensor([[0.1, 0.1, 0.9], [5.0, 8.0, 9.0]]), atol=1e-6)
    assert torch.equal(ob.nearest_k(src, dst, 1)[1][:, 0], ob.nearest_index(src, dst))


def _kd_case():
    from unitex_b200.bake import generate_box_views_c2ws, generate_intrinsics
    v, f, uv, fuv = two_spheres(12, 24)
    c2ws = generate_box_views_c2ws(2.8)[[0, 1, 4, 2, 3, 5]]
    intr = generate_intrinsics(1.0, 1.0, fov=False)
    H = W = 64
    mats = torch.matmul(ob.intr_to_proj_ortho(intr), ob.c2w_to_w2c(c2ws))
    vh = torch.cat([torch.from_numpy(v), torch.ones(len(v), 1)], -1)
    rast = ob.rasterize(torch.matmul(vh, mats.permute(0, 2, 1)).numpy(), f, H, W)
    img = torch.from_numpy(analytic_color(ob.interpolate(v, rast, f)) * (rast[..., 3:4] > 0)).float()
    return v, f, uv, fuv, c2ws, intr, img, H, W


def test_kdtree_bake_recovers_analytic_colour():
    """bake_mv_to_uv_kdtree (renderer_inverse.py:367-433): nearest pixel-cloud points carry the analytic colour of their own
    position, so the baked atlas approximates the field for both merge rules."""
    v, f, uv, fuv, c2ws, intr, img, H, W = _kd_case()
    pos2d = None
    errs = {}
    for km in ("order_mean", "mean"):
        out = ob.infer(v, f, uv, fuv, c2ws, intr, img, H, W, 64, 64, method="kdtree", kdtree_method=km)
        m2 = out["mask_2d"][0, ..., 0]
        if pos2d is None:
            pos2d = ob.interpolate(v, out["rast_2d"].numpy(), f)[0]
        want = torch.from_numpy(analytic_color(pos2d)).float()
        seen = out["mask_2d_visiable"].any(dim=0)[..., 0]
        errs[km] = (out["color_2d"][0] - want).abs().max(-1).values[seen].mean().item()
        assert torch.equal(out["color_2d"][0][m2], out["pre_pull_push"][0][m2])          # pull-push leaves the charts alone
        assert (out["color_2d"][0][~m2] >= 0).all()
    assert errs["order_mean"] < 0.03 and errs["mean"] < 0.03


def test_query_field_hook_contract():
    """register_query_field (:93-103): f(vertices_visiable [Nv,3], colors_visiable [Nv,C], vertices_invisiable [Ni,3]) ->
    [Ni,C]; the reproject bake hands it the owned / unowned texels (:609-614), `order_mean` the same split (:427-432),
    `mean` the union pixel cloud and every covered texel (:387-389)."""
    v, f, uv, fuv, c2ws, intr, img, H, W = _kd_case()
    seen_args = {}

    def field(vv, cv, vi):
        seen_args["shapes"] = (tuple(vv.shape), tuple(cv.shape), tuple(vi.shape))
        return torch.full((vi.shape[0], 3), 0.125)

    base = ob.infer(v, f, uv, fuv, c2ws, intr, img, H, W, 64, 64)
    owned, m2 = base["owner"] >= 0, base["mask_2d"][0, ..., 0]
    n_own, n_un = int((owned & m2).sum()), int((~owned & m2).sum())
    assert n_un > 0
    out = ob.infer(v, f, uv, fuv, c2ws, intr, img, H, W, 64, 64, query_field=field)
    assert seen_args["shapes"] == ((n_own, 3), (n_own, 3), (n_un, 3))
    unowned_inner = ~owned & m2 & ~base["seam"][0, ..., 0]
    assert (out["pre_blur"][0][~owned & m2] == 0.125).all() and unowned_inner.any()
    out = ob.infer(v, f, uv, fuv, c2ws, intr, img, H, W, 64, 64, method="kdtree", query_field=field)
    assert seen_args["shapes"] == ((n_own, 3), (n_own, 3), (n_un, 3))
    assert (out["pre_pull_push"][0][~owned & m2] == 0.125).all()
    out = ob.infer(v, f, uv, fuv, c2ws, intr, img, H, W, 64, 64, method="kdtree", kdtree_method="mean", query_field=field)
    n_cloud = int((img.abs().sum(-1) > 0).sum())
    assert seen_args["shapes"][2] == (int(m2.sum()), 3) and seen_args["shapes"][0][0] >= n_cloud
    assert (out["pre_pull_push"][0][m2] == 0.125).all()
