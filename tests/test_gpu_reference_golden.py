"""GPU: the CUDA path against fixtures produced by the REFERENCE'S OWN Python (tests/golden/ref_*.npz; generator
tests/golden/make_reference_golden.py, run in the build container where /root/reference exists -- nothing here reads it).
Bake: every `infer` variant of NVDiffRendererInverse (renderer_inverse.py:635-726): masks bit-exact, colours |d| < 2e-4.
FLUX: PBRFluxPipeline.__call__ (flux_piplines/*/pipeline.py:404-700) from PIL images: latents PSNR >= 40 dB (north_star)."""
import os

import numpy as np
import pytest
import torch

from tests.bake_meshes import two_spheres

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")
COLOR_ATOL = 2e-4


def _unpack(bits, shape):
    return np.unpackbits(bits)[: int(np.prod(shape))].reshape(shape).astype(bool)


def test_bake_variants_match_reference_infer(lib):
    from unitex_b200 import bake as ub
    z, zi = np.load(os.path.join(G, "ref_bake.npz")), np.load(os.path.join(G, "bake_two_spheres.npz"))
    v, f, uv, fuv = two_spheres(10, 20)
    c2ws = ub.generate_box_views_c2ws(2.8)
    assert np.array_equal(c2ws.cpu().numpy(), z["c2ws_all"])
    c2ws, intr = c2ws[[0, 1, 4, 2, 3, 5]], ub.generate_intrinsics(1.0, 1.0, fov=False)
    img = torch.from_numpy(zi["image"])
    calls = []

    def field(vv, cv, vi):
        calls.append((vv.shape[0], vi.double().sum(0).cpu().numpy()))
        return (0.25 + 0.5 * torch.sigmoid(vi * 3.0)).to(vi)

    r = ub.NVDiffRendererInverse(pbr_mesh=ub.BakeMesh(v, f, uv, fuv))
    r.register_query_field(field)
    common = dict(H=48, W=48, H2D=64, W2D=64, perspective=False, ray_normal_angle_threhold=100.0, filt_gradient_points=False)
    variants = {
        "reproject": dict(method="reproject"),
        "kdtree_order_mean": dict(method="kdtree", kdtree_method="order_mean", kdtree_n_neighbors_visiable=9, kdtree_n_neighbors_invisiable=32),
        "kdtree_mean": dict(method="kdtree", kdtree_method="mean", kdtree_n_neighbors=32),
        "kdtree_mvpaint": dict(method="kdtree", kdtree_method="mvpaint", kdtree_n_neighbors=8),
        "reproject_gaussian": dict(method="reproject", reproject_method="gaussian"),
        "reproject_inpaint": dict(method="reproject", reproject_inpainting=True),
        "kdtree_inpaint": dict(method="kdtree", kdtree_method="order_mean", kdtree_n_neighbors_visiable=9, kdtree_inpainting=True),
    }
    for name, kw in variants.items():
        _, vis, m2, col = r.infer(r.pbr_mesh, c2ws, intr, img, **common, **kw)
        torch.cuda.synchronize()
        err = np.abs(col.cpu().numpy() - z[f"{name}.color_2d"]).max()
        assert err < COLOR_ATOL, (name, err)
        assert np.array_equal(vis.cpu().numpy(), _unpack(z["mask_2d_visiable"], (6, 64, 64, 1))), name
        assert np.array_equal(m2.cpu().numpy(), _unpack(z["mask_2d"], (1, 64, 64, 1))), name
    on = (img.sum(-1, keepdim=True) > 0)
    img9 = torch.cat([img, 0.5 * img + 0.1 * on, (1.0 - img) * on], dim=-1)                 # PBR attributes, 9 channels (:711-719)
    for name in ("reproject", "kdtree_order_mean"):
        _, vis, m2, col = r.infer(r.pbr_mesh, c2ws, intr, img9, **common, **variants[name])
        torch.cuda.synchronize()
        assert col.shape == (1, 64, 64, 9)
        err = np.abs(col.cpu().numpy() - z[f"{name}.pbr9.color_2d"]).max()
        assert err < COLOR_ATOL, (name, "pbr9", err)
    assert [c[0] for c in calls] == z["field.n_visible"].tolist()          # the field sees the same visible / query sets
    assert np.array_equal(np.stack([c[1] for c in calls]), z["field.query_sum"])   # ... bit-identical query positions, same order
    mv = r.mv_to_pcd(c2ws, intr, (48, 48), image_attrs=img, perspective=False, filt_gradient_points=False)
    assert np.array_equal(mv["alpha_visiable"].cpu().numpy() > 0, _unpack(z["mv.alpha_visiable"], (6, 48, 48, 1)))
    score, index = ub.knn(torch.from_numpy(z["fn.knn_src"]), torch.from_numpy(z["fn.knn_dst"]), k=4)
    assert np.array_equal(index.cpu().numpy().astype(np.int32), z["fn.knn_index"])


def test_pipeline_call_matches_reference_call(lib):
    from PIL import Image
    from flux_piplines.delight.pipeline import PBRFluxPipeline as DelightPipeline
    from flux_piplines.texturing.pipeline import PBRFluxPipeline
    from oracle import flux_dit as fd
    from oracle import flux_sampler as fs
    from oracle import vae as ov
    from unitex_b200.flux import FluxConfig, FluxTransformer
    from unitex_b200.vae import AutoencoderKLB200
    z = np.load(os.path.join(G, "ref_flux_call.npz"))
    ocfg = fd.FluxConfig(num_layers=1, num_single_layers=1, num_attention_heads=2)     # text widths 4096 / 768 as the reference hard-codes
    P = {k: v.to(torch.bfloat16).float() for k, v in fd.init_params(ocfg, 21, norm_weight_std=0.1).items()}
    vcfg = ov.VaeConfig.tiny()
    VP = {k: v.to(torch.bfloat16).float() for k, v in ov.init_params(vcfg, 5).items()}
    vae = AutoencoderKLB200(VP, vcfg.block_out_channels, vcfg.layers_per_block, vcfg.latent_channels, vcfg.in_channels,
                            vcfg.norm_num_groups, vcfg.scaling_factor, vcfg.shift_factor)
    tr = FluxTransformer(FluxConfig(num_layers=1, num_single_layers=1, num_attention_heads=2)).load_state_dict(P)
    ctrl, dual = Image.fromarray(z["control_image"]), Image.fromarray(z["dual_image"])
    for task, cls, d in (("texturing", PBRFluxPipeline, dual), ("delight", DelightPipeline, None)):
        pipe = cls(tr, vae)
        kw = dict(prompt="[MVFLUX]", control_image=ctrl, dual_image=d, height=128, width=128, n_rows=1, n_cols=1,
                  num_inference_steps=3, guidance_scale=3.5, max_sequence_length=128)
        lat = pipe(**kw, generator=torch.Generator().manual_seed(63), output_type="latent").images
        torch.cuda.synchronize()
        ref = torch.from_numpy(z[f"{task}.latents"])
        assert lat.shape == ref.shape
        db = fs.psnr(lat.float().cpu(), ref)
        assert db >= 40.0, f"{task}: latent PSNR {db:.1f} dB vs the reference's own __call__"
        img = pipe(**kw, generator=torch.Generator().manual_seed(63)).images[0]
        a, b = np.asarray(img).astype(np.float64), z[f"{task}.image"].astype(np.float64)
        assert a.shape == b.shape
        db_img = 10 * np.log10(255.0 ** 2 / max(np.mean((a - b) ** 2), 1e-12))
        assert db_img >= 30.0, f"{task}: image PSNR {db_img:.1f} dB"     # both sides are bf16 VAE chains + uint8 quantisation
        print(f"{task}: latent {db:.1f} dB, image {db_img:.1f} dB")


def test_export_condition_matches_reference(lib):
    """b1: this repo's VideoExporter.export_condition (CUDA rasteriser) against the reference's own export_condition run
    (video/export_nvdiffrast_video.py:900-999): coverage exact, uint8 G-buffers within 1 LSB."""
    from unitex_b200.export import VideoExporter
    z = np.load(os.path.join(G, "ref_glue.npz"))
    v, f, _, _ = two_spheres(10, 20)
    for name, kw in (("six", dict(n_views=6, n_rows=2, n_cols=3)), ("four", dict(n_views=4, n_rows=2, n_cols=2)),
                     ("four_persp", dict(n_views=4, n_rows=2, n_cols=2, perspective=True)),
                     ("orbit8", dict(n_views=8, n_rows=2, n_cols=4, orbit=True))):
        kw = dict(dict(perspective=False, orbit=False), **kw)
        out = VideoExporter().export_condition((v, f), geometry_scale=0.95, H=64, W=64, fov_deg=49.1, scale=1.0,
                                               background="grey", return_image=True, return_camera=True, **kw)
        assert np.array_equal(np.asarray(out["alpha"]), z[f"cond.{name}.alpha"]), name
        for k in ("ccm", "normal"):
            d = np.abs(np.asarray(out[k]).astype(np.int16) - z[f"cond.{name}.{k}"].astype(np.int16))
            assert d.max() <= 1 and (d > 0).mean() < 0.01, (name, k, int(d.max()), float((d > 0).mean()))
        assert np.array_equal(out["c2ws"].cpu().numpy(), z[f"cond.{name}.c2ws"])
        assert np.array_equal(out["intrinsics"].cpu().numpy(), z[f"cond.{name}.intrinsics"])
        assert out["perspective"] is kw["perspective"]


def test_reproject_glue_matches_reference(lib, tmp_path):
    """b11: what `reproject_and_query_field` (pipeline.py:312-360) hands to the bake entry point -- view slicing of mv_rgb.png,
    the infer kwargs, the files it leaves -- against the reference's own function driven with the same fake renderer."""
    import ast
    import types
    from PIL import Image
    import pipeline as drop_in
    from tests.glue_fakes import FakeFlux, glue_inputs, sha
    from unitex_b200 import export as ux
    z = np.load(os.path.join(G, "ref_glue.npz"))
    normal, ccm, ref = glue_inputs()
    d = str(tmp_path)
    for n, a in (("mv_normal.png", normal), ("mv_ccm.png", ccm), ("processed_image.png", ref)):
        Image.fromarray(a).save(os.path.join(d, n))
    me = types.SimpleNamespace(pipeline=FakeFlux(), pipeline_name="texture_plus", adapter_names=["texture", "delight"],
                               weights_for_texture=[1.0, 0.0], weights_for_delight=[0.0, 1.0], generator=None, super_resolutions=False)
    drop_in.CustomRGBTextureFullPipeline.infer_mv(me, d, os.path.join(d, "processed_image.png"), os.path.join(d, "mv_normal.png"),
                                                  os.path.join(d, "mv_ccm.png"))                    # writes mv_rgb.png (hash-checked on CPU)
    v, f, uv, fuv = two_spheres(6, 12)
    ux.save_obj(os.path.join(d, "processed_mesh.obj"), v, f, uv * 0.5 + 0.5, fuv)
    torch.save({"c2ws": torch.from_numpy(z["cond.six.c2ws"]), "intrinsics": torch.from_numpy(z["cond.six.intrinsics"]), "perspective": False},
               os.path.join(d, "camera_info.pth"))
    rec = {}

    class FakeInverse:
        pbr_mesh = None

        def update_from_file(self, path):
            rec["mesh"] = os.path.basename(path)

        def infer(self, blank, **kw):
            rec.update(kw)
            n = kw["c2ws"].shape[0]
            return None, torch.zeros(n, 8, 8, 1, dtype=torch.bool), torch.ones(1, 8, 8, 1, dtype=torch.bool), torch.full((1, 8, 8, 3), 0.25)

        def clear(self):
            rec["cleared"] = True
    me2 = types.SimpleNamespace(inverse_renderer=FakeInverse())
    drop_in.CustomRGBTextureFullPipeline.reproject_and_query_field(me2, d, os.path.join(d, "processed_mesh.obj"), os.path.join(d, "mv_rgb.png"),
                                                                   os.path.join(d, "camera_info.pth"), method="reproject", inpainting=False)
    assert sha(rec["image_attrs"].cpu().numpy()) == str(z["rq.image_attrs_sha"])
    assert np.array_equal(rec["image_attrs"].cpu().numpy()[:, ::64, ::64], z["rq.image_attrs_probe"])
    want = dict(ast.literal_eval(str(z["rq.kwargs"])))
    got = {k: v for k, v in rec.items() if k not in ("image_attrs", "c2ws", "intrinsics")}
    assert got == want, (got, want)
    assert np.array_equal(rec["c2ws"].cpu().numpy(), z["cond.six.c2ws"])
    assert sorted(n for n in os.listdir(d) if n.endswith((".glb", "_mask.png", "_uv.png"))) == ast.literal_eval(str(z["rq.files"]))


def test_infer_filt_gradient_points_matches_reference(lib):
    """filt_gradient_points=True (the default of the reference's infer signature): gradient-filtered view masks, then the bake."""
    from unitex_b200 import bake as ub
    z = np.load(os.path.join(G, "ref_bake.npz"))
    v, f, uv, fuv = two_spheres(10, 20)
    c2ws, intr = ub.generate_box_views_c2ws(2.8)[[0, 1, 4, 2, 3, 5]], ub.generate_intrinsics(1.0, 1.0, fov=False)
    img = torch.from_numpy(z["filt.image"].astype(np.float32))
    r = ub.NVDiffRendererInverse(pbr_mesh=ub.BakeMesh(v, f, uv, fuv))
    mv = r.mv_to_pcd(c2ws, intr, (128, 128), image_attrs=img, perspective=False, grad_norm_threhold=0.2,
                     ray_normal_angle_threhold=100.0, filt_gradient_points=True)
    want = _unpack(z["filt.alpha_visiable"], (6, 128, 128, 1))
    got = mv["alpha_visiable"].cpu().numpy() > 0
    assert (got != want).sum() == 0, int((got != want).sum())
    _, vis, m2, col = r.infer(r.pbr_mesh, c2ws, intr, img, H=128, W=128, H2D=64, W2D=64, perspective=False, grad_norm_threhold=0.2,
                              ray_normal_angle_threhold=100.0, method="reproject", filt_gradient_points=True)
    torch.cuda.synchronize()
    assert np.array_equal(vis.cpu().numpy(), _unpack(z["filt.mask_2d_visiable"], (6, 64, 64, 1)))
    assert np.abs(col.cpu().numpy() - z["filt.reproject.color_2d"]).max() < COLOR_ATOL


def test_infer_perspective_matches_reference(lib):
    """perspective=True (the other default of the reference's infer signature): pinhole views, rays leave the camera position."""
    from unitex_b200 import bake as ub
    z = np.load(os.path.join(G, "ref_bake.npz"))
    v, f, uv, fuv = two_spheres(10, 20)
    c2ws, intr = ub.generate_box_views_c2ws(2.8)[[0, 1, 4, 2, 3, 5]], ub.generate_intrinsics(49.1, 49.1, fov=True, degree=True)
    img = torch.from_numpy(z["persp.image"].astype(np.float32))
    r = ub.NVDiffRendererInverse(pbr_mesh=ub.BakeMesh(v, f, uv, fuv))
    common = dict(H=48, W=48, H2D=64, W2D=64, perspective=True, ray_normal_angle_threhold=100.0, filt_gradient_points=False)
    for name, kw in (("reproject", dict(method="reproject")),
                     ("kdtree_order_mean", dict(method="kdtree", kdtree_method="order_mean", kdtree_n_neighbors_visiable=9, kdtree_n_neighbors_invisiable=32))):
        _, vis, m2, col = r.infer(r.pbr_mesh, c2ws, intr, img, **common, **kw)
        torch.cuda.synchronize()
        assert np.array_equal(vis.cpu().numpy(), _unpack(z["persp.mask_2d_visiable"], (6, 64, 64, 1))), name
        err = np.abs(col.cpu().numpy() - z[f"persp.{name}.color_2d"]).max()
        assert err < COLOR_ATOL, (name, err)


def test_raytracing_plugin_matches_reference_wrapper(lib):
    """RayTracing(vertices, faces).intersects_closest(rays_o, rays_d) against the reference's dispatcher + APRMIS wrapper run
    (raytracing/__init__.py:12-80, rt_aprmis/__init__.py:10-86): shapes, dtypes, -1 for misses, ids / positions / uv bit for bit."""
    from unitex_b200.bake import RayTracing
    z = np.load(os.path.join(G, "ref_bake.npz"))
    v, f, _, _ = two_spheres(10, 20)
    rt = RayTracing(torch.from_numpy(v), torch.from_numpy(f).long())
    hit, front, tri_idx, loc, uv = rt.intersects_closest(torch.from_numpy(z["rt.rays_o"]), torch.from_numpy(z["rt.rays_d"]))
    assert front is None and hit.dtype == torch.bool and tri_idx.dtype == torch.int64
    assert hit.shape == (3, 50) and tri_idx.shape == (3, 50) and loc.shape == (3, 50, 3) and uv.shape == (3, 50, 2)
    assert np.array_equal(hit.cpu().numpy(), z["rt.hit"]) and 5 < int(z["rt.hit"].sum()) < 145
    assert np.array_equal(tri_idx.cpu().numpy().astype(np.int32), z["rt.tri_idx"])
    assert np.array_equal(loc.cpu().numpy(), z["rt.loc"]) and np.array_equal(uv.cpu().numpy(), z["rt.uv"])
