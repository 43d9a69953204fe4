"""CPU: the oracle against fixtures produced by the REFERENCE'S OWN Python (tests/golden/ref_*.npz, generated in the build
container by tests/golden/make_reference_golden.py through ref_harness.py: the reference's renderer_inverse / PBRMesh / knn /
pull_push / lens_blur / camera code and its PBRFluxPipeline.__call__ / attention processor, with the absent third-party
packages supplied by the oracle's restatements).  Everything that is torch on both sides is compared bit for bit."""
import os

import numpy as np
import torch

from oracle import bake as ob
from oracle import flux_dit as fd
from oracle import flux_sampler as fs
from tests.bake_meshes import two_spheres

G = os.path.join(os.path.dirname(__file__), "golden")


def _unpack(bits, shape):
    return np.unpackbits(bits)[: int(np.prod(shape))].reshape(shape).astype(bool)


def test_cameras_match_reference():
    from unitex_b200.bake import c2w_to_w2c, generate_box_views_c2ws, generate_intrinsics, intr_to_proj
    z = np.load(os.path.join(G, "ref_bake.npz"))
    c2ws = generate_box_views_c2ws(2.8)
    assert np.array_equal(c2ws.numpy(), z["c2ws_all"])                                   # camera/generator.py:153-185
    intr = generate_intrinsics(1.0, 1.0, fov=False)
    assert np.array_equal(intr.numpy(), z["intrinsics"])                                 # :93-114
    assert np.array_equal(intr_to_proj(intr, perspective=False).numpy(), z["proj"])      # camera/conversion.py:8-28
    assert np.array_equal(ob.intr_to_proj_ortho(intr).numpy(), z["proj"])
    assert np.array_equal(c2w_to_w2c(c2ws[[0, 1, 4, 2, 3, 5]]).numpy(), z["w2c"])        # :50-57
    assert np.array_equal(ob.c2w_to_w2c(c2ws[[0, 1, 4, 2, 3, 5]]).numpy(), z["w2c"])


def test_mesh_normals_match_reference():
    from unitex_b200.export import vertex_normals
    z = np.load(os.path.join(G, "ref_bake.npz"))
    v, f, _, _ = two_spheres(10, 20)
    n = vertex_normals(torch.from_numpy(v), torch.from_numpy(f).long())                  # structure_v2.py:63-71
    assert np.abs(n.numpy() - z["vertex_normals"]).max() < 1e-6
    from unitex_b200.bake import area_weighted_vertex_normals                             # the bake's own copy: slot-wise like the reference
    assert np.array_equal(area_weighted_vertex_normals(torch.from_numpy(v), torch.from_numpy(f)).numpy(), z["vertex_normals"])


def test_torch_tail_functions_bit_exact():
    z = np.load(os.path.join(G, "ref_bake.npz"))
    x, m = torch.from_numpy(z["fn.x"]), torch.from_numpy(z["fn.mask"])
    assert np.array_equal(ob.lens_blur_torch(x).numpy(), z["fn.lens_blur"])              # image/lens_blur.py:260-280
    kd, km = ob.pull_push(x * m, m)                                                      # texture/stitching/mip.py:51-96
    assert np.array_equal(kd.numpy(), z["fn.pull_push"]) and np.array_equal(km.numpy(), z["fn.pull_push_mask"])
    assert np.array_equal(ob.boundary_mask(torch.from_numpy(z["fn.bmask_in"]), 3).numpy(), z["fn.bmask"])     # renderer_inverse.py:435-444
    _, idx = ob.nearest_k(torch.from_numpy(z["fn.knn_src"]), torch.from_numpy(z["fn.knn_dst"]), 4)
    assert np.array_equal(idx.numpy().astype(np.int32), z["fn.knn_index"])


def test_infer_every_variant_matches_reference():
    """oracle/bake.py::infer against NVDiffRendererInverse.infer itself (renderer_inverse.py:635-726)."""
    from unitex_b200.bake import generate_box_views_c2ws, generate_intrinsics
    z, zi = np.load(os.path.join(G, "ref_bake.npz")), np.load(os.path.join(G, "bake_two_spheres.npz"))
    v, f, uv, fuv = two_spheres(10, 20)
    c2ws = generate_box_views_c2ws(2.8)[[0, 1, 4, 2, 3, 5]]
    intr = generate_intrinsics(1.0, 1.0, fov=False)
    img = torch.from_numpy(zi["image"])
    calls = []

    def field(pv, cv, pi):
        calls.append((pv.shape[0], pi.double().sum(0).numpy()))
        return 0.25 + 0.5 * torch.sigmoid(pi * 3.0)

    variants = {
        "reproject": dict(method="reproject"),
        "kdtree_order_mean": dict(method="kdtree", kdtree_method="order_mean", k_vis=9, k_invis=32),
        "kdtree_mean": dict(method="kdtree", kdtree_method="mean", k_all=32),
        "kdtree_mvpaint": dict(method="kdtree", kdtree_method="mvpaint", k_all=8),
        "reproject_gaussian": dict(method="reproject", reproject_method="gaussian"),
        "reproject_inpaint": dict(method="reproject", query_field=field),
        "kdtree_inpaint": dict(method="kdtree", kdtree_method="order_mean", k_vis=9, query_field=field),
    }
    for name, kw in variants.items():
        out = ob.infer(v, f, uv, fuv, c2ws, intr, img, 48, 48, 64, 64, **kw)
        assert np.array_equal(out["color_2d"].numpy(), z[f"{name}.color_2d"]), name
        if name == "reproject":
            assert np.array_equal(out["mask_2d_visiable"].numpy(), _unpack(z["mask_2d_visiable"], (6, 64, 64, 1)))
            assert np.array_equal(out["mask_2d"].numpy(), _unpack(z["mask_2d"], (1, 64, 64, 1)))
            assert np.array_equal(out["alpha_mv"].numpy() > 0, _unpack(z["mv.alpha_visiable"], (6, 48, 48, 1)))
    assert [c[0] for c in calls] == z["field.n_visible"].tolist()
    assert np.array_equal(np.stack([c[1] for c in calls]), z["field.query_sum"])
    on = (img.sum(-1, keepdim=True) > 0)
    img9 = torch.cat([img, 0.5 * img + 0.1 * on, (1.0 - img) * on], dim=-1)                 # image_attrs.shape[-1] == 9 (:711-719)
    for name in ("reproject", "kdtree_order_mean"):
        out = ob.infer(v, f, uv, fuv, c2ws, intr, img9, 48, 48, 64, 64, **variants[name])
        assert np.array_equal(out["color_2d"].numpy(), z[f"{name}.pbr9.color_2d"]), name


def test_attention_processor_matches_reference():
    """oracle joint_attention against NativeFluxAttnProcessor2_0.__call__ (attention_processor.py:24-110)."""
    z = np.load(os.path.join(G, "ref_attention.npz"))
    cfg = fd.FluxConfig(num_layers=1, num_single_layers=1, num_attention_heads=2)
    P = fd.init_params(cfg, 33, norm_weight_std=0.1)
    cos, sin = fd.rope_table(torch.from_numpy(z["ids"]), cfg)
    x, ctx = torch.from_numpy(z["x"]), torch.from_numpy(z["ctx"])
    ox, oc = fd.joint_attention(P, "transformer_blocks.0.attn.", cfg, x, ctx, cos, sin)
    assert np.array_equal(ox.numpy(), z["out_x"]) and np.array_equal(oc.numpy(), z["out_ctx"])
    Ps = {k.replace("transformer_blocks.0.", "single_transformer_blocks.0."): v for k, v in P.items() if k.startswith("transformer_blocks.0.attn.")}
    o = fd.joint_attention(Ps, "single_transformer_blocks.0.attn.", cfg, torch.cat([ctx, x], 1), None, cos, sin)
    assert np.array_equal(o.numpy(), z["out_single"])


def _flux_setup():
    from oracle import vae as ov
    cfg = fd.FluxConfig(num_layers=1, num_single_layers=1, num_attention_heads=2)
    P = {k: v.to(torch.bfloat16) for k, v in fd.init_params(cfg, 21, norm_weight_std=0.1).items()}
    vcfg = ov.VaeConfig.tiny()
    VP = {k: v.to(torch.bfloat16) for k, v in ov.init_params(vcfg, 5).items()}
    return cfg, P, vcfg, VP


def test_sampler_matches_reference_call():
    """oracle/flux_sampler.py against PBRFluxPipeline.__call__ itself (flux_piplines/{texturing,delight}/pipeline.py:404-700):
    schedule, bf16 timesteps, id offsets, token order, condition re-imposition, Euler update -- bit for bit in bf16."""
    from PIL import Image
    z = np.load(os.path.join(G, "ref_flux_call.npz"))
    cfg, P, vcfg, VP = _flux_setup()
    ctrl, dual = Image.fromarray(z["control_image"]), Image.fromarray(z["dual_image"])
    for task, d in (("texturing", dual), ("delight", None)):
        assert abs(fs.calculate_shift(64) - float(z[f"{task}.mu"])) < 1e-12
        assert np.array_equal(fs.flow_match_sigmas(3, 64), z[f"{task}.sigmas"])
        ids = fs.build_ids(16, 16, (16, 16), (8, 8) if d is not None else None)
        assert np.array_equal(ids.numpy(), z[f"{task}.img_ids"])
        t = torch.from_numpy(fs.flow_match_sigmas(3, 64)[:3] * np.float32(1000.0)).to(torch.bfloat16) / 1000
        assert np.array_equal(t.float().numpy()[:, None], z[f"{task}.timesteps"])
        lat = fs.pipeline_call(P, cfg, VP, vcfg, ctrl, d, 128, 128, 3, torch.Generator().manual_seed(63), S_txt=128)
        assert np.array_equal(lat.float().numpy(), z[f"{task}.latents"]), task
        img = fs.pipeline_call(P, cfg, VP, vcfg, ctrl, d, 128, 128, 3, torch.Generator().manual_seed(63), S_txt=128, output_type="pil")
        assert np.array_equal(img[0], z[f"{task}.image"]), task


def test_export_condition_matches_reference():
    """oracle export_condition against VideoExporter.export_condition itself (video/export_nvdiffrast_video.py:900-999)."""
    from unitex_b200.export import vertex_normals
    z = np.load(os.path.join(G, "ref_glue.npz"))
    v, f, _, _ = two_spheres(10, 20)
    vn = vertex_normals(torch.from_numpy(v), torch.from_numpy(f).long()).numpy()
    for name, kw in (("six", dict(n_views=6, n_rows=2, n_cols=3)), ("four", dict(n_views=4, n_rows=2, n_cols=2)),
                     ("four_persp", dict(n_views=4, n_rows=2, n_cols=2, perspective=True)),
                     ("orbit8", dict(n_views=8, n_rows=2, n_cols=4, orbit=True))):
        out = ob.export_condition(v, f, vn, geometry_scale=0.95, H=64, W=64, scale=1.0, **kw)
        for k in ("alpha", "ccm", "normal"):
            assert np.array_equal(out[k], z[f"cond.{name}.{k}"]), (name, k)
        assert np.array_equal(out["c2ws"].numpy(), z[f"cond.{name}.c2ws"])
        assert np.array_equal(out["intrinsics"].numpy(), z[f"cond.{name}.intrinsics"])


def test_infer_mv_glue_matches_reference(tmp_path):
    """This repo's drop-in `infer_mv` (pipeline.py) against the reference's own `infer_mv` (pipeline.py:231-291) driven with the
    same fake FLUX object: control strip (blend + view permutation + flip), both calls' kwargs, adapter switching, mv_rgb grid."""
    from PIL import Image
    import types
    import pipeline as drop_in
    from tests.glue_fakes import FakeFlux, glue_inputs, sha
    z = np.load(os.path.join(G, "ref_glue.npz"))
    normal, ccm, ref = glue_inputs()
    d = str(tmp_path)
    Image.fromarray(normal).save(os.path.join(d, "mv_normal.png"))
    Image.fromarray(ccm).save(os.path.join(d, "mv_ccm.png"))
    Image.fromarray(ref).save(os.path.join(d, "processed_image.png"))
    fake = FakeFlux()
    me = types.SimpleNamespace(pipeline=fake, pipeline_name="texture_plus", adapter_names=["texture", "delight"],
                               weights_for_texture=[1.0, 0.0], weights_for_delight=[0.0, 1.0], generator=None, super_resolutions=False)
    drop_in.CustomRGBTextureFullPipeline.infer_mv(me, d, os.path.join(d, "processed_image.png"), os.path.join(d, "mv_normal.png"),
                                                  os.path.join(d, "mv_ccm.png"))
    assert sha(np.array(fake.calls[0]["control_image"])) == str(z["mv.strip_sha"])
    assert sha(np.array(fake.calls[0]["dual_image"])) == str(z["mv.dual_sha"])
    assert sha(np.array(fake.calls[1]["control_image"])) == str(z["mv.second_control_sha"])
    assert fake.calls[1].get("dual_image") is None and not bool(z["mv.second_has_dual"])
    assert repr(sorted((k, v) for k, v in fake.calls[0].items() if k not in ("control_image", "dual_image"))) == str(z["mv.kwargs"])
    assert repr(fake.adapters) == str(z["mv.adapters"])
    assert sha(np.array(Image.open(os.path.join(d, "mv_rgb.png")))) == str(z["mv.rgb_sha"])
    assert sha(np.array(Image.open(os.path.join(d, "mv_rgb_w_light.png")))) == str(z["mv.w_light_sha"])


def test_infer_filt_gradient_points_matches_reference():
    """filt_gradient_points=True, the default of the reference's infer signature (renderer_inverse.py:657): mv_to_pcd's gradient
    filter with its x-only 31-pixel erosion (:188-214), then the reproject bake."""
    from unitex_b200.bake import generate_box_views_c2ws, generate_intrinsics
    z = np.load(os.path.join(G, "ref_bake.npz"))
    v, f, uv, fuv = two_spheres(10, 20)
    c2ws = generate_box_views_c2ws(2.8)[[0, 1, 4, 2, 3, 5]]
    intr = generate_intrinsics(1.0, 1.0, fov=False)
    img = torch.from_numpy(z["filt.image"].astype(np.float32))
    out = ob.infer(v, f, uv, fuv, c2ws, intr, img, 128, 128, 64, 64, method="reproject", filt_gradient_points=True, grad_norm_threhold=0.2)
    assert np.array_equal(out["alpha_mv"].numpy() > 0, _unpack(z["filt.alpha_visiable"], (6, 128, 128, 1)))
    assert np.array_equal(out["mask_2d_visiable"].numpy(), _unpack(z["filt.mask_2d_visiable"], (6, 64, 64, 1)))
    assert np.array_equal(out["color_2d"].numpy(), z["filt.reproject.color_2d"])


def test_infer_perspective_matches_reference():
    """perspective=True, the other default of the reference's infer signature (renderer_inverse.py:639): pinhole projection
    (camera/conversion.py:11-18), rays from the camera position (:279-284)."""
    from unitex_b200.bake import generate_box_views_c2ws, generate_intrinsics, intr_to_proj
    z = np.load(os.path.join(G, "ref_bake.npz"))
    v, f, uv, fuv = two_spheres(10, 20)
    c2ws = generate_box_views_c2ws(2.8)[[0, 1, 4, 2, 3, 5]]
    intr = generate_intrinsics(49.1, 49.1, fov=True, degree=True)
    assert np.array_equal(intr.numpy(), z["persp.intrinsics"])
    assert np.array_equal(intr_to_proj(intr, perspective=True).numpy(), z["persp.proj"])
    assert np.array_equal(ob.intr_to_proj_persp(intr).numpy(), z["persp.proj"])
    img = torch.from_numpy(z["persp.image"].astype(np.float32))
    for name, kw in (("reproject", dict(method="reproject")), ("kdtree_order_mean", dict(method="kdtree", kdtree_method="order_mean", k_vis=9, k_invis=32))):
        out = ob.infer(v, f, uv, fuv, c2ws, intr, img, 48, 48, 64, 64, perspective=True, **kw)
        assert np.array_equal(out["color_2d"].numpy(), z[f"persp.{name}.color_2d"]), name
    assert np.array_equal(out["mask_2d_visiable"].numpy(), _unpack(z["persp.mask_2d_visiable"], (6, 64, 64, 1)))
    assert out["mask_2d_visiable"].sum() > 1000


def test_lora_target_list_matches_reference_source():
    """Only where the reference tree is present (the build container): the LoRA target modules the oracle and the product merge
    are the ones the reference trains (flux_piplines/texturing/trainer.py:282-304)."""
    import ast
    import re
    import pytest
    path = "/root/reference/flux_piplines/texturing/trainer.py"
    if not os.path.exists(path):
        pytest.skip("reference tree absent")
    src = open(path).read()
    targets = ast.literal_eval(re.search(r"target_modules = (\[\s*#[^\n]*\n.*?\])", src, re.S).group(1))
    saved = ast.literal_eval(re.search(r"modules_to_save = (\[.*?\])", src, re.S).group(1))
    assert set(targets) == set(fs.LORA_TARGETS_DOUBLE)
    assert set(fs.LORA_TARGETS_SINGLE) <= set(targets)                  # single blocks only own attn.to_q/k/v of that list
    assert saved[0] == "x_embedder"                                      # the one modules_to_save entry that has parameters


def test_preprocess_reference_image_matches_reference(tmp_path):
    """The matte-driven crop / scale / paste of the reference image (pipeline.py:182-196, image/process_image.py:31-81) against the
    reference's own functions run with a stand-in matte source."""
    import types
    from PIL import Image
    import pipeline as drop_in
    from tests.glue_fakes import reference_rgba, sha
    z = np.load(os.path.join(G, "ref_glue.npz"))
    src = str(tmp_path / "ref_rgba.png")
    Image.fromarray(reference_rgba(), mode="RGBA").save(src)
    drop_in.CustomRGBTextureFullPipeline.preprocess_reference_image(types.SimpleNamespace(), str(tmp_path), src)
    a_full, a_small = np.array(Image.open(tmp_path / "rembg_image.png")), np.array(Image.open(tmp_path / "processed_image.png"))
    assert a_full.shape == tuple(z["pre.rembg_shape"]) and a_small.shape == tuple(z["pre.processed_shape"])
    assert np.array_equal(a_full[::64, ::64], z["pre.rembg_probe"]) and np.array_equal(a_small[::32, ::32], z["pre.processed_probe"])
    assert sha(a_full) == str(z["pre.rembg_sha"]) and sha(a_small) == str(z["pre.processed_sha"])


def test_preprocess_reference_image_matte_hook_matches_reference(tmp_path):
    """An RGB input without alpha + the constructor's `rembg_session` hook (what the reference builds from RMBG-2.0 / rembg,
    pipeline.py:146-149): the same call the fixture was made with -- `rembg_session(image)` hands back the matte -- so the
    outputs equal the reference's own `preprocess_reference_image` run byte for byte.  Without a session: a warning, not silence."""
    import types
    import warnings
    from PIL import Image
    import pipeline as drop_in
    from tests.glue_fakes import reference_rgba, sha
    z = np.load(os.path.join(G, "ref_glue.npz"))
    rgba = Image.fromarray(reference_rgba(), mode="RGBA")
    src = str(tmp_path / "ref_rgb.png")
    rgba.convert("RGB").save(src)
    me = types.SimpleNamespace(rembg_session=lambda image: rgba)
    drop_in.CustomRGBTextureFullPipeline.preprocess_reference_image(me, str(tmp_path), src)
    a_full, a_small = np.array(Image.open(tmp_path / "rembg_image.png")), np.array(Image.open(tmp_path / "processed_image.png"))
    assert sha(a_full) == str(z["pre.rembg_sha"]) and sha(a_small) == str(z["pre.processed_sha"])
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        drop_in.CustomRGBTextureFullPipeline.preprocess_reference_image(types.SimpleNamespace(rembg_session=None), str(tmp_path), src)
    assert any("un-matted" in str(x.message) for x in w)
