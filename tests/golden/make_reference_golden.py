"""Golden fixtures produced by the REFERENCE'S OWN Python, run in the build container through `ref_harness.py`
(absent third-party packages stubbed; their arithmetic supplied by the oracle's restatements -- see the harness header
for exactly what is real and what is substituted).       python tests/golden/make_reference_golden.py

Writes tests/golden/ref_*.npz.  tests/test_reference_golden_cpu.py pins the oracle against them (bit-exact where the code
path is torch on both sides); tests/test_gpu_reference_golden.py compares the CUDA path with them on the GPU box.
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
import ref_harness as rh                                      # noqa: E402
from oracle import flux_dit as fd                             # noqa: E402
from oracle import vae as ov                                  # noqa: E402
from tests.bake_meshes import two_spheres                     # noqa: E402
from tests.glue_fakes import FakeFlux, glue_inputs, reference_rgba, sha as _sha   # noqa: E402


def bake():
    """NVDiffRendererInverse.infer (renderer_inverse.py:635-726) -- the reference class, its PBRMesh / PointCloud / knn /
    pull_push / lens_blur_torch / camera code -- on the two-sphere case of bake_two_spheres.npz, every variant the path has."""
    ri = rh.texturetools()
    from texturetools.camera.conversion import c2w_to_w2c, intr_to_proj
    from texturetools.camera.generator import generate_box_views_c2ws, generate_intrinsics
    from texturetools.image.lens_blur import lens_blur_torch
    from texturetools.mesh.structure_v2 import PBRMesh
    from texturetools.pcd.knn import knn
    from texturetools.texture.stitching.mip import pull_push

    v, f, uv, fuv = two_spheres(10, 20)
    z = np.load(os.path.join(HERE, "bake_two_spheres.npz"))
    img = torch.from_numpy(z["image"])
    c2ws_all = generate_box_views_c2ws(radius=2.8)                       # generator.py:153-185
    c2ws = c2ws_all[[0, 1, 4, 2, 3, 5]]
    intr = generate_intrinsics(1.0, 1.0, fov=False)                      # generator.py:93-114
    out = {"c2ws_all": c2ws_all.numpy(), "intrinsics": intr.numpy(),
           "proj": intr_to_proj(intr, perspective=False).numpy(), "w2c": c2w_to_w2c(c2ws).numpy()}
    mesh = PBRMesh(torch.from_numpy(v), torch.from_numpy(f).long(), torch.from_numpy(uv), torch.from_numpy(fuv).long())
    out["face_normals"] = mesh.normals.numpy()
    out["vertex_normals"] = mesh.vertex_normals.numpy()                  # structure_v2.py:63-71
    r = ri.NVDiffRendererInverse(device="cpu", pbr_mesh=mesh)
    common = dict(H=48, W=48, H2D=64, W2D=64, perspective=False, ray_normal_angle_threhold=100.0, filt_gradient_points=False)

    def field(vertices_visiable, colors_visiable, vertices_invisiable):  # register_query_field contract :93-103
        field.calls.append((vertices_visiable.clone(), colors_visiable.clone(), vertices_invisiable.clone()))
        return 0.25 + 0.5 * torch.sigmoid(vertices_invisiable * 3.0)
    field.calls = []

    variants = {
        "reproject": dict(method="reproject"),
        "kdtree_order_mean": dict(method="kdtree", kdtree_method="order_mean", kdtree_n_neighbors_visiable=9, kdtree_n_neighbors_invisiable=32),
        "kdtree_mean": dict(method="kdtree", kdtree_method="mean", kdtree_n_neighbors=32),
        "kdtree_mvpaint": dict(method="kdtree", kdtree_method="mvpaint", kdtree_n_neighbors=8),
        "reproject_gaussian": dict(method="reproject", reproject_method="gaussian"),
        "reproject_inpaint": dict(method="reproject", reproject_inpainting=True),
        "kdtree_inpaint": dict(method="kdtree", kdtree_method="order_mean", kdtree_n_neighbors_visiable=9, kdtree_inpainting=True),
    }
    r.register_query_field(field)
    for name, kw in variants.items():
        _, vis, m2d, col = r.infer(None, c2ws, intr, img, **common, **kw)
        out[f"{name}.color_2d"] = col.numpy()
        if name == "reproject":
            out["mask_2d_visiable"] = np.packbits(vis.numpy())
            out["mask_2d"] = np.packbits(m2d.numpy())
    img9 = torch.cat([img, 0.5 * img + 0.1 * (img.sum(-1, keepdim=True) > 0), (1.0 - img) * (img.sum(-1, keepdim=True) > 0)], dim=-1)
    for name, kw in (("reproject", dict(method="reproject")), ("kdtree_order_mean", variants["kdtree_order_mean"])):
        r.register_query_field(None)
        _, _, _, col = r.infer(None, c2ws, intr, img9, **common, **kw)              # image_attrs.shape[-1] == 9 (:711-719)
        out[f"{name}.pbr9.color_2d"] = col.numpy()
    out["field.n_visible"] = np.array([c[0].shape[0] for c in field.calls])
    out["field.query_sum"] = np.stack([c[2].double().sum(0).numpy() for c in field.calls])
    # filt_gradient_points=True -- the DEFAULT of infer's signature (:657); needs views wide enough to survive the 31-pixel erosion
    big = dict(common, H=128, W=128, filt_gradient_points=True, grad_norm_threhold=0.2)
    from oracle import bake as ob
    from tests.bake_meshes import analytic_color
    mats = torch.matmul(intr_to_proj(intr, perspective=False), c2w_to_w2c(c2ws))
    rast128 = ob.rasterize(torch.matmul(torch.cat([torch.from_numpy(v), torch.ones(len(v), 1)], -1), mats.permute(0, 2, 1)).numpy(), f, 128, 128)
    img128 = torch.from_numpy(analytic_color(ob.interpolate(v, rast128, f)) * (rast128[..., 3:4] > 0)).float()
    out["filt.image"] = img128.numpy().astype(np.float16)
    img128 = torch.from_numpy(out["filt.image"].astype(np.float32))
    mvf = r.mv_to_pcd(c2ws, intr, (128, 128), image_attrs=img128, perspective=False, grad_norm_threhold=0.2,
                      ray_normal_angle_threhold=100.0, filt_gradient_points=True)
    out["filt.alpha_visiable"] = np.packbits(mvf["alpha_visiable"].numpy() > 0)
    r.register_query_field(None)
    _, visf, _, colf = r.infer(None, c2ws, intr, img128, **big, method="reproject")
    out["filt.reproject.color_2d"] = colf.numpy()
    out["filt.mask_2d_visiable"] = np.packbits(visf.numpy())
    # perspective=True -- the other DEFAULT of infer's signature (:639): pinhole views, rays from the camera position (:279-284)
    intr_p = generate_intrinsics(49.1, 49.1, fov=True, degree=True)
    out["persp.intrinsics"] = intr_p.numpy()
    out["persp.proj"] = intr_to_proj(intr_p, perspective=True).numpy()
    mats_p = torch.matmul(intr_to_proj(intr_p, perspective=True), c2w_to_w2c(c2ws))
    rast_p = ob.rasterize(torch.matmul(torch.cat([torch.from_numpy(v), torch.ones(len(v), 1)], -1), mats_p.permute(0, 2, 1)).numpy(), f, 48, 48)
    img_p = torch.from_numpy((analytic_color(ob.interpolate(v, rast_p, f)) * (rast_p[..., 3:4] > 0)).astype(np.float16).astype(np.float32))
    out["persp.image"] = img_p.numpy().astype(np.float16)
    for name in ("reproject", "kdtree_order_mean"):
        _, visp, m2p, colp = r.infer(None, c2ws, intr_p, img_p, **dict(common, perspective=True), **variants[name])
        out[f"persp.{name}.color_2d"] = colp.numpy()
    out["persp.mask_2d_visiable"] = np.packbits(visp.numpy())
    # mv_to_pcd as the shipped path calls it (filt_gradient_points=False, pipeline.py:343-347)
    mv0 = r.mv_to_pcd(c2ws, intr, (48, 48), image_attrs=img, perspective=False, filt_gradient_points=False)
    out["mv.alpha_visiable"] = np.packbits(mv0["alpha_visiable"].numpy() > 0)

    # function-level vectors (pure torch in the reference, no stand-ins involved)
    g = torch.Generator().manual_seed(1)
    x = torch.rand(1, 3, 40, 56, generator=g)
    m = torch.rand(1, 1, 40, 56, generator=g) > 0.6
    out["fn.x"], out["fn.mask"] = x.numpy(), m.numpy()
    out["fn.lens_blur"] = lens_blur_torch(x).numpy()                     # image/lens_blur.py:260-280
    kd, km = pull_push(x * m, m)                                         # texture/stitching/mip.py:51-96
    out["fn.pull_push"], out["fn.pull_push_mask"] = kd.numpy(), km.numpy()
    mb = torch.rand(2, 24, 24, 1, generator=g) > 0.5
    out["fn.bmask_in"] = mb.numpy()
    out["fn.bmask"] = r.get_boundary_mask(mb, kernel_size=3).numpy()     # renderer_inverse.py:435-444
    src, dst = torch.rand(300, 3, generator=g), torch.rand(40, 3, generator=g)
    score, index = knn(src, dst, k=4)                                    # pcd/knn/__init__.py:104-114 (wrapper real, tree [ext])
    out["fn.knn_src"], out["fn.knn_dst"], out["fn.knn_index"] = src.numpy(), dst.numpy(), index.numpy().astype(np.int32)
    # the ray-tracer plug-in as the reference exposes it: RayTracing dispatcher -> APRMISRayTracing wrapper (raytracing/__init__.py:12-80,
    # rt_aprmis/__init__.py:10-86; the Slang launch inside is the oracle's restatement), batch-shaped rays, misses included
    from texturetools.raytracing import RayTracing
    rt = RayTracing(torch.from_numpy(v), torch.from_numpy(f).long())
    ro = torch.rand(3, 50, 3, generator=g) * 2.4 - 1.2
    rdir = torch.nn.functional.normalize(torch.rand(3, 50, 3, generator=g) - 0.5 - 0.4 * ro, dim=-1)
    hit, front, tri_idx, loc, ruv = rt.intersects_closest(ro, rdir)
    assert front is None and hit.dtype == torch.bool and tri_idx.dtype == torch.int64
    out["rt.rays_o"], out["rt.rays_d"] = ro.numpy(), rdir.numpy()
    out["rt.hit"], out["rt.tri_idx"], out["rt.loc"], out["rt.uv"] = hit.numpy(), tri_idx.numpy().astype(np.int32), loc.numpy(), ruv.numpy()
    np.savez_compressed(os.path.join(HERE, "ref_bake.npz"), **{k: (a.astype(np.float32) if a.dtype == np.float64 and not k.startswith("field") else a)
                                                               for k, a in out.items()})
    return out


class _Dist:
    """DiagonalGaussianDistribution [ext]: sample = mean + std * randn_tensor(mean.shape, generator, dtype=mean.dtype)."""
    def __init__(self, mean, logvar):
        self.mean, self.logvar = mean, logvar

    def sample(self, generator=None):
        from diffusers.utils.torch_utils import randn_tensor
        n = randn_tensor(self.mean.shape, generator=generator, device=self.mean.device, dtype=self.mean.dtype)
        return ov.sample(self.mean, self.logvar, n)


def flux_inputs():
    from PIL import Image
    rng = np.random.default_rng(0)
    yy, xx = np.mgrid[0:128, 0:128]
    base = np.stack([127 + 100 * np.sin(xx / 9.0), 127 + 100 * np.cos(yy / 7.0), 127 + 90 * np.sin((xx + yy) / 11.0)], -1)
    ctrl = np.clip(base + rng.normal(0, 6, base.shape), 0, 255).astype(np.uint8)
    dual = np.clip(base[::2, ::2][:, ::-1] + rng.normal(0, 6, (64, 64, 3)), 0, 255).astype(np.uint8)
    return Image.fromarray(ctrl), Image.fromarray(dual)


FLUX_CFG = dict(num_layers=1, num_single_layers=1, num_attention_heads=2)     # text widths stay 4096 / 768: the reference hard-codes them (:538-543)
FLUX_SEED, VAE_SEED, S_TXT, STEPS = 21, 5, 128, 3


def flux_weights():
    cfg = fd.FluxConfig(**FLUX_CFG)
    P = {k: v.to(torch.bfloat16) for k, v in fd.init_params(cfg, FLUX_SEED, norm_weight_std=0.1).items()}
    vcfg = ov.VaeConfig.tiny()
    VP = {k: v.to(torch.bfloat16) for k, v in ov.init_params(vcfg, VAE_SEED).items()}
    return cfg, P, vcfg, VP


def flux():
    """PBRFluxPipeline.__call__ (flux_piplines/{texturing,delight}/pipeline.py:404-700) -- the reference's class, its __init__,
    prepare_latents_and_image_ids, pack / unpack / ids, calculate_shift, retrieve_timesteps and the condition-token loop --
    around bf16 eager stand-ins for the diffusers modules (tiny FLUX: 1+1 blocks, 2 heads; tiny VAE)."""
    from diffusers.schedulers.scheduling_flow_match_euler_discrete import FlowMatchEulerDiscreteScheduler
    cfg, P, vcfg, VP = flux_weights()
    ctrl, dual = flux_inputs()

    class Vae:
        config = rh._Cfg(latent_channels=vcfg.latent_channels, block_out_channels=vcfg.block_out_channels,
                         scaling_factor=vcfg.scaling_factor, shift_factor=vcfg.shift_factor)
        dtype, device = torch.bfloat16, torch.device("cpu")

        def encode(self, image):
            return types.SimpleNamespace(latent_dist=_Dist(*ov.encode_moments(VP, vcfg, image)))

        def decode(self, z, return_dict=False):
            return (ov.decode(VP, vcfg, z),)

    class Transformer:
        config = rh._Cfg(guidance_embeds=True)
        calls = []

        def __call__(self, hidden_states, timestep, guidance, pooled_projections, encoder_hidden_states, txt_ids, img_ids,
                     joint_attention_kwargs=None, return_dict=False):
            self.calls.append(dict(timestep=timestep.clone(), hidden=hidden_states.clone(), img_ids=img_ids.clone(), txt_ids=txt_ids.clone()))
            return (fd.flux_forward(P, cfg, hidden_states, timestep, guidance, pooled_projections, encoder_hidden_states, txt_ids, img_ids),)

    out = {"control_image": np.asarray(ctrl), "dual_image": np.asarray(dual)}
    for task, use_dual in (("texturing", True), ("delight", False)):
        mod = rh.flux(task)
        tr = Transformer()
        tr.calls = []
        pipe = mod.PBRFluxPipeline(FlowMatchEulerDiscreteScheduler(), Vae(), None, None, None, None, tr)
        kw = dict(prompt="[MVFLUX]", control_image=ctrl, dual_image=dual if use_dual else None, height=128, width=128, n_rows=1, n_cols=1,
                  num_inference_steps=STEPS, guidance_scale=3.5, max_sequence_length=S_TXT)
        lat = pipe(**kw, generator=torch.Generator().manual_seed(63), output_type="latent").images
        n_tok = tr.calls[0]["hidden"].shape[1]
        out[f"{task}.latents"] = lat.float().numpy()
        out[f"{task}.timesteps"] = torch.stack([c["timestep"] for c in tr.calls[:STEPS]]).float().numpy()      # bf16(t)/1000 as fed (:643,:648)
        out[f"{task}.img_ids"] = tr.calls[0]["img_ids"].float().numpy()
        out[f"{task}.tokens_step0"] = tr.calls[0]["hidden"].float().numpy()                                       # [noise | control | dual] packed
        out[f"{task}.sigmas"] = pipe.scheduler.sigmas.numpy()
        assert lat.shape == (1, 64, 64) and n_tok == (64 + 64 + (16 if use_dual else 0))
        img = pipe(**kw, generator=torch.Generator().manual_seed(63), output_type="pil").images[0]
        out[f"{task}.image"] = np.asarray(img)
        out[f"{task}.mu"] = np.float64(mod.calculate_shift(64, 256, 4096, 0.5, 1.15))
    np.savez_compressed(os.path.join(HERE, "ref_flux_call.npz"), **out)
    return out


def attention():
    """NativeFluxAttnProcessor2_0.__call__ (attention_processor.py:24-110) -- real -- on nn.Linear projections; RMSNorm and
    apply_rotary_emb are diffusers [ext] and come from the oracle."""
    import diffusers.models.embeddings as emb
    emb.apply_rotary_emb = lambda x, freqs: fd.apply_rope(x, freqs[0], freqs[1])
    ap = rh.flux_attention("texturing")
    cfg = fd.FluxConfig(**FLUX_CFG)
    P = fd.init_params(cfg, 33, norm_weight_std=0.1)
    pre = "transformer_blocks.0.attn."

    class Lin:
        def __init__(self, n):
            self.n = n

        def __call__(self, x):
            return torch.nn.functional.linear(x, P[pre + self.n + ".weight"], P[pre + self.n + ".bias"])

    class Norm:
        def __init__(self, n):
            self.n = n

        def __call__(self, x):
            return fd.rms_norm(x, P[pre + self.n + ".weight"])

    attn = types.SimpleNamespace(heads=2, to_q=Lin("to_q"), to_k=Lin("to_k"), to_v=Lin("to_v"), add_q_proj=Lin("add_q_proj"),
                                 add_k_proj=Lin("add_k_proj"), add_v_proj=Lin("add_v_proj"), norm_q=Norm("norm_q"), norm_k=Norm("norm_k"),
                                 norm_added_q=Norm("norm_added_q"), norm_added_k=Norm("norm_added_k"),
                                 to_out=[Lin("to_out.0"), lambda x: x], to_add_out=Lin("to_add_out"))
    g = torch.Generator().manual_seed(2)
    x, ctx = torch.randn(1, 96, 256, generator=g), torch.randn(1, 32, 256, generator=g)
    ids = torch.cat([torch.zeros(32, 3), torch.stack([torch.zeros(96), torch.arange(96) // 12, torch.arange(96) % 12], -1)])
    cos, sin = fd.rope_table(ids, cfg)
    hx, hc = ap.NativeFluxAttnProcessor2_0()(attn, x, encoder_hidden_states=ctx, image_rotary_emb=(cos, sin))
    single = ap.NativeFluxAttnProcessor2_0()(attn, torch.cat([ctx, x], 1), image_rotary_emb=(cos, sin))
    np.savez_compressed(os.path.join(HERE, "ref_attention.npz"), x=x.numpy(), ctx=ctx.numpy(), ids=ids.numpy(), out_x=hx.numpy(),
                        out_ctx=hc.numpy(), out_single=single.numpy())


def glue():
    """export_condition (video/export_nvdiffrast_video.py:900-999 over renderer_base.simple_rendering :101-200 and
    mesh/structure.py scale_to_bbox / apply_transform), and the top-level pipeline.py glue: infer_mv (:231-291, grid
    re-ordering a9 + the two calls) and reproject_and_query_field (:312-360, view slicing + infer kwargs)."""
    import tempfile
    from PIL import Image
    from unitex_b200.export import vertex_normals
    v, f, uv, fuv = two_spheres(10, 20)
    vn = vertex_normals(torch.from_numpy(v), torch.from_numpy(f).long()).numpy()
    ve = rh.video_exporter(v, f, vn)
    out = {}
    for name, kw in (("six", dict(n_views=6, n_rows=2, n_cols=3)), ("four", dict(n_views=4, n_rows=2, n_cols=2)),
                     ("four_persp", dict(n_views=4, n_rows=2, n_cols=2, perspective=True)),
                     ("orbit8", dict(n_views=8, n_rows=2, n_cols=4, orbit=True))):
        kw = dict(dict(perspective=False, orbit=False), **kw)
        r = ve.export_condition("mesh.obj", geometry_scale=0.95, H=64, W=64, fov_deg=49.1, scale=1.0,
                                background="grey", return_info=False, return_image=True, return_mesh=False, return_camera=True, **kw)
        for k in ("alpha", "ccm", "normal"):
            out[f"cond.{name}.{k}"] = np.asarray(r[k])
        out[f"cond.{name}.c2ws"], out[f"cond.{name}.intrinsics"] = r["c2ws"].numpy(), r["intrinsics"].numpy()
        assert r["perspective"] is kw["perspective"]

    mod = rh.top_level_pipeline()
    normal, ccm, ref = glue_inputs()
    with tempfile.TemporaryDirectory() as d:
        Image.fromarray(normal).save(os.path.join(d, "mv_normal.png"))
        Image.fromarray(ccm).save(os.path.join(d, "mv_ccm.png"))
        Image.fromarray(ref).save(os.path.join(d, "processed_image.png"))
        fake = FakeFlux()
        me = types.SimpleNamespace(pipeline=fake, pipeline_name="texture_plus", adapter_names=["texture", "delight"],
                                   weights_for_texture=[1.0, 0.0], weights_for_delight=[0.0, 1.0], generator=None, super_resolutions=False)
        mod.CustomRGBTextureFullPipeline.infer_mv(me, d, os.path.join(d, "processed_image.png"), os.path.join(d, "mv_normal.png"),
                                                  os.path.join(d, "mv_ccm.png"))
        out["mv.strip_sha"] = np.array(_sha(np.array(fake.calls[0]["control_image"])))
        out["mv.strip_probe"] = np.array(fake.calls[0]["control_image"])[::64, ::64].copy()
        out["mv.dual_sha"] = np.array(_sha(np.array(fake.calls[0]["dual_image"])))
        out["mv.second_control_sha"] = np.array(_sha(np.array(fake.calls[1]["control_image"])))
        out["mv.second_has_dual"] = np.array("dual_image" in fake.calls[1] and fake.calls[1]["dual_image"] is not None)
        out["mv.kwargs"] = np.array(repr(sorted((k, v) for k, v in fake.calls[0].items() if k not in ("control_image", "dual_image"))))
        out["mv.adapters"] = np.array(repr(fake.adapters))
        out["mv.rgb_sha"] = np.array(_sha(np.array(Image.open(os.path.join(d, "mv_rgb.png")))))
        out["mv.rgb_probe"] = np.array(Image.open(os.path.join(d, "mv_rgb.png")))[::64, ::64].copy()
        out["mv.w_light_sha"] = np.array(_sha(np.array(Image.open(os.path.join(d, "mv_rgb_w_light.png")))))

        # preprocess_reference_image (:182-196) over image/process_image.py::preprocess (:31-81), the matte supplied by a stand-in for
        # the background-removal network [ext]: it hands back the test image's own alpha
        rgba = Image.fromarray(reference_rgba(), mode="RGBA")
        rgba.save(os.path.join(d, "ref_rgba.png"))
        me3 = types.SimpleNamespace(rembg_session=lambda image: rgba)
        mod.CustomRGBTextureFullPipeline.preprocess_reference_image(me3, d, os.path.join(d, "ref_rgba.png"))
        a_full, a_small = np.array(Image.open(os.path.join(d, "rembg_image.png"))), np.array(Image.open(os.path.join(d, "processed_image.png")))
        out["pre.rembg_sha"], out["pre.rembg_shape"], out["pre.rembg_probe"] = np.array(_sha(a_full)), np.array(a_full.shape), a_full[::64, ::64].copy()
        out["pre.processed_sha"], out["pre.processed_shape"], out["pre.processed_probe"] = np.array(_sha(a_small)), np.array(a_small.shape), a_small[::32, ::32].copy()

        # reproject_and_query_field: what the bake entry point is handed
        rec = {}

        class FakeInverse:
            def update_from_file(self, path):
                rec["mesh"] = os.path.basename(path)

            def infer(self, blank, **kw):
                rec.update(kw)
                t = types.SimpleNamespace(export=lambda path: open(path, "wb").close())
                n = kw["c2ws"].shape[0]
                return t, torch.zeros(n, 8, 8, 1, dtype=torch.bool), torch.ones(1, 8, 8, 1, dtype=torch.bool), torch.full((1, 8, 8, 3), 0.25)

            def clear(self):
                rec["cleared"] = True
        cam = {"c2ws": torch.from_numpy(out["cond.six.c2ws"]), "intrinsics": torch.from_numpy(out["cond.six.intrinsics"]), "perspective": False}
        torch.save(cam, os.path.join(d, "camera_info.pth"))
        me2 = types.SimpleNamespace(inverse_renderer=FakeInverse())
        mod.CustomRGBTextureFullPipeline.reproject_and_query_field(me2, d, os.path.join(d, "processed_mesh.obj"), os.path.join(d, "mv_rgb.png"),
                                                                   os.path.join(d, "camera_info.pth"), method="reproject", inpainting=False)
        out["rq.image_attrs_sha"] = np.array(_sha(rec["image_attrs"].numpy()))
        out["rq.image_attrs_probe"] = rec["image_attrs"].numpy()[:, ::64, ::64].copy()
        out["rq.kwargs"] = np.array(repr(sorted((k, v) for k, v in rec.items() if k not in ("image_attrs", "c2ws", "intrinsics"))))
        out["rq.files"] = np.array(repr(sorted(n for n in os.listdir(d) if n.endswith((".glb", "_mask.png", "_uv.png")))))
    np.savez_compressed(os.path.join(HERE, "ref_glue.npz"), **out)
    return out


if __name__ == "__main__":
    torch.set_num_threads(8)
    bake()
    attention()
    flux()
    glue()
    print(sorted(n for n in os.listdir(HERE) if n.startswith("ref_")))
