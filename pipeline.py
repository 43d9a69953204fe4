"""`CustomRGBTextureFullPipeline`: drop-in for the reference's top-level orchestrator (reference pipeline.py:141-632,
run.py:1-10) over the B200-native hot path.  Same constructor / `__call__(save_dir, input_image_path, input_mesh_path,
clear_cache)` signature, same step sequence ['step_1_1', 'step_2_ablition'], same cache file names.

What is NOT re-implemented (out of scope, CPU third-party): background matting (rembg / RMBG-2.0: a `rembg_session`
hook takes the reference's own object), mesh decimation + chart-based UV unwrapping (open3d / xatlas: a UV-less mesh gets a
per-triangle atlas instead, unitex_b200/uv_atlas.py), the orbit mp4, super-resolution (TSD_SR).
"""
from __future__ import annotations

import os
import shutil
import warnings
from typing import Optional, Tuple

import numpy as np
import torch
from PIL import Image

from flux_piplines.texturing.pipeline import PBRFluxPipeline
from unitex_b200 import bake as ub
from unitex_b200 import export as ux
from unitex_b200.flux import FluxConfig


def build_pipeline(pretrain_models=None, pipeline_name="texture_plus", model="rgb", super_resolutions=False,
                   add_lora_path=None, add_lora_weights=None, speedup_mode=False):
    """reference pipeline.py:81-127.  `pretrain_models='random'` (or a FluxConfig) builds random-init weights + random LoRAs
    so the whole pipeline runs without the HF checkpoints (tests, smoke)."""
    if pretrain_models == "random" or isinstance(pretrain_models, FluxConfig):
        cfg = pretrain_models if isinstance(pretrain_models, FluxConfig) else FluxConfig()
        pipeline = PBRFluxPipeline.from_random(cfg, seed=0, with_vae=True)
        g = torch.Generator().manual_seed(1)
        for name in ("texture", "delight"):
            n = "transformer_blocks.0.attn.to_q"
            D = cfg.inner_dim
            pipeline.load_lora_weights({f"transformer.{n}.lora_A.weight": torch.randn(8, D, generator=g) * 0.02,
                                        f"transformer.{n}.lora_B.weight": torch.randn(D, 8, generator=g) * 0.02}, adapter_name=name)
    else:
        lora_id = f"{pretrain_models}/UniTex/texture_gen/pytorch_lora_weights.safetensors"
        lora_id_delight = f"{pretrain_models}/UniTex/delight/pytorch_lora_weights.safetensors"
        ckpt_id = f"{pretrain_models}/black-forest-labs/FLUX.1-dev"
        pipeline = PBRFluxPipeline.from_pretrained(ckpt_id, text_encoder=None, text_encoder_2=None, torch_dtype=torch.bfloat16)
        pipeline.load_lora_weights(lora_id, adapter_name="texture")
        pipeline.load_lora_weights(lora_id_delight, adapter_name="delight")
    weights_for_texture, weights_for_delight, adapter_names = [1.0, 0.0], [0.0, 1.0], ["texture", "delight"]
    for i, p in enumerate(add_lora_path or []):
        pipeline.load_lora_weights(p, adapter_name=f"add_lora_{i}")
        adapter_names.append(f"add_lora_{i}")
        weights_for_texture.append(add_lora_weights[i])
        weights_for_delight.append(add_lora_weights[i])
    pipeline._num_inference_steps = 28
    return pipeline, weights_for_texture, weights_for_delight, adapter_names


def reference_image_on_canvas(rgb: Image.Image, alpha: Image.Image, H=2048, W=2048, scale=0.8, color="white") -> Image.Image:
    """image/process_image.py:31-81 (`preprocess` with a given alpha): bounding box of the matte, uniform scale so that it spans
    `scale` of the canvas, object pasted centred on a `color` canvas through the matte; the matte itself becomes the alpha channel."""
    m = np.asarray(alpha)
    rows, cols = np.where(m.sum(-1) > 0)[0], np.where(m.sum(-2) > 0)[0]
    x1, y1, x2, y2 = cols.min(), rows.min(), cols.max(), rows.max()
    dy, dx = y2 - y1, x2 - x1
    s = min(H * scale / dy, W * scale / dx)
    Ht, Wt = int(dy * s), int(dx * s)
    ox, oy = int((W - Wt) / 2), int((H - Ht) / 2)
    box_src, box_dst = (int(x1), int(y1), int(x2), int(y2)), (ox, oy, ox + Wt, oy + Ht)
    rgbc = rgb.crop(box_src).resize((Wt, Ht))
    alphac = alpha.crop(box_src).resize((Wt, Ht))
    alphat = Image.new("L", (W, H))
    alphat.paste(alphac, box_dst)
    out = Image.new("RGBA", (W, H), color)
    out.paste(rgbc, box_dst, alphac)
    out.putalpha(alphat)
    return out


class RGBTextureFullPipelineBase:
    step_seq = []

    def __init__(self, pretrain_models=None, pipeline_name="texture_plus", super_resolutions=False, seed=0, speedup_mode=None,
                 add_lora_path=None, add_lora_weights=None, enable_rembg=False, rembg_session=None):
        """`rembg_session`: the matte source the reference builds itself (reference pipeline.py:146-149: rembg when
        `enable_rembg`, else RMBG-2.0 -- both background-removal NETWORKS, out of scope here): any callable
        `image -> RGBA PIL image` (what `preprocess` calls, image/process_image.py:52).  Without one, `enable_rembg=True`
        cannot be honoured and raises; with `enable_rembg=False` an input image that has no alpha channel of its own is
        used un-matted, with a warning (see `preprocess_reference_image`)."""
        if enable_rembg and rembg_session is None:
            raise NotImplementedError("enable_rembg=True needs a matte source: pass rembg_session=<callable image -> RGBA image> "
                                      "(the rembg / RMBG-2.0 networks are out of scope of the B200 hot path)")
        self.rembg_session = rembg_session
        if super_resolutions:
            raise NotImplementedError("TSD_SR super-resolution is out of scope (reference run.py:4 leaves it off)")
        (self.pipeline, self.weights_for_texture, self.weights_for_delight, self.adapter_names) = build_pipeline(
            pretrain_models=pretrain_models, pipeline_name=pipeline_name, speedup_mode=speedup_mode, add_lora_path=add_lora_path,
            add_lora_weights=add_lora_weights, model="rgb")
        self.pipeline_name = pipeline_name
        self.video_exporter = ux.VideoExporter()
        self.inverse_renderer = ub.NVDiffRendererInverse(device="cuda")
        self.seed = seed
        self.generator = torch.Generator().manual_seed(seed)
        self.super_resolutions = False

    # ------------------------------------------------------------------ step pieces
    def preprocess_blank_mesh(self, save_dir, input_mesh_path, min_faces=20_000, max_faces=200_000, scale=0.95):
        """reference :170-179 -> geometry/uv/uv_atlas.py:131-194: the bounding box is centred and its longest side scaled to
        2*scale (float64, like the open3d transform there) -- the bake's cameras assume that frame -- and the mesh is written as
        processed_mesh.obj.  A mesh without UVs gets a per-triangle atlas (the reference's open3d clean-up / decimation /
        compute_uvatlas chain [ext] is out of scope), with a warning."""
        V, F, UV, Ft = ub.load_mesh(input_mesh_path)           # .obj or .glb, like the reference's test cases
        if len(UV) == 0:
            # the reference unwraps with open3d / UVAtlas after cleaning and decimating (uv_atlas.py:149-175: CPU third party, out
            # of scope); here the mesh gets the simplest valid atlas, one slot per triangle (unitex_b200/uv_atlas.py)
            from unitex_b200.uv_atlas import per_triangle_atlas
            warnings.warn("preprocess_blank_mesh: the mesh has no UVs -- the reference would decimate it to <= 200 000 faces and unwrap it "
                          "with open3d's compute_uvatlas; here every triangle gets its own atlas slot (a valid but less economical "
                          "layout, every edge a seam)", stacklevel=2)
            UV, Ft = per_triangle_atlas(len(F), 2048)
        V = np.asarray(V, dtype=np.float64)
        aaa, bbb = V.min(0), V.max(0)
        sss = (bbb - aaa).max() / (2.0 * scale)
        V = V / sss - (aaa + bbb) / (2.0 * sss)
        ux.save_obj(os.path.join(save_dir, "processed_mesh.obj"), V, F, UV, Ft)

    def preprocess_reference_image(self, save_dir, input_image_path, scale=0.95, color="grey"):
        """reference :182-196 -> image/process_image.py:31-81 (`preprocess`, restated in `reference_image_on_canvas` and pinned
        against the reference's own function, tests/golden/ref_glue.npz): the matte's bounding box is cropped, scaled so the
        object spans `scale` of the 1024^2 canvas and pasted on the `color` canvas through the matte.  Matte source, in the
        reference's order (:48-54): the image's own alpha channel if it has a real one; else `self.rembg_session(image)`
        (the constructor's hook for the background-removal network the reference runs there).  With neither, the image
        is centred on the canvas un-matted and a warning says so: the conditioning then differs from the reference's."""
        src = Image.open(input_image_path)
        has_alpha = src.mode == "RGBA" and (np.asarray(src.getchannel("A")) > 0).sum() < src.size[0] * src.size[1] - 8
        if has_alpha:
            src = src.resize((1024, 1024))                                   # :183 resizes before the matte is used
            out = reference_image_on_canvas(src.convert("RGB"), src.getchannel("A"), 1024, 1024, scale, color)
        elif self.rembg_session is not None:
            rgb = src.convert("RGB").resize((1024, 1024))                    # :183
            alpha = self.rembg_session(rgb).getchannel("A")                  # image/process_image.py:52-53
            out = reference_image_on_canvas(rgb, alpha, 1024, 1024, scale, color)
        else:
            warnings.warn("preprocess_reference_image: the input image has no alpha channel and no rembg_session was given -- "
                          "the reference would matte it with RMBG-2.0 / rembg, crop and rescale; here it is used un-matted, so "
                          "rembg_image.png / processed_image.png differ from the reference's", stacklevel=2)
            img = src.convert("RGB")
            out = Image.new("RGB", (1024, 1024), color)                      # PIL grey = (128, 128, 128), as image/process_image.py:68
            im = img.copy()
            im.thumbnail((1024, 1024))
            out.paste(im, ((1024 - im.width) // 2, (1024 - im.height) // 2))
        out.save(os.path.join(save_dir, "rembg_image.png"))
        out.convert("RGB").resize((512, 512)).save(os.path.join(save_dir, "processed_image.png"))

    def render_geometry_images(self, save_dir, input_mesh_path, geometry_scale=0.95, scale=1.0, color="grey"):
        """reference :199-228."""
        out = self.video_exporter.export_condition(input_mesh_path, geometry_scale=geometry_scale, n_views=6, n_rows=2, n_cols=3,
                                                   H=512, W=512, fov_deg=49.1, scale=scale, perspective=False, orbit=False,
                                                   background=color, return_info=False, return_image=True, return_camera=True)
        out["alpha"].save(os.path.join(save_dir, "mv_alpha.png"))
        out["ccm"].save(os.path.join(save_dir, "mv_ccm.png"))
        out["normal"].save(os.path.join(save_dir, "mv_normal.png"))
        torch.save({"c2ws": out["c2ws"], "intrinsics": out["intrinsics"], "perspective": out["perspective"]},
                   os.path.join(save_dir, "camera_info.pth"))

    def infer_mv(self, save_dir, input_image_path, input_mv_image_path, add_input_mv_image_path):
        """reference :231-291: texture_gen call, then delight call, same generator (draw order matters)."""
        reference_image = Image.open(input_image_path).convert("RGB")
        strip = ux.control_grid_to_strip(np.array(Image.open(input_mv_image_path).convert("RGB")),
                                         np.array(Image.open(add_input_mv_image_path).convert("RGB")))
        steps = getattr(self.pipeline, "_num_inference_steps", 28)
        kw = dict(prompt="[MVFLUX]", prompt_embeds=None, pooled_prompt_embeds=None, height=512, width=3072, n_rows=1, n_cols=6,
                  num_inference_steps=steps, guidance_scale=3.5, max_sequence_length=512, generator=self.generator)
        self.pipeline.set_adapters(adapter_names=self.adapter_names, adapter_weights=self.weights_for_texture)
        out_image = self.pipeline(control_image=Image.fromarray(strip), dual_image=reference_image, **kw).images[0]
        out_image.save(os.path.join(save_dir, "mv_rgb_w_light.png"))
        self.pipeline.set_adapters(adapter_names=self.adapter_names, adapter_weights=self.weights_for_delight)
        out_delighted = self.pipeline(control_image=out_image, **kw).images[0]
        grid = ux.strip_to_view_grid(np.array(out_delighted))
        Image.fromarray(grid).save(os.path.join(save_dir, "mv_rgb.png"))
        return grid                                                          # uint8 [1024, 1536, 3]: the asset's finished view tile

    def reproject_and_query_field(self, save_dir, input_mesh_path, input_mv_image_path, camera_info_path, method="reproject",
                                  inpainting=False):
        """reference :312-360."""
        os.makedirs(save_dir, exist_ok=True)
        img = torch.from_numpy(np.asarray(Image.open(input_mv_image_path).convert("RGB")).astype(np.float32) / 255.0).cuda()
        H, W, C = img.shape
        HP, WP = H // 2, W // 3
        image_attrs = img.reshape(2, HP, 3, WP, C).permute(0, 2, 1, 3, 4).reshape(6, HP, WP, C)
        cam = torch.load(camera_info_path, weights_only=True, map_location="cuda")
        self.inverse_renderer.update_from_file(input_mesh_path)
        _, reprojected_uv, visable_mask, completed_uv_map = self.inverse_renderer.infer(
            self.inverse_renderer.pbr_mesh, c2ws=cam["c2ws"].cpu(), intrinsics=cam["intrinsics"].cpu(), image_attrs=image_attrs,
            perspective=cam["perspective"], H=HP, W=WP, H2D=2048, W2D=2048, method=method, kdtree_inpainting=inpainting,
            reproject_inpainting=inpainting, kdtree_n_neighbors=8, kdtree_n_neighbors_visiable=4,
            grad_norm_threhold=0.15, ray_normal_angle_threhold=100, filt_gradient_points=inpainting)
        V, F, UV, Ft = ub.load_mesh(input_mesh_path)
        # the GLB base colour goes through the reference's `tensor_to_image` (renderer_utils.py:64: clamp * 255 -> uint8, i.e.
        # truncation); the three PNGs below through torchvision's save_image (round half up)
        atlas = (completed_uv_map[0].clamp(0, 1) * 255.0).to(torch.uint8).cpu().numpy()
        ux.save_glb(os.path.join(save_dir, "textured_mesh.glb"), V, F, UV, Ft, atlas)

        def save(t, name):
            a = (t.float() * 255.0).add(0.5).clamp(0, 255).to(torch.uint8).cpu().numpy()     # torchvision save_image's quantiser
            Image.fromarray(a[..., 0] if a.shape[-1] == 1 else a).save(os.path.join(save_dir, name))
        save(reprojected_uv.any(dim=0), "visable_uv_mask.png")
        save(visable_mask[0], "valid_uv_mask.png")
        save(completed_uv_map[0], "completed_uv.png")
        self.inverse_renderer.clear()

    # ------------------------------------------------------------------ steps (reference :568-575, :624-629)
    def step_1_1(self, save_dir, input_image_path, input_mesh_path):
        cache = os.path.join(save_dir, "cache")
        self.preprocess_blank_mesh(cache, input_mesh_path)
        self.preprocess_reference_image(cache, input_image_path)
        self.render_geometry_images(cache, os.path.join(cache, "processed_mesh.obj"))
        return self.infer_mv(cache, os.path.join(cache, "processed_image.png"), os.path.join(cache, "mv_normal.png"),
                             os.path.join(cache, "mv_ccm.png"))

    def step_2_ablition(self, save_dir, input_image_path, input_mesh_path):
        cache = os.path.join(save_dir, "cache")
        self.reproject_and_query_field(os.path.join(cache, "wo_LTM"), os.path.join(cache, "processed_mesh.obj"),
                                       os.path.join(cache, "mv_rgb.png"), os.path.join(cache, "camera_info.pth"),
                                       method="reproject", inpainting=False)

    def __call__(self, save_dir: str, input_image_path: str, input_mesh_path: str, clear_cache=False) -> Tuple[str, str]:
        """reference :594-617."""
        cache = os.path.join(save_dir, "cache")
        os.makedirs(cache, exist_ok=True)
        for step in self.step_seq:
            getattr(self, step)(save_dir, input_image_path, input_mesh_path)
        shutil.copy(os.path.join(cache, "rembg_image.png"), os.path.join(save_dir, "rembg_image.png"))
        shutil.copy(os.path.join(cache, "mv_rgb.png"), os.path.join(save_dir, "mv_rgb.png"))
        shutil.copy(os.path.join(cache, "wo_LTM", "textured_mesh.glb"), os.path.join(save_dir, "textured_mesh.glb"))
        if clear_cache:
            shutil.rmtree(cache)
        return os.path.join(save_dir, "rembg_image.png"), os.path.join(save_dir, "textured_mesh.glb")


    # ------------------------------------------------------------------ multi-asset, multi-GPU (north_star)
    def run_batch(self, assets, clear_cache=False, base_seed: Optional[int] = None):
        """A batch of assets over the ranks of one box.  `assets`: list of (save_dir, input_image_path, input_mesh_path), the
        arguments of `__call__` -- the reference runs them one after another on one GPU (run.py:5-10).

        The path shards by INDEPENDENT assets: rank r runs the two FLUX calls + VAE decode of assets r, r + world, ...
        (`parallel.shard_grids`) with no data-path collective during the 28 steps (all views of one asset share one attention
        sequence, flux_piplines/texturing/pipeline.py:630-656, so views are never split).  Then the path's ONE exchange: a
        single all-gather of the finished uint8 view tiles (mv_rgb, [1024,1536,3] = 4.7 MB per asset) so that every rank
        holds every asset's tile before UV projection; each rank then bakes its own assets FROM THE GATHERED tile.
        Asset g draws from a generator seeded `base_seed + g` (run.py:5 uses 63 for its one asset), so the result of an
        asset does not depend on the world size or on which rank ran it -- the 2-GPU test compares against 1 GPU.
        Returns [(rembg_image.png, textured_mesh.glb)] for ALL assets, and leaves the gathered tiles in `self.last_tiles`.
        Works without torch.distributed (world = 1)."""
        import torch.distributed as dist
        from unitex_b200 import parallel as par
        world = dist.get_world_size() if dist.is_initialized() else 1
        rank = dist.get_rank() if dist.is_initialized() else 0
        base = self.seed if base_seed is None else base_seed
        mine = par.shard_grids(len(assets), rank, world)
        dev = self.inverse_renderer.device
        local = []
        for g in mine:
            save_dir, image_path, mesh_path = assets[g]
            os.makedirs(os.path.join(save_dir, "cache"), exist_ok=True)
            self.generator.manual_seed(par.grid_seed(base, g))
            grid = self.step_1_1(save_dir, image_path, mesh_path)
            local.append(torch.from_numpy(np.ascontiguousarray(grid)).to(dev))
        shape = tuple(local[0].shape) if local else (1024, 1536, 3)
        tiles = par.all_gather_grid_tiles(local, len(assets), shape, torch.uint8, dev)
        self.last_tiles = tiles
        for g in mine:
            save_dir, image_path, mesh_path = assets[g]
            cache = os.path.join(save_dir, "cache")
            Image.fromarray(tiles[g].cpu().numpy()).save(os.path.join(cache, "mv_rgb.png"))      # the gathered tile is what gets baked
            self.step_2_ablition(save_dir, image_path, mesh_path)
            shutil.copy(os.path.join(cache, "rembg_image.png"), os.path.join(save_dir, "rembg_image.png"))
            shutil.copy(os.path.join(cache, "mv_rgb.png"), os.path.join(save_dir, "mv_rgb.png"))
            shutil.copy(os.path.join(cache, "wo_LTM", "textured_mesh.glb"), os.path.join(save_dir, "textured_mesh.glb"))
            if clear_cache:
                shutil.rmtree(cache)
        if world > 1:
            dist.barrier()
        return [(os.path.join(a[0], "rembg_image.png"), os.path.join(a[0], "textured_mesh.glb")) for a in assets]


class CustomRGBTextureFullPipeline(RGBTextureFullPipelineBase):
    step_seq = ["step_1_1", "step_2_ablition"]       # reference pipeline.py:620-629
