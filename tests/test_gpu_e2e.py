"""GPU: the whole drop-in path CustomRGBTextureFullPipeline(...)(save_dir, image, mesh) with random-init weights:
G-buffer render -> texture_gen + delight FLUX calls (VAE encode/decode inside) -> UV bake -> GLB (reference run.py:1-10)."""
import os
import struct

import numpy as np
import pytest
import torch

from tests.bake_meshes import two_spheres

pytestmark = pytest.mark.gpu


def test_custom_rgb_texture_full_pipeline(lib, tmp_path):
    from PIL import Image
    from pipeline import CustomRGBTextureFullPipeline
    from unitex_b200.export import save_obj
    from unitex_b200.flux import FluxConfig
    v, f, uv, fuv = two_spheres(24, 48)
    mesh_path, img_path = str(tmp_path / "mesh.obj"), str(tmp_path / "image.png")
    save_obj(mesh_path, v * 3.7 + np.array([5.0, -2.0, 1.0], np.float32), f, (uv + 1) / 2, fuv)   # arbitrary frame: preprocess_blank_mesh normalises it
    Image.fromarray(np.random.default_rng(0).integers(0, 255, (256, 256, 3), dtype=np.uint8)).save(img_path)
    cfg = FluxConfig(num_layers=1, num_single_layers=1, num_attention_heads=2, joint_attention_dim=256, pooled_projection_dim=64)
    pipe = CustomRGBTextureFullPipeline(pretrain_models=cfg, super_resolutions=False, seed=63)
    pipe.pipeline._num_inference_steps = 2
    save_dir = str(tmp_path / "out")
    png, glb = pipe(save_dir, img_path, mesh_path, clear_cache=False)
    torch.cuda.synchronize()
    assert png.endswith("rembg_image.png") and glb.endswith("textured_mesh.glb") and os.path.exists(png) and os.path.exists(glb)
    cache = os.path.join(save_dir, "cache")
    for name in ("processed_mesh.obj", "processed_image.png", "mv_alpha.png", "mv_ccm.png", "mv_normal.png", "camera_info.pth",
                 "mv_rgb_w_light.png", "mv_rgb.png", "wo_LTM/textured_mesh.glb", "wo_LTM/visable_uv_mask.png",
                 "wo_LTM/valid_uv_mask.png", "wo_LTM/completed_uv.png"):
        assert os.path.exists(os.path.join(cache, name)), name
    assert Image.open(os.path.join(cache, "mv_rgb.png")).size == (1536, 1024)
    assert Image.open(os.path.join(cache, "mv_rgb_w_light.png")).size == (3072, 512)
    alpha = np.asarray(Image.open(os.path.join(cache, "mv_alpha.png")))
    assert alpha.shape == (1024, 1536) and 0.05 < (alpha > 0).mean() < 0.9
    valid = np.asarray(Image.open(os.path.join(cache, "wo_LTM", "valid_uv_mask.png")))
    vis = np.asarray(Image.open(os.path.join(cache, "wo_LTM", "visable_uv_mask.png")))
    assert valid.shape == (2048, 2048) and (vis > 0).sum() > 0.5 * (valid > 0).sum()
    assert struct.unpack("<I", open(glb, "rb").read(4))[0] == 0x46546C67


def test_full_pipeline_on_a_mesh_without_uvs(lib, tmp_path):
    """A UV-less mesh goes through the drop-in call too: per-triangle atlas (unitex_b200/uv_atlas.py) instead of the reference's
    open3d / UVAtlas chain [ext], with a warning; every cache file and the GLB are written, most covered texels are seen by a view."""
    from PIL import Image
    from pipeline import CustomRGBTextureFullPipeline
    from unitex_b200.export import save_obj
    from unitex_b200.flux import FluxConfig
    v, f, _, _ = two_spheres(16, 32)
    mesh_path, img_path = str(tmp_path / "mesh.obj"), str(tmp_path / "image.png")
    save_obj(mesh_path, v, f)                                              # no vt lines
    Image.fromarray(np.random.default_rng(1).integers(0, 255, (128, 128, 3), dtype=np.uint8)).save(img_path)
    cfg = FluxConfig(num_layers=1, num_single_layers=1, num_attention_heads=2, joint_attention_dim=256, pooled_projection_dim=64)
    pipe = CustomRGBTextureFullPipeline(pretrain_models=cfg, seed=63)
    pipe.pipeline._num_inference_steps = 2
    with pytest.warns(UserWarning, match="no UVs"):
        png, glb = pipe(str(tmp_path / "out"), img_path, mesh_path)
    torch.cuda.synchronize()
    assert os.path.exists(glb) and struct.unpack("<I", open(glb, "rb").read(4))[0] == 0x46546C67
    cache = os.path.join(str(tmp_path / "out"), "cache", "wo_LTM")
    valid = np.asarray(Image.open(os.path.join(cache, "valid_uv_mask.png")))
    vis = np.asarray(Image.open(os.path.join(cache, "visable_uv_mask.png")))
    assert (valid > 0).sum() > 20_000 and (vis > 0).sum() > 0.5 * (valid > 0).sum()
