"""Generates the committed golden fixtures from the ORACLE (the reference itself cannot be imported or built in this image:
diffusers / peft / nvdiffrast / slangtorch are absent -- SURVEY 8c -- so these pin the oracle against regressions and give
the GPU tests fixed inputs/outputs; they are NOT outputs of the reference).   python tests/golden/make_golden.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import bake as ob                      # noqa: E402
from oracle import flux_dit as fd                  # noqa: E402
from oracle import flux_sampler as fs              # noqa: E402
from tests.bake_meshes import analytic_color, two_spheres   # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def dit():
    cfg = fd.FluxConfig.tiny(1, 1)
    P = {k: v.to(torch.bfloat16).float() for k, v in fd.init_params(cfg, 0, norm_weight_std=0.1).items()}
    ids = fs.build_ids(16, 16, (16, 16), None)
    g = torch.Generator().manual_seed(63)
    noise = torch.randn(1, 64, 64, generator=g).to(torch.bfloat16).float()
    cond = torch.randn(1, 64, 64, generator=g).to(torch.bfloat16).float()
    out = fs.denoise(P, cfg, noise, cond, ids, num_steps=2, S_txt=128)
    np.savez_compressed(os.path.join(HERE, "dit_tiny_denoise.npz"), noise=noise.numpy(), cond=cond.numpy(), ids=ids.numpy(),
                        out=out.numpy().astype(np.float32), sigmas=fs.flow_match_sigmas(2, 64))


def bake():
    from unitex_b200.bake import generate_box_views_c2ws, generate_intrinsics
    v, f, uv, fuv = two_spheres(10, 20)
    c2ws = generate_box_views_c2ws(2.8)[[0, 1, 4, 2, 3, 5]]
    intr = generate_intrinsics(1.0, 1.0, fov=False)
    H = W = 48
    mats = torch.matmul(ob.intr_to_proj_ortho(intr), ob.c2w_to_w2c(c2ws))
    vh = torch.cat([torch.from_numpy(v), torch.ones(len(v), 1)], -1)
    rast = ob.rasterize(torch.matmul(vh, mats.permute(0, 2, 1)).numpy(), f, H, W)
    img = torch.from_numpy(analytic_color(ob.interpolate(v, rast, f)) * (rast[..., 3:4] > 0)).float()
    out = ob.infer_reproject(v, f, uv, fuv, c2ws, intr, img, H, W, 64, 64)
    info, aabb, _ = ob.lbvh_build(v, f)
    np.savez_compressed(os.path.join(HERE, "bake_two_spheres.npz"), image=img.numpy(), tid_2d=out["tid_2d"].numpy().astype(np.int32),
                        mask_vis=np.packbits(out["mask_2d_visiable"].numpy()), owner=out["owner"].numpy().astype(np.int8),
                        nn_index=out["nn_index"].numpy().astype(np.int32), color_2d=out["color_2d"].numpy().astype(np.float16),
                        lbvh_info=info, rast_mv_id=rast[..., 3].astype(np.int32))


def bake_kdtree():
    """bake_mv_to_uv_kdtree (`order_mean`, k = 9 visible / 32 invisible) on the same case + a k-NN table."""
    from unitex_b200.bake import generate_box_views_c2ws, generate_intrinsics
    z = np.load(os.path.join(HERE, "bake_two_spheres.npz"))
    v, f, uv, fuv = two_spheres(10, 20)
    c2ws = generate_box_views_c2ws(2.8)[[0, 1, 4, 2, 3, 5]]
    intr = generate_intrinsics(1.0, 1.0, fov=False)
    out = ob.infer(v, f, uv, fuv, c2ws, intr, torch.from_numpy(z["image"]), 48, 48, 64, 64, method="kdtree",
                   kdtree_method="order_mean", k_vis=9, k_invis=32)
    g = torch.Generator().manual_seed(5)
    src, dst = torch.rand(500, 3, generator=g), torch.rand(64, 3, generator=g)
    src[17] = src[3]
    dist, idx = ob.nearest_k(src, dst, 8)
    np.savez_compressed(os.path.join(HERE, "bake_kdtree.npz"), color_2d=out["color_2d"].numpy().astype(np.float16),
                        knn_src=src.numpy(), knn_dst=dst.numpy(), knn_index=idx.numpy().astype(np.int32), knn_dist=dist.numpy())


if __name__ == "__main__":
    dit()
    bake()
    bake_kdtree()
    print(sorted(os.listdir(HERE)))
