"""CPU: the oracle reproduces the committed golden fixtures (tests/golden/make_golden.py made them FROM the oracle -- the
reference cannot run in this image; parity is unpinned upstream, these pin us)."""
import os

import numpy as np
import torch

from oracle import bake as ob
from oracle import flux_dit as fd
from oracle import flux_sampler as fs
from tests.bake_meshes import two_spheres

G = os.path.join(os.path.dirname(__file__), "golden")


def test_dit_golden():
    z = np.load(os.path.join(G, "dit_tiny_denoise.npz"))
    cfg = fd.FluxConfig.tiny(1, 1)
    P = {k: v.to(torch.bfloat16).float() for k, v in fd.init_params(cfg, 0, norm_weight_std=0.1).items()}
    out = fs.denoise(P, cfg, torch.from_numpy(z["noise"]), torch.from_numpy(z["cond"]), torch.from_numpy(z["ids"]), num_steps=2, S_txt=128)
    assert np.array_equal(z["sigmas"], fs.flow_match_sigmas(2, 64))
    assert fs.psnr(out, torch.from_numpy(z["out"])) > 80.0          # fp32 CPU matmul order may differ between hosts


def test_bake_golden_integers_exact():
    from unitex_b200.bake import generate_box_views_c2ws, generate_intrinsics
    z = np.load(os.path.join(G, "bake_two_spheres.npz"))
    v, f, uv, fuv = two_spheres(10, 20)
    c2ws = generate_box_views_c2ws(2.8)[[0, 1, 4, 2, 3, 5]]
    intr = generate_intrinsics(1.0, 1.0, fov=False)
    out = ob.infer_reproject(v, f, uv, fuv, c2ws, intr, torch.from_numpy(z["image"]), 48, 48, 64, 64)
    assert np.array_equal(out["tid_2d"].numpy().astype(np.int32), z["tid_2d"])
    assert np.array_equal(np.packbits(out["mask_2d_visiable"].numpy()), z["mask_vis"])
    assert np.array_equal(out["owner"].numpy().astype(np.int8), z["owner"])
    assert np.array_equal(out["nn_index"].numpy().astype(np.int32), z["nn_index"])
    assert np.abs(out["color_2d"].numpy() - z["color_2d"].astype(np.float32)).max() < 2e-3
    info, _, _ = ob.lbvh_build(v, f)
    assert np.array_equal(info, z["lbvh_info"])


def test_bake_kdtree_golden():
    from unitex_b200.bake import generate_box_views_c2ws, generate_intrinsics
    z, zk = np.load(os.path.join(G, "bake_two_spheres.npz")), np.load(os.path.join(G, "bake_kdtree.npz"))
    v, f, uv, fuv = two_spheres(10, 20)
    c2ws = generate_box_views_c2ws(2.8)[[0, 1, 4, 2, 3, 5]]
    intr = generate_intrinsics(1.0, 1.0, fov=False)
    out = ob.infer(v, f, uv, fuv, c2ws, intr, torch.from_numpy(z["image"]), 48, 48, 64, 64, method="kdtree",
                   kdtree_method="order_mean", k_vis=9, k_invis=32)
    assert np.abs(out["color_2d"].numpy() - zk["color_2d"].astype(np.float32)).max() < 2e-3
    dist, idx = ob.nearest_k(torch.from_numpy(zk["knn_src"]), torch.from_numpy(zk["knn_dst"]), 8)
    assert np.array_equal(idx.numpy().astype(np.int32), zk["knn_index"]) and np.array_equal(dist.numpy(), zk["knn_dist"])
