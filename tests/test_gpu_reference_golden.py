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
