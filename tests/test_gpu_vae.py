"""GPU parity of the B200 VAE (im2col + tcgen05 GEMM convs, GroupNorm+SiLU, mid attention) against the fp32 oracle.
Tolerance: PSNR >= 40 dB on the decoded image / encoded moments, or within 3 dB of what the reference's own bf16 eager
chain scores on the same yardstick (the reference runs the VAE in bf16: pipeline.py:102, :690)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _setup(seed=0):
    from oracle import flux_sampler as fs
    from oracle import vae as ov
    from unitex_b200.vae import AutoencoderKLB200
    cfg = ov.VaeConfig.tiny()
    P = {k: v.to(torch.bfloat16).float() for k, v in ov.init_params(cfg, seed).items()}
    eng = AutoencoderKLB200(P, cfg.block_out_channels, cfg.layers_per_block, cfg.latent_channels, cfg.in_channels,
                            cfg.norm_num_groups, cfg.scaling_factor, cfg.shift_factor)
    return fs, ov, cfg, P, eng


def _bar(fs, ours, ref, bf16_chain):
    db, db16 = fs.psnr(ours, ref), fs.psnr(bf16_chain, ref)
    assert db >= min(40.0, db16 - 3.0), f"PSNR {db:.1f} dB (bf16 eager oracle: {db16:.1f} dB)"
    return db


def test_decode_matches_oracle(lib):
    fs, ov, cfg, P, eng = _setup()
    z = torch.randn(1, 16, 8, 16, generator=torch.Generator().manual_seed(1)).to(torch.bfloat16)
    img = eng.decode(z.cuda())
    torch.cuda.synchronize()
    assert img.shape == (1, 3, 64, 128) and torch.isfinite(img.float()).all()
    Pg = {k: v.cuda() for k, v in P.items()}
    ref = ov.decode(Pg, cfg, z.float().cuda())
    chain = ov.decode({k: v.to(torch.bfloat16) for k, v in Pg.items()}, cfg, z.cuda()).float()
    _bar(fs, img.float(), ref, chain)


def test_encode_matches_oracle_and_sampling(lib):
    fs, ov, cfg, P, eng = _setup(3)
    g = torch.Generator().manual_seed(2)
    img = (torch.rand(1, 3, 64, 64, generator=g) * 2 - 1).to(torch.bfloat16)
    mean, logvar = eng.encode_moments(img.cuda())
    torch.cuda.synchronize()
    Pg = {k: v.cuda() for k, v in P.items()}
    rm, rl = ov.encode_moments(Pg, cfg, img.float().cuda())
    cm, cl = ov.encode_moments({k: v.to(torch.bfloat16) for k, v in Pg.items()}, cfg, img.cuda())
    assert mean.shape == (1, 16, 8, 8)
    _bar(fs, mean, rm, cm.float())
    _bar(fs, logvar, rl, cl.float())
    # latent_dist.sample(generator): same draw as a CPU generator would give diffusers' randn_tensor
    z = eng.encode_sample(img.cuda(), torch.Generator().manual_seed(9))
    noise = torch.randn(mean.shape, generator=torch.Generator().manual_seed(9), dtype=torch.bfloat16).cuda().float()
    assert torch.allclose(z.float(), (mean + torch.exp(0.5 * logvar) * noise).to(torch.bfloat16).float())


def test_pipeline_with_images_end_to_end(lib):
    """PIL in -> VAE encode -> denoise -> VAE decode -> PIL out through the reference's call signature."""
    from PIL import Image
    import numpy as np
    from flux_piplines.texturing.pipeline import PBRFluxPipeline
    from oracle import flux_dit as fd
    from unitex_b200.flux import FluxConfig, FluxTransformer
    fs, ov, cfg, P, vae = _setup(5)
    ocfg = fd.FluxConfig.tiny(1, 1)
    Pt = {k: v.to(torch.bfloat16).float() for k, v in fd.init_params(ocfg, 0, norm_weight_std=0.1).items()}
    tr = FluxTransformer(FluxConfig(num_layers=1, num_single_layers=1, num_attention_heads=2, joint_attention_dim=256,
                                    pooled_projection_dim=64)).load_state_dict(Pt)
    pipe = PBRFluxPipeline(tr, vae)
    rng = np.random.default_rng(0)
    ctrl = Image.fromarray(rng.integers(0, 255, (128, 128, 3), dtype=np.uint8))
    dual = Image.fromarray(rng.integers(0, 255, (64, 64, 3), dtype=np.uint8))
    out = pipe(prompt="[MVFLUX]", control_image=ctrl, dual_image=dual, height=128, width=128, n_rows=1, n_cols=1,
               num_inference_steps=2, guidance_scale=3.5, max_sequence_length=128,
               generator=torch.Generator().manual_seed(63)).images
    assert len(out) == 1 and out[0].size == (128, 128)
    a = np.asarray(out[0])
    assert a.dtype == np.uint8 and a.std() > 1.0


@pytest.mark.parametrize("N,H,W,C,Co,with_res", [(1, 16, 128, 64, 128, False), (2, 8, 32, 128, 64, True), (1, 4, 256, 64, 256, True),
                                                 (1, 16, 16, 64, 64, False), (3, 2, 64, 192, 128, False)])
def test_implicit_conv3x3_matches_conv2d(lib, N, H, W, C, Co, with_res):
    """The implicit-GEMM convolution (TMA boxes shifted by the tap, zero fill as padding) against torch conv2d in fp32 on the
    same bf16-rounded operands; M = N*H*W is not always a multiple of the 128-row tile (tail boxes past the last image)."""
    from unitex_b200 import _lib
    L = _lib.load()
    g = torch.Generator(device="cuda").manual_seed(N * 1000 + W)
    x = torch.randn(N, H, W, C, device="cuda", generator=g).to(torch.bfloat16)
    w = (torch.randn(Co, 3, 3, C, device="cuda", generator=g) * 0.05).to(torch.bfloat16)
    b = torch.randn(Co, device="cuda", generator=g).to(torch.bfloat16)
    res = torch.randn(N * H * W, Co, device="cuda", generator=g).to(torch.bfloat16) if with_res else None
    ref = torch.nn.functional.conv2d(x.float().permute(0, 3, 1, 2), w.float().permute(0, 3, 1, 2), b.float(), padding=1)
    ref = ref.permute(0, 2, 3, 1).reshape(N * H * W, Co)
    y = res.clone() if with_res else torch.empty(N * H * W, Co, device="cuda", dtype=torch.bfloat16)
    one = torch.ones(Co, device="cuda")
    if with_res:
        ref = ref + res.float()
    st = torch.cuda.current_stream().cuda_stream
    _lib.check(L.utx_conv3x3_nhwc(x.data_ptr(), N, H, W, C, w.data_ptr(), b.data_ptr(), Co, y.data_ptr(), Co,
                                  one.data_ptr() if with_res else None, y.data_ptr() if with_res else None, Co if with_res else 0, st),
               "utx_conv3x3_nhwc")
    torch.cuda.synchronize()
    err = (y.float() - ref).abs().max().item()
    assert err <= 2e-2 * ref.abs().max().item() + 1e-2, f"max abs err {err}"


@pytest.mark.parametrize("M,N,lds", [(64, 16384, 16384), (33, 1000, 1024), (17, 4096, 4096), (9, 1002, 1002), (5, 17000, 17000)])
def test_softmax_rows(lib, M, N, lds):
    """The mid block's row softmax (fp32 scores -> bf16 probabilities): the register-resident kernel (N % 4 == 0, N <= 16384,
    ragged last float4 block, padded rows) and the three-pass kernel behind the same entry point."""
    import ctypes as C
    from unitex_b200 import _lib
    g = torch.Generator(device="cuda").manual_seed(N)
    S = torch.randn(M, lds, device="cuda", generator=g) * 4.0
    P = torch.full((M, lds), 7.0, device="cuda", dtype=torch.bfloat16)
    _lib.check(_lib.load().utx_softmax_rows(C.c_void_p(S.data_ptr()), lds, C.c_void_p(P.data_ptr()), lds, M, N, None), "utx_softmax_rows")
    torch.cuda.synchronize()
    want = torch.softmax(S[:, :N].double(), dim=-1)
    got = P[:, :N].double()
    assert (got - want).abs().max().item() <= 2.0 ** -8 * want.max().item() + 1e-7      # one bf16 rounding of values <= max
    assert (got.sum(-1) - 1.0).abs().max().item() < 2.0 ** -8 + 1e-4                  # every term within half a bf16 ulp (<= 2^-8 relative)
    assert (P[:, N:] == 7.0).all()                                                    # nothing written past the row
