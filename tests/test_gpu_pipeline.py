"""GPU: the PBRFluxPipeline drop-in (reference flux_piplines/texturing/pipeline.py:404-700) end to end in latent space
against the oracle's restatement of the same call."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_pipeline_call_matches_oracle(lib):
    from flux_piplines.texturing.pipeline import PBRFluxPipeline
    from flux_piplines.delight.pipeline import PBRFluxPipeline as DelightPipeline
    from oracle import flux_dit as fd
    from oracle import flux_sampler as fs
    from unitex_b200.flux import FluxConfig, FluxTransformer
    assert DelightPipeline is PBRFluxPipeline
    ocfg = fd.FluxConfig.tiny(2, 2)
    P = {k: v.to(torch.bfloat16).float() for k, v in fd.init_params(ocfg, 3, norm_weight_std=0.1).items()}
    L_tex = fs.init_lora(P, ocfg, rank=8, seed=11, std=0.05)
    L_del = fs.init_lora(P, ocfg, rank=8, seed=12, std=0.05)
    cfg = FluxConfig(num_layers=2, num_single_layers=2, num_attention_heads=2, joint_attention_dim=256, pooled_projection_dim=64)
    pipe = PBRFluxPipeline(FluxTransformer(cfg).load_state_dict(P))
    pipe.load_lora_weights({"transformer." + k: v for k, v in L_tex.items()}, adapter_name="texture")
    pipe.load_lora_weights({"transformer." + k: v for k, v in L_del.items()}, adapter_name="delight")
    H = W = 256                                  # 1-view 256^2 plumbing config (BASELINE configs[0]) -> 16x16 tokens
    g = torch.Generator().manual_seed(7)
    ctrl = torch.randn(1, 16, 32, 32, generator=g)
    dual = torch.randn(1, 16, 32, 32, generator=g)
    steps = 3
    for names, weights, L, use_dual in ((["texture", "delight"], [1.0, 0.0], L_tex, True),
                                         (["texture", "delight"], [0.0, 1.0], L_del, False)):
        pipe.set_adapters(names, weights)                                   # reference pipeline.py:245,263
        out = pipe(prompt="[MVFLUX]", control_latents=ctrl, dual_latents=dual if use_dual else None, height=H, width=W,
                   n_rows=1, n_cols=1, num_inference_steps=steps, guidance_scale=3.5, max_sequence_length=128,
                   generator=torch.Generator().manual_seed(63), output_type="latent").images
        torch.cuda.synchronize()
        # oracle restatement of the same call
        noise = torch.randn((1, 16, 32, 32), generator=torch.Generator().manual_seed(63), dtype=torch.bfloat16)
        noise_p = fs.pack_latents(noise)
        cond = [fs.pack_latents(ctrl.to(torch.bfloat16))] + ([fs.pack_latents(dual.to(torch.bfloat16))] if use_dual else [])
        ids = fs.build_ids(32, 32, (32, 32), (32, 32) if use_dual else None)
        Pm = fs.merge_lora({k: v.to(torch.bfloat16) for k, v in P.items()}, L, 1.0)
        Pm = {k: v.float().cuda() for k, v in Pm.items()}
        ref = fs.denoise(Pm, ocfg, noise_p.float().cuda(), torch.cat(cond, 1).float().cuda(), ids.cuda(), num_steps=steps, S_txt=128)
        assert out.shape == (1, 256, 64)
        db = fs.psnr(out.float(), ref)
        assert db >= 40.0, f"{weights}: PSNR {db:.1f} dB"
    assert len(pipe._merged) == 2           # one resident merged weight set per adapter vector
