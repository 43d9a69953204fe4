"""CPU: pins the DiT oracle with analytic known-answer tests (the reference ships no golden vectors)."""
import math

import numpy as np
import torch

from oracle import flux_dit as fd
from oracle import flux_sampler as fs


def test_flops_match_survey_table():
    cfg = fd.FluxConfig()
    assert abs(fd.dit_flops(cfg, 9728) / 1e12 - 191.90) < 0.02
    assert abs(fd.dit_flops(cfg, 8704) / 1e12 - 165.46) < 0.02
    assert abs(fd.dit_flops(cfg, 13824) / 1e12 - 312.35) < 0.02


def test_calculate_shift_and_sigmas_closed_form():
    # pipeline.py:59-69 with the FLUX.1-dev scheduler config: mu(6144)=1.4967, mu(4096)=1.15, mu(256)=0.5
    assert abs(fs.calculate_shift(6144) - 1.4966667) < 1e-6
    assert abs(fs.calculate_shift(4096) - 1.15) < 1e-12
    assert abs(fs.calculate_shift(256) - 0.5) < 1e-12
    for S in (256, 4096, 6144):
        sig = fs.flow_match_sigmas(28, S)
        assert sig.shape == (29,) and sig[0] == 1.0 and sig[-1] == 0.0
        mu = fs.calculate_shift(S)
        for i in (1, 7, 27):
            s = 1.0 - i / 28.0 * (1 - 1 / 28.0) * 28 / 27 if False else np.linspace(1, 1 / 28, 28)[i]
            want = math.exp(mu) / (math.exp(mu) + (1 / s - 1))
            assert abs(sig[i] - want) < 1e-6
        assert np.all(np.diff(sig) < 0)


def test_pack_unpack_roundtrip_and_layout():
    lat = torch.arange(1 * 16 * 8 * 12, dtype=torch.float32).view(1, 16, 8, 12)
    p = fs.pack_latents(lat)
    assert p.shape == (1, 4 * 6, 64)
    # token (y,x) channel c*4 + dy*2 + dx  == lat[c, 2y+dy, 2x+dx]   (pipeline.py:240-249)
    assert p[0, 1 * 6 + 2, 5 * 4 + 1 * 2 + 0] == lat[0, 5, 3, 4]
    assert torch.equal(fs.unpack_latents(p, 64, 96), lat)


def test_ids_offsets_follow_reference():
    # noise (0,0); control (y+HL/2, x); dual (y+HL/2, x+WL/2)   pipeline.py:303-312,335-344,384-393
    ids = fs.build_ids(128, 128, control_hw=(128, 128), dual_hw=(64, 64))
    assert ids.shape == (4096 + 4096 + 1024, 3)
    assert ids[:, 0].abs().max() == 0
    assert ids[0].tolist() == [0, 0, 0] and ids[4095].tolist() == [0, 63, 63]
    assert ids[4096].tolist() == [0, 64, 0] and ids[8191].tolist() == [0, 127, 63]
    assert ids[8192].tolist() == [0, 64, 64] and ids[-1].tolist() == [0, 95, 95]
    # bf16 ids (pipeline.py:571) are exact in this range
    assert torch.equal(fs.build_ids(128, 128, (128, 128), (64, 64), dtype=torch.bfloat16).float(), ids)


def test_rope_closed_form():
    cfg = fd.FluxConfig.tiny()
    ids = torch.tensor([[0., 3., 7.], [0., 0., 0.]])
    cos, sin = fd.rope_table(ids, cfg)
    assert cos.shape == (2, 128)
    assert torch.allclose(cos[1], torch.ones(128)) and torch.allclose(sin[1], torch.zeros(128))
    # axis 1 (dims 16..72): pair i has angle 3 * 1e4^(-2i/56)
    for i in (0, 5, 27):
        a = 3.0 * 10000.0 ** (-2 * i / 56)
        assert abs(cos[0, 16 + 2 * i].item() - math.cos(a)) < 1e-6 and cos[0, 16 + 2 * i] == cos[0, 16 + 2 * i + 1]
        assert abs(sin[0, 16 + 2 * i].item() - math.sin(a)) < 1e-6
    x = torch.randn(1, 1, 2, 128)
    y = fd.apply_rope(x, cos, sin)
    assert torch.allclose(y[0, 0, 1], x[0, 0, 1])
    # rotation of pair (x0,x1) by angle a
    a = 7.0 * 10000.0 ** (-2 * 2 / 56)
    x0, x1 = x[0, 0, 0, 72 + 4].item(), x[0, 0, 0, 72 + 5].item()
    assert abs(y[0, 0, 0, 72 + 4].item() - (x0 * math.cos(a) - x1 * math.sin(a))) < 1e-5
    assert abs(y[0, 0, 0, 72 + 5].item() - (x1 * math.cos(a) + x0 * math.sin(a))) < 1e-5
    # norms preserved
    assert torch.allclose(y.norm(), x.norm(), rtol=1e-5)


def test_sinusoid_and_rmsnorm():
    e = fd.sinusoid_256(torch.tensor([500.0]))
    assert e.shape == (1, 256)
    assert abs(e[0, 0].item() - math.cos(500.0)) < 1e-4 and abs(e[0, 128].item() - math.sin(500.0)) < 1e-4
    f5 = math.exp(-math.log(10000.0) * 5 / 128)
    assert abs(e[0, 5].item() - math.cos(500 * f5)) < 1e-4
    x = torch.randn(3, 128)
    w = torch.rand(128) + 0.5
    y = fd.rms_norm(x, w)
    assert torch.allclose(y, x / torch.sqrt((x * x).mean(-1, keepdim=True) + 1e-6) * w, atol=1e-6)


def test_attention_matches_explicit_softmax():
    cfg = fd.FluxConfig.tiny()
    P = fd.init_params(cfg, 1, norm_weight_std=0.1)
    x = torch.randn(1, 24, cfg.inner_dim)
    ctx = torch.randn(1, 8, cfg.inner_dim)
    ids = torch.cat([torch.zeros(8, 3), fs.latent_image_ids(4, 6)])
    cos, sin = fd.rope_table(ids, cfg)
    ao, co = fd.joint_attention(P, "transformer_blocks.0.attn.", cfg, x, ctx, cos, sin)
    # explicit re-derivation (attention_processor.py:43-99)
    H = cfg.num_attention_heads
    pre = "transformer_blocks.0.attn."
    lin = lambda n, t: torch.nn.functional.linear(t, P[pre + n + ".weight"], P[pre + n + ".bias"])
    hd = lambda t: t.view(1, -1, H, 128).transpose(1, 2)
    q = torch.cat([fd.rms_norm(hd(lin("add_q_proj", ctx)), P[pre + "norm_added_q.weight"]),
                   fd.rms_norm(hd(lin("to_q", x)), P[pre + "norm_q.weight"])], 2)
    k = torch.cat([fd.rms_norm(hd(lin("add_k_proj", ctx)), P[pre + "norm_added_k.weight"]),
                   fd.rms_norm(hd(lin("to_k", x)), P[pre + "norm_k.weight"])], 2)
    v = torch.cat([hd(lin("add_v_proj", ctx)), hd(lin("to_v", x))], 2)
    q, k = fd.apply_rope(q, cos, sin), fd.apply_rope(k, cos, sin)
    w = torch.softmax(q @ k.transpose(-1, -2) / math.sqrt(128), -1)
    o = (w @ v).transpose(1, 2).reshape(1, 32, -1)
    assert torch.allclose(ao, lin("to_out.0", o[:, 8:]), atol=1e-5)
    assert torch.allclose(co, lin("to_add_out", o[:, :8]), atol=1e-5)


def test_lora_merge_equals_unmerged_evaluation():
    cfg = fd.FluxConfig.tiny(1, 1)
    P = fd.init_params(cfg, 2)
    L = fs.init_lora(P, cfg, rank=4, seed=3, std=0.05)
    Pm = fs.merge_lora(P, L, scale=0.7)
    x = torch.randn(5, cfg.inner_dim)
    n = "transformer_blocks.0.attn.to_q"
    want = torch.nn.functional.linear(x, P[n + ".weight"], P[n + ".bias"]) + 0.7 * (x @ L[n + ".lora_A.weight"].T) @ L[n + ".lora_B.weight"].T
    got = torch.nn.functional.linear(x, Pm[n + ".weight"], Pm[n + ".bias"])
    assert torch.allclose(got, want, atol=1e-5)
    assert torch.equal(Pm["x_embedder.weight"], L["x_embedder.weight"])
    assert torch.equal(Pm["single_transformer_blocks.0.proj_mlp.weight"], P["single_transformer_blocks.0.proj_mlp.weight"])


def test_denoise_condition_tokens_and_bf16_yardstick():
    cfg = fd.FluxConfig.tiny()
    P = fd.init_params(cfg, 0, norm_weight_std=0.1)
    ids = fs.build_ids(16, 16, (16, 16), (16, 16))
    g = torch.Generator().manual_seed(63)
    noise, cond = torch.randn(1, 64, 64, generator=g), torch.randn(1, 128, 64, generator=g)
    tr = []
    out = fs.denoise(P, cfg, noise, cond, ids, num_steps=2, S_txt=128, trace=tr)
    assert out.shape == (1, 64, 64) and torch.isfinite(out).all() and len(tr) == 2
    # one Euler step by hand
    sig = fs.flow_match_sigmas(2, 64)
    t_in = (torch.tensor([sig[0] * 1000.0]).to(torch.bfloat16) / 1000)
    lat = torch.cat([noise, cond], 1)
    v = fd.flux_forward(P, cfg, lat, t_in, torch.tensor([3.5]), torch.zeros(1, cfg.pooled_projection_dim),
                        torch.zeros(1, 128, cfg.joint_attention_dim), torch.zeros(128, 3), ids)
    want = (lat + (sig[1] - sig[0]) * v)[:, :64]
    assert torch.allclose(tr[0], want, atol=1e-5)
    Pb = {k: v.bfloat16() for k, v in P.items()}
    outb = fs.denoise(Pb, cfg, noise, cond, ids, num_steps=2, S_txt=128)
    assert fs.psnr(outb.float(), out) > 40.0      # the reference's own bf16 rounding stays above the 40 dB bar
