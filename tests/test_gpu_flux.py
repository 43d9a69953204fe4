"""GPU parity of the whole DiT forward / Euler loop against the fp32 oracle (north_star: PSNR >= 40 dB on the latent)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

PSNR_MIN_DB = 40.0     # BASELINE.json north_star tolerance


def _setup(layers=(2, 2), heads=2, HL=16, WL=16, ctrl=(16, 16), dual=(16, 16), s_txt=128, seed=0, lora=False):
    from oracle import flux_dit as fd
    from oracle import flux_sampler as fs
    from unitex_b200.flux import FluxConfig, FluxTransformer
    ocfg = fd.FluxConfig.tiny(layers[0], layers[1], heads)
    P = fd.init_params(ocfg, seed, norm_weight_std=0.1)
    # the engine stores bf16: the oracle sees the same bf16-rounded weights, evaluated in fp32
    P = {k: v.to(torch.bfloat16).float() for k, v in P.items()}
    if lora:
        L = fs.init_lora(P, ocfg, rank=8, seed=seed + 1, std=0.05)
    cfg = FluxConfig(num_layers=layers[0], num_single_layers=layers[1], num_attention_heads=heads,
                     joint_attention_dim=ocfg.joint_attention_dim, pooled_projection_dim=ocfg.pooled_projection_dim)
    eng = FluxTransformer(cfg).load_state_dict(P)
    if lora:
        eng.merge_lora_(L, 0.8)
        P = fs.merge_lora({k: v.to(torch.bfloat16) for k, v in P.items()}, L, 0.8)
        P = {k: v.float() for k, v in P.items()}
    img_ids = fs.build_ids(HL, WL, ctrl, dual)
    g = torch.Generator().manual_seed(63)
    s_noise = (HL // 2) * (WL // 2)
    noise = torch.randn(1, s_noise, 64, generator=g).to(torch.bfloat16)
    cond = torch.randn(1, img_ids.shape[0] - s_noise, 64, generator=g).to(torch.bfloat16)
    return fd, fs, ocfg, P, eng, img_ids, noise, cond, s_txt, s_noise


@pytest.mark.parametrize("lora", [False, True])
def test_forward_matches_oracle(lib, lora):
    fd, fs, ocfg, P, eng, img_ids, noise, cond, s_txt, s_noise = _setup(lora=lora)
    ids = torch.cat([torch.zeros(s_txt, 3), img_ids])
    g = torch.Generator().manual_seed(5)
    enc = (torch.randn(s_txt, ocfg.joint_attention_dim, generator=g) * 0.5).to(torch.bfloat16)
    pooled = torch.randn(ocfg.pooled_projection_dim, generator=g)
    eng.prepare(ids, enc, pooled, s_txt=s_txt)
    lat = torch.cat([noise, cond], 1)[0].cuda().contiguous()
    t_in = float(torch.tensor(0.73).to(torch.bfloat16))
    v = eng.forward(lat, t_in, 3.5)
    torch.cuda.synchronize()
    Pg = {k: w.cuda() for k, w in P.items()}
    ref = fd.flux_forward(Pg, ocfg, lat[None].float(), torch.tensor([t_in]).cuda(), torch.tensor([3.5]).cuda(),
                          pooled[None].cuda(), enc[None].float().cuda(), torch.zeros(s_txt, 3).cuda(), img_ids.cuda())[0]
    db = fs.psnr(v.float(), ref)
    assert torch.isfinite(v.float()).all()
    assert db >= PSNR_MIN_DB, f"forward PSNR {db:.1f} dB"


def test_denoise_matches_oracle_and_keeps_condition(lib):
    fd, fs, ocfg, P, eng, img_ids, noise, cond, s_txt, s_noise = _setup(layers=(2, 3), HL=32, WL=16, ctrl=(32, 16), dual=None)
    ids = torch.cat([torch.zeros(s_txt, 3), img_ids])
    eng.prepare(ids, None, None, s_txt=s_txt)
    lat = torch.cat([noise, cond], 1)[0].cuda().contiguous()
    lat0 = lat.clone()
    steps = 4
    sig = fs.flow_match_sigmas(steps, s_noise)
    eng.denoise_(lat, s_noise, sig, 3.5)
    torch.cuda.synchronize()
    assert torch.equal(lat[s_noise:], lat0[s_noise:])          # clean condition tokens never change (:644-645)
    Pg = {k: w.cuda() for k, w in P.items()}
    ref = fs.denoise(Pg, ocfg, noise.float().cuda(), cond.float().cuda(), img_ids.cuda(), num_steps=steps, S_txt=s_txt)[0]
    db = fs.psnr(lat[:s_noise].float(), ref)
    assert db >= PSNR_MIN_DB, f"denoise PSNR {db:.1f} dB"


def test_graph_replay_is_bit_identical_to_eager(lib, monkeypatch):
    """utx_flux_denoise replays a captured CUDA graph of the step from the second step on; the per-step scalars come from
    device memory.  Same bits as the eager launches, step by step and as one call."""
    fd, fs, ocfg, P, eng, img_ids, noise, cond, s_txt, s_noise = _setup(layers=(2, 2), HL=32, WL=16, ctrl=(32, 16), dual=(16, 16))
    ids = torch.cat([torch.zeros(s_txt, 3), img_ids])
    sig = fs.flow_match_sigmas(6, s_noise)
    lat0 = torch.cat([noise, cond], 1)[0].cuda().contiguous()

    def run(e, one_call):
        e.prepare(ids, None, None, s_txt=s_txt)
        lat = lat0.clone()
        if one_call:
            e.denoise_(lat, s_noise, sig, 3.5)
        else:
            for i in range(6):
                e.denoise_(lat, s_noise, sig[i:i + 2], 3.5)
        torch.cuda.synchronize()
        return lat

    a = run(eng, False)
    assert eng.graph_replays() >= 4                          # steps 2..6 of the same (latents, s_noise) ran as graph launches
    b = run(eng, True)
    monkeypatch.setenv("UTX_FLUX_GRAPH", "0")
    from unitex_b200.flux import FluxConfig, FluxTransformer
    eager = FluxTransformer(eng.cfg).load_state_dict(P)      # a fresh handle reads the knob
    c = run(eager, False)
    assert eager.graph_replays() == 0
    assert torch.equal(a, c) and torch.equal(b, c)


def test_merge_lora_key_handling(lib):
    """Adapter files: `.alpha` entries scale their module (kohya style), unknown suffixes / modules raise instead of being
    broadcast into a bias, `modules_to_save` replacements are applied only when asked (the last active adapter's copy wins)."""
    fd, fs, ocfg, P, eng, *_ = _setup(layers=(1, 1))
    n = "transformer_blocks.0.attn.to_q"
    g = torch.Generator().manual_seed(2)
    A, B = torch.randn(4, 256, generator=g) * 0.1, torch.randn(256, 4, generator=g) * 0.1
    k, r0, r1 = eng.where[n]
    W0 = eng.T["w_" + k][r0:r1].clone()
    eng.merge_lora_({n + ".lora_A.weight": A, n + ".lora_B.weight": B, n + ".alpha": torch.tensor(2.0)}, 1.0)
    want = (W0.float().cpu() + (2.0 / 4) * (B @ A)).to(torch.bfloat16)
    assert (eng.T["w_" + k][r0:r1].cpu().float() - want.float()).abs().max().item() <= 2 ** -7 * want.float().abs().max().item()
    with pytest.raises(KeyError):
        eng.merge_lora_({n + ".lora_A.weight": A, n + ".lora_B.weight": B, n + ".dora_scale": torch.ones(256)}, 1.0)
    with pytest.raises(KeyError):
        eng.merge_lora_({"no_such_block.lora_A.weight": A, "no_such_block.lora_B.weight": B}, 1.0)
    xk = eng.where["x_embedder"][0]
    X0 = eng.T["w_" + xk].clone()
    rep = {"x_embedder.weight": torch.full((256, 64), 0.5), "x_embedder.bias": torch.zeros(256)}
    eng.merge_lora_(rep, 1.0, replace_modules=False)
    assert torch.equal(eng.T["w_" + xk], X0)
    eng.merge_lora_(rep, 1.0)
    assert torch.equal(eng.T["w_" + xk].float().cpu(), torch.full((256, 64), 0.5))
    with pytest.raises(ValueError):
        eng.merge_lora_({"x_embedder.weight": torch.zeros(3, 3)}, 1.0)


def test_product_path_does_not_import_oracle():
    import subprocess, sys
    code = ("import sys; import unitex_b200.flux, unitex_b200.ops; "
            "assert not any(m == 'oracle' or m.startswith('oracle.') for m in sys.modules), 'oracle imported'")
    subprocess.run([sys.executable, "-c", code], check=True)


def test_denoise_matches_committed_golden(lib):
    """CUDA path vs tests/golden/dit_tiny_denoise.npz (oracle-generated fixture, make_golden.py)."""
    import os
    import numpy as np
    from oracle import flux_dit as fd
    from oracle import flux_sampler as fs
    from unitex_b200.flux import FluxConfig, FluxTransformer
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "dit_tiny_denoise.npz"))
    ocfg = fd.FluxConfig.tiny(1, 1)
    P = {k: v.to(torch.bfloat16).float() for k, v in fd.init_params(ocfg, 0, norm_weight_std=0.1).items()}
    eng = FluxTransformer(FluxConfig(num_layers=1, num_single_layers=1, num_attention_heads=2, joint_attention_dim=256,
                                     pooled_projection_dim=64)).load_state_dict(P)
    eng.prepare(torch.cat([torch.zeros(128, 3), torch.from_numpy(z["ids"])]), None, None, s_txt=128)
    lat = torch.cat([torch.from_numpy(z["noise"]), torch.from_numpy(z["cond"])], 1)[0].to(torch.bfloat16).cuda().contiguous()
    eng.denoise_(lat, 64, z["sigmas"], 3.5)
    torch.cuda.synchronize()
    db = fs.psnr(lat[:64].float().cpu(), torch.from_numpy(z["out"])[0])
    assert db >= PSNR_MIN_DB, db


def test_ragged_token_counts_and_empty_calls(lib):
    """Token counts that are not multiples of the 128-row tiles (txt 77, img 333) and degenerate calls through the C ABI."""
    from oracle import flux_dit as fd
    from oracle import flux_sampler as fs
    from unitex_b200 import ops
    from unitex_b200.flux import FluxConfig, FluxTransformer
    ocfg = fd.FluxConfig.tiny(1, 1)
    P = {k: v.to(torch.bfloat16).float() for k, v in fd.init_params(ocfg, 7, norm_weight_std=0.1).items()}
    eng = FluxTransformer(FluxConfig(num_layers=1, num_single_layers=1, num_attention_heads=2, joint_attention_dim=256,
                                     pooled_projection_dim=64)).load_state_dict(P)
    s_txt, s_img = 77, 333
    g = torch.Generator().manual_seed(1)
    img_ids = torch.stack([torch.zeros(s_img), torch.randint(0, 40, (s_img,), generator=g).float(),
                           torch.randint(0, 40, (s_img,), generator=g).float()], -1)
    enc = (torch.randn(s_txt, 256, generator=g) * 0.3).to(torch.bfloat16)
    lat = torch.randn(s_img, 64, generator=g).to(torch.bfloat16).cuda()
    eng.prepare(torch.cat([torch.zeros(s_txt, 3), img_ids]), enc, None, s_txt=s_txt)
    v = eng.forward(lat, 0.5, 3.5)
    torch.cuda.synchronize()
    Pg = {k: w.cuda() for k, w in P.items()}
    ref = fd.flux_forward(Pg, ocfg, lat[None].float(), torch.tensor([0.5]).cuda(), torch.tensor([3.5]).cuda(),
                          torch.zeros(1, 64).cuda(), enc[None].float().cuda(), torch.zeros(s_txt, 3).cuda(), img_ids.cuda())[0]
    assert fs.psnr(v.float(), ref) >= PSNR_MIN_DB
    # zero steps is a no-op; zero-row GEMM / Euler calls succeed without launching
    before = lat.clone()
    eng.denoise_(lat, 100, [1.0], 3.5)
    torch.cuda.synchronize()
    assert torch.equal(lat, before)
    X = torch.zeros(4, 256, device="cuda", dtype=torch.bfloat16)
    Y = torch.ones(4, 256, device="cuda", dtype=torch.bfloat16)
    ops.gemm(X[:0], torch.zeros(256, 256, device="cuda", dtype=torch.bfloat16), out=Y[:0])     # M = 0: nothing launched
    assert (Y == 1).all()
    ops.euler_update_(lat, lat, 0, 0.1)
    torch.cuda.synchronize()
    assert torch.equal(lat, before)
