"""Oracle: latent packing, position ids, flow-match Euler schedule, LoRA merge and
the condition-token denoise loop of `PBRFluxPipeline.__call__`.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Follows
flux_piplines/texturing/pipeline.py:59-69 (calculate_shift), :240-275 (pack /
unpack / ids), :277-402 (condition latents + id offsets), :580-684 (loop).
The scheduler is diffusers' FlowMatchEulerDiscreteScheduler [ext] with the
FLUX.1-dev scheduler_config (base_shift 0.5, max_shift 1.15, 256/4096).
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import numpy as np
import torch

from .flux_dit import FluxConfig, flux_forward


def calculate_shift(image_seq_len, base_seq_len=256, max_seq_len=4096, base_shift=0.5, max_shift=1.15):
    """pipeline.py:59-69; called with the scheduler config values (:596-602)."""
    m = (max_shift - base_shift) / (max_seq_len - base_seq_len)
    b = base_shift - m * base_seq_len
    return image_seq_len * m + b


def flow_match_sigmas(num_steps: int, image_seq_len: int) -> np.ndarray:
    """sigmas = linspace(1, 1/N, N) (:594) -> time shift exp(mu)/(exp(mu)+(1/s-1)) [ext
    set_timesteps, use_dynamic_shifting] -> float32, with the terminal 0 appended."""
    sig = np.linspace(1.0, 1.0 / num_steps, num_steps).astype(np.float32)     # [ext] set_timesteps: np.array(sigmas).astype(np.float32)
    mu = calculate_shift(image_seq_len)
    sig = math.exp(mu) / (math.exp(mu) + (1 / sig - 1) ** 1.0)                # [ext] time_shift(mu, 1.0, sigmas): evaluated in fp32
    return np.concatenate([sig.astype(np.float32), np.zeros(1, np.float32)])


def pack_latents(lat: torch.Tensor) -> torch.Tensor:
    """:240-249 (pixel_shuffle=True): [B,C,H,W] -> [B,(H/2)(W/2),4C]."""
    B, C, H, W = lat.shape
    lat = lat.view(B, C, H // 2, 2, W // 2, 2).permute(0, 2, 4, 1, 3, 5)
    return lat.reshape(B, (H // 2) * (W // 2), C * 4)


def unpack_latents(lat: torch.Tensor, height: int, width: int, vae_scale_factor: int = 8) -> torch.Tensor:
    """:251-265."""
    B, _, ch = lat.shape
    h = 2 * (int(height) // (vae_scale_factor * 2))
    w = 2 * (int(width) // (vae_scale_factor * 2))
    lat = lat.view(B, h // 2, w // 2, ch // 4, 2, 2).permute(0, 3, 1, 4, 2, 5)
    return lat.reshape(B, ch // 4, h, w)


def latent_image_ids(height: int, width: int, offset_x=0, offset_y=0, offset_z=0, dtype=torch.float32):
    """:267-275: ids[...,1]=row+offset_y, ids[...,2]=col+offset_x, [H*W,3] in `dtype`
    (the reference builds them in bf16, :571 -- exact for integers <= 256)."""
    ids = torch.zeros(height, width, 3)
    ids[..., 1] += torch.arange(offset_y, offset_y + height)[:, None]
    ids[..., 2] += torch.arange(offset_x, offset_x + width)[None, :]
    if offset_z != 0:
        ids[..., 0] += offset_z
    return ids.reshape(height * width, 3).to(dtype)


def build_ids(HL: int, WL: int, control_hw=None, dual_hw=None, dtype=torch.float32):
    """Token order and id offsets of :292-393 + :580-582: noise (0,0); control
    (y + HL/2, x + 0); dual (y + HL/2, x + WL/2); condition = cat[control, dual]."""
    ids = [latent_image_ids(HL // 2, WL // 2, dtype=dtype)]
    if control_hw is not None:
        ids.append(latent_image_ids(control_hw[0] // 2, control_hw[1] // 2, offset_x=0, offset_y=HL // 2, dtype=dtype))
    if dual_hw is not None:
        ids.append(latent_image_ids(dual_hw[0] // 2, dual_hw[1] // 2, offset_x=WL // 2, offset_y=HL // 2, dtype=dtype))
    return torch.cat(ids, dim=0)


def euler_step(latents: torch.Tensor, v: torch.Tensor, sigma: float, sigma_next: float) -> torch.Tensor:
    """FlowMatchEulerDiscreteScheduler.step [ext]: `sample.float() + (sigma_next - sigma) * model_output`, cast to
    model_output.dtype.  The sigmas are 0-dim fp32 tensors there, so by torch's promotion rules the product
    `(sigma_next - sigma) * model_output` is formed in model_output's dtype: with the bf16 transformer of the reference the
    sigma difference is cast to bf16, multiplied, and the increment rounded to bf16 BEFORE the fp32 add.  Kept (0-dim tensors
    here too), so an fp32 `v` sees no extra rounding."""
    dt = torch.tensor(np.float32(sigma_next)) - torch.tensor(np.float32(sigma))
    x = latents.to(torch.float32) + dt.to(v.device) * v
    return x.to(v.dtype)


def merge_lora(P: Dict[str, torch.Tensor], lora: Dict[str, torch.Tensor], scale: float) -> Dict[str, torch.Tensor]:
    """peft merge rule for one active adapter (pipeline.py:108-118,245,263; A.2):
    W' = W + scale * B @ A per target (fp32 math, rounded to W's dtype); entries
    `<name>.weight` / `<name>.bias` in `lora` without lora_A/B are modules_to_save
    replacements (trainer.py:297-304: x_embedder)."""
    out = dict(P)
    names = {k[: -len(".lora_A.weight")] for k in lora if k.endswith(".lora_A.weight")}
    for n in sorted(names):
        A, B = lora[n + ".lora_A.weight"].float(), lora[n + ".lora_B.weight"].float()
        W = P[n + ".weight"]
        out[n + ".weight"] = (W.float() + scale * (B @ A)).to(W.dtype)
    for k, v in lora.items():
        if ".lora_" not in k:
            out[k] = v.to(P[k].dtype)
    return out


LORA_TARGETS_DOUBLE = ("attn.to_q", "attn.to_k", "attn.to_v", "attn.to_out.0", "attn.add_q_proj",
                       "attn.add_k_proj", "attn.add_v_proj", "attn.to_add_out", "ff.net.0.proj",
                       "ff.net.2", "ff_context.net.0.proj", "ff_context.net.2")   # trainer.py:283-296
LORA_TARGETS_SINGLE = ("attn.to_q", "attn.to_k", "attn.to_v")


def init_lora(P: Dict[str, torch.Tensor], cfg: FluxConfig, rank: int, seed: int, std: float = 0.02, device="cpu"):
    """Random LoRA adapter over the reference's target list (+ x_embedder replacement)."""
    g = torch.Generator(device=device).manual_seed(seed)
    L = {}
    names = [f"transformer_blocks.{i}.{t}" for i in range(cfg.num_layers) for t in LORA_TARGETS_DOUBLE]
    names += [f"single_transformer_blocks.{i}.{t}" for i in range(cfg.num_single_layers) for t in LORA_TARGETS_SINGLE]
    for n in names:
        o, i = P[n + ".weight"].shape
        L[n + ".lora_A.weight"] = torch.randn(rank, i, generator=g, device=device) * std
        L[n + ".lora_B.weight"] = torch.randn(o, rank, generator=g, device=device) * std
    L["x_embedder.weight"] = torch.randn(P["x_embedder.weight"].shape, generator=g, device=device) * std
    L["x_embedder.bias"] = torch.randn(P["x_embedder.bias"].shape, generator=g, device=device) * std
    return L


@torch.no_grad()
def denoise(P, cfg: FluxConfig, noise: torch.Tensor, condition: Optional[torch.Tensor], img_ids: torch.Tensor,
            num_steps: int = 28, guidance_scale: float = 3.5, S_txt: int = 512,
            enc: Optional[torch.Tensor] = None, pooled: Optional[torch.Tensor] = None, trace=None):
    """The hot loop :630-684.  noise [B,S_noise,64], condition [B,S_cond,64] (clean,
    re-imposed every step :644-645).  Returns denoised noise tokens [B,S_noise,64]."""
    dt = P["x_embedder.weight"].dtype
    B, S_noise, _ = noise.shape
    if enc is None:                                   # :538-543 zero text embeddings
        enc = torch.zeros(B, S_txt, cfg.joint_attention_dim, dtype=dt, device=noise.device)
    if pooled is None:
        pooled = torch.zeros(B, cfg.pooled_projection_dim, dtype=dt, device=noise.device)
    txt_ids = torch.zeros(enc.shape[1], 3, device=noise.device)
    sig = flow_match_sigmas(num_steps, S_noise)                      # mu from noise tokens only (:595)
    guidance = torch.full([B], guidance_scale, dtype=torch.float32, device=noise.device)
    latents = noise.to(dt) if condition is None else torch.cat([noise.to(dt), condition.to(dt)], dim=1)
    for i in range(num_steps):
        if condition is not None:
            latents = torch.cat([latents[:, :S_noise], condition.to(dt)], dim=1)
        # :643 t.to(latents.dtype) and :648 timestep / 1000 -- both in bf16 in the reference whatever `dt` is here
        t = torch.full([B], float(np.float32(sig[i]) * np.float32(1000.0)), dtype=torch.float32, device=noise.device).to(torch.bfloat16)
        v = flux_forward(P, cfg, latents, t / 1000, guidance, pooled, enc, txt_ids, img_ids)
        latents = euler_step(latents, v, sig[i], sig[i + 1])
        if trace is not None:
            trace.append(latents[:, :S_noise].clone())
    return latents[:, :S_noise]


def psnr(x: torch.Tensor, ref: torch.Tensor) -> float:
    """PSNR in dB with peak = dynamic range of the reference tensor."""
    x, ref = x.double(), ref.double()
    mse = torch.mean((x - ref) ** 2).item()
    peak = (ref.max() - ref.min()).item()
    if mse == 0:
        return float("inf")
    return 10.0 * math.log10(peak * peak / mse)


@torch.no_grad()
def pipeline_call(P, cfg: FluxConfig, VP, vcfg, control_image, dual_image, height: int, width: int, num_steps: int,
                  generator: torch.Generator, guidance_scale: float = 3.5, S_txt: int = 512, output_type: str = "latent"):
    """`PBRFluxPipeline.__call__` :502-700 from PIL images, with the VAE on both ends (pinned against the reference's own
    `__call__` by tests/golden/ref_flux_call.npz).  Order of generator draws: noise (:292), dual sample (:308), control
    sample (:364).  `control_image` / `dual_image`: PIL (sizes multiples of 16) or None.  P / VP in the run dtype (the
    reference runs everything in bf16: pipeline.py:102)."""
    from . import vae as ov
    dt = P["x_embedder.weight"].dtype

    def prep(img):                                   # VaeImageProcessor.preprocess [ext]: uint8 -> [-1, 1], no resize (:301, :357)
        a = torch.from_numpy(np.asarray(img.convert("RGB")).astype(np.float32) / 255.0)
        return (a.permute(2, 0, 1)[None] * 2.0 - 1.0).to(dt)

    def encode(img):                                 # _encode_vae_image :226-238
        mean, logvar = ov.encode_moments(VP, vcfg, prep(img))
        z = ov.sample(mean, logvar, torch.randn(mean.shape, generator=generator, dtype=mean.dtype))
        return ((z - vcfg.shift_factor) * vcfg.scaling_factor).to(dt)

    HL, WL = 2 * (height // 16), 2 * (width // 16)
    noise = pack_latents(torch.randn((1, 16, HL, WL), generator=generator, dtype=dt))
    dual = encode(dual_image) if dual_image is not None else None
    control = encode(control_image) if control_image is not None else None
    cond = [pack_latents(z) for z in (control, dual) if z is not None]                       # condition = cat[control, dual] :575-577
    ids = build_ids(HL, WL, tuple(control.shape[-2:]) if control is not None else None,
                    tuple(dual.shape[-2:]) if dual is not None else None, dtype=dt)
    lat = denoise(P, cfg, noise, torch.cat(cond, 1) if cond else None, ids, num_steps, guidance_scale, S_txt)
    if output_type == "latent":
        return lat
    z = unpack_latents(lat, height, width) / vcfg.scaling_factor + vcfg.shift_factor         # :688-689
    img = ov.decode(VP, vcfg, z.to(dt))
    img = (img.float() / 2 + 0.5).clamp(0, 1)                                                # VaeImageProcessor.postprocess [ext]
    return (img.permute(0, 2, 3, 1).numpy() * 255).round().astype("uint8")
