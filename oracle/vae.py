"""Oracle: eager PyTorch restatement of the FLUX AutoencoderKL (TEST INFRASTRUCTURE, see oracle/__init__.py).

PARITY UNPINNED: the arithmetic is diffusers' `AutoencoderKL` [ext, absent]; restated from the published FLUX.1-dev VAE
config (latent_channels 16, block_out_channels (128,256,512,512), layers_per_block 2, norm_num_groups 32, act silu,
scaling_factor 0.3611, shift_factor 0.1159, no quant convs).  Reference call sites:
flux_piplines/texturing/pipeline.py:226-238 (`_encode_vae_image`: latent_dist.sample(generator), (z - shift) * scale)
and :688-692 (decode of latents / scale + shift).  Parameter names follow the diffusers state dict.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Tuple

import torch
import torch.nn.functional as F


@dataclass(frozen=True)
class VaeConfig:
    in_channels: int = 3
    latent_channels: int = 16
    block_out_channels: Tuple[int, ...] = (128, 256, 512, 512)
    layers_per_block: int = 2
    norm_num_groups: int = 32
    scaling_factor: float = 0.3611
    shift_factor: float = 0.1159

    @staticmethod
    def tiny() -> "VaeConfig":
        return VaeConfig(block_out_channels=(64, 64, 128, 128))


def _resnet_names(prefix, cin, cout):
    out = [(prefix + "norm1", "gn", cin), (prefix + "conv1", "conv3", (cout, cin)), (prefix + "norm2", "gn", cout),
           (prefix + "conv2", "conv3", (cout, cout))]
    if cin != cout:
        out.append((prefix + "conv_shortcut", "conv1", (cout, cin)))
    return out


def _mid_names(prefix, c):
    out = _resnet_names(prefix + "resnets.0.", c, c)
    out += [(prefix + "attentions.0.group_norm", "gn", c)]
    out += [(prefix + f"attentions.0.{n}", "lin", (c, c)) for n in ("to_q", "to_k", "to_v", "to_out.0")]
    return out + _resnet_names(prefix + "resnets.1.", c, c)


def param_specs(cfg: VaeConfig):
    boc = cfg.block_out_channels
    S = [("encoder.conv_in", "conv3", (boc[0], cfg.in_channels))]
    cin = boc[0]
    for i, c in enumerate(boc):
        for j in range(cfg.layers_per_block):
            S += _resnet_names(f"encoder.down_blocks.{i}.resnets.{j}.", cin, c)
            cin = c
        if i < len(boc) - 1:
            S.append((f"encoder.down_blocks.{i}.downsamplers.0.conv", "conv3", (c, c)))
    S += _mid_names("encoder.mid_block.", boc[-1])
    S += [("encoder.conv_norm_out", "gn", boc[-1]), ("encoder.conv_out", "conv3", (2 * cfg.latent_channels, boc[-1]))]
    rb = list(reversed(boc))
    S.append(("decoder.conv_in", "conv3", (rb[0], cfg.latent_channels)))
    S += _mid_names("decoder.mid_block.", rb[0])
    cin = rb[0]
    for i, c in enumerate(rb):
        for j in range(cfg.layers_per_block + 1):
            S += _resnet_names(f"decoder.up_blocks.{i}.resnets.{j}.", cin, c)
            cin = c
        if i < len(rb) - 1:
            S.append((f"decoder.up_blocks.{i}.upsamplers.0.conv", "conv3", (c, c)))
    S += [("decoder.conv_norm_out", "gn", rb[-1]), ("decoder.conv_out", "conv3", (cfg.in_channels, rb[-1]))]
    return S


def init_params(cfg: VaeConfig, seed: int = 0, dtype=torch.float32) -> Dict[str, torch.Tensor]:
    g = torch.Generator().manual_seed(seed)
    P = {}
    for name, kind, shp in param_specs(cfg):
        if kind == "gn":
            P[name + ".weight"] = (1 + 0.1 * torch.randn(shp, generator=g)).to(dtype)
            P[name + ".bias"] = (0.1 * torch.randn(shp, generator=g)).to(dtype)
        else:
            k = {"conv3": 3, "conv1": 1}.get(kind)
            o, i = shp
            fan = i * (k * k if k else 1)
            w = torch.randn((o, i, k, k) if k else (o, i), generator=g) / fan ** 0.5
            P[name + ".weight"] = w.to(dtype)
            P[name + ".bias"] = (0.05 * torch.randn(o, generator=g)).to(dtype)
    return P


def _gn(P, n, x, groups):
    return F.group_norm(x, groups, P[n + ".weight"], P[n + ".bias"], eps=1e-6)


def _resnet(P, p, x, groups):
    h = F.conv2d(F.silu(_gn(P, p + "norm1", x, groups)), P[p + "conv1.weight"], P[p + "conv1.bias"], padding=1)
    h = F.conv2d(F.silu(_gn(P, p + "norm2", h, groups)), P[p + "conv2.weight"], P[p + "conv2.bias"], padding=1)
    if (p + "conv_shortcut.weight") in P:
        x = F.conv2d(x, P[p + "conv_shortcut.weight"], P[p + "conv_shortcut.bias"])
    return x + h


def _attn(P, p, x, groups):
    B, Cc, H, W = x.shape
    h = _gn(P, p + "group_norm", x, groups).view(B, Cc, H * W).transpose(1, 2)
    q, k, v = (F.linear(h, P[p + n + ".weight"], P[p + n + ".bias"]) for n in ("to_q", "to_k", "to_v"))
    o = F.scaled_dot_product_attention(q[:, None], k[:, None], v[:, None])[:, 0]
    o = F.linear(o, P[p + "to_out.0.weight"], P[p + "to_out.0.bias"])
    return x + o.transpose(1, 2).reshape(B, Cc, H, W)


def _mid(P, p, x, groups):
    x = _resnet(P, p + "resnets.0.", x, groups)
    x = _attn(P, p + "attentions.0.", x, groups)
    return _resnet(P, p + "resnets.1.", x, groups)


@torch.no_grad()
def decode(P, cfg: VaeConfig, z: torch.Tensor) -> torch.Tensor:
    """AutoencoderKL.decode [ext]: z [B,16,h,w] (already / scale + shift) -> image [B,3,8h,8w]."""
    g = cfg.norm_num_groups
    x = F.conv2d(z, P["decoder.conv_in.weight"], P["decoder.conv_in.bias"], padding=1)
    x = _mid(P, "decoder.mid_block.", x, g)
    n = len(cfg.block_out_channels)
    for i in range(n):
        for j in range(cfg.layers_per_block + 1):
            x = _resnet(P, f"decoder.up_blocks.{i}.resnets.{j}.", x, g)
        if i < n - 1:
            x = F.interpolate(x, scale_factor=2.0, mode="nearest")
            x = F.conv2d(x, P[f"decoder.up_blocks.{i}.upsamplers.0.conv.weight"], P[f"decoder.up_blocks.{i}.upsamplers.0.conv.bias"], padding=1)
    x = F.silu(_gn(P, "decoder.conv_norm_out", x, g))
    return F.conv2d(x, P["decoder.conv_out.weight"], P["decoder.conv_out.bias"], padding=1)


@torch.no_grad()
def encode_moments(P, cfg: VaeConfig, img: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """AutoencoderKL.encode [ext] -> (mean, logvar clamped to [-30, 20])."""
    g = cfg.norm_num_groups
    x = F.conv2d(img, P["encoder.conv_in.weight"], P["encoder.conv_in.bias"], padding=1)
    n = len(cfg.block_out_channels)
    for i in range(n):
        for j in range(cfg.layers_per_block):
            x = _resnet(P, f"encoder.down_blocks.{i}.resnets.{j}.", x, g)
        if i < n - 1:
            x = F.pad(x, (0, 1, 0, 1))                                  # Downsample2D(padding=0): asymmetric pad, stride 2
            x = F.conv2d(x, P[f"encoder.down_blocks.{i}.downsamplers.0.conv.weight"], P[f"encoder.down_blocks.{i}.downsamplers.0.conv.bias"], stride=2)
    x = _mid(P, "encoder.mid_block.", x, g)
    x = F.silu(_gn(P, "encoder.conv_norm_out", x, g))
    m = F.conv2d(x, P["encoder.conv_out.weight"], P["encoder.conv_out.bias"], padding=1)
    mean, logvar = m.chunk(2, dim=1)
    return mean, logvar.clamp(-30.0, 20.0)


def sample(mean, logvar, noise):
    """DiagonalGaussianDistribution.sample [ext]: mean + exp(0.5 logvar) * randn."""
    return mean + torch.exp(0.5 * logvar) * noise
