"""FLUX AutoencoderKL on B200: host orchestration over libunitex_b200.so (im2col + tcgen05 GEMM convolutions, GroupNorm+SiLU,
single-head mid-block attention as two GEMMs around an fp32 row softmax).  Replaces `self.vae.encode / self.vae.decode`
of the reference sampler (flux_piplines/texturing/pipeline.py:226-238, :688-692; diffusers AutoencoderKL [ext], FLUX config:
latent 16, blocks (128,256,512,512), 2 layers/block, GN32, scaling 0.3611, shift 0.1159, no quant convs).

Layout: activations NHWC bf16 ([N*H*W, C] matrices), conv weights re-arranged once to [Cout, ky, kx, Cin] (K padded to a
multiple of 64, Cout to a multiple of 8).  torch is used for allocation and the final NCHW views only.
"""
from __future__ import annotations

import json
import os
from typing import Dict, Optional

import torch

from . import _lib, ops
from .ops import _p, _stream


def _ceil(v, m):
    return (v + m - 1) // m * m


class AutoencoderKLB200:
    def __init__(self, state_dict: Dict[str, torch.Tensor], block_out_channels=(128, 256, 512, 512), layers_per_block=2,
                 latent_channels=16, in_channels=3, norm_num_groups=32, scaling_factor=0.3611, shift_factor=0.1159,
                 device="cuda"):
        self.device = torch.device(device)
        self.boc = tuple(block_out_channels)
        self.layers = layers_per_block
        self.latent_channels, self.in_channels, self.groups = latent_channels, in_channels, norm_num_groups
        self.scaling_factor, self.shift_factor = scaling_factor, shift_factor
        self.dtype = torch.bfloat16
        self.lib = _lib.load()
        self.implicit_conv = os.environ.get("UTX_VAE_IM2COL", "0") != "1"   # debug knob: force the im2col path
        self.W: Dict[str, torch.Tensor] = {}
        for k, v in state_dict.items():
            if not k.endswith(".weight"):
                continue
            n = k[:-len(".weight")]
            b = state_dict[n + ".bias"]
            if v.dim() == 4:                                   # conv: [Cout,Cin,k,k] -> [Cout_pad, K_pad] with K = (ky,kx,cin)
                co, ci, kh, kw = v.shape
                w = v.permute(0, 2, 3, 1).reshape(co, kh * kw * ci)
                wp = torch.zeros(_ceil(co, 8), _ceil(w.shape[1], 64))
                wp[:co, :w.shape[1]] = w
                bp = torch.zeros(_ceil(co, 8))
                bp[:co] = b
                self.W[n + ".w"], self.W[n + ".b"] = wp.to(self.device, torch.bfloat16), bp.to(self.device, torch.bfloat16)
            elif v.dim() == 2:                                 # attention linears
                self.W[n + ".w"], self.W[n + ".b"] = v.to(self.device, torch.bfloat16).contiguous(), b.to(self.device, torch.bfloat16)
            else:                                              # GroupNorm affine (kept fp32)
                self.W[n + ".w"], self.W[n + ".b"] = v.to(self.device, torch.float32), b.to(self.device, torch.float32)
        self._ones: Dict[int, torch.Tensor] = {}
        self._stats = torch.zeros(64 * 2 * 8, device=self.device, dtype=torch.float64)

    # ------------------------------------------------------------------ construction
    @classmethod
    def from_pretrained(cls, vae_dir: str, device="cuda"):
        from safetensors.torch import load_file
        cfg = json.load(open(os.path.join(vae_dir, "config.json")))
        sd = {}
        for f in sorted(os.listdir(vae_dir)):
            if f.endswith(".safetensors"):
                sd.update(load_file(os.path.join(vae_dir, f)))
        return cls(sd, cfg["block_out_channels"], cfg["layers_per_block"], cfg["latent_channels"], cfg["in_channels"],
                   cfg["norm_num_groups"], cfg.get("scaling_factor", 0.3611), cfg.get("shift_factor", 0.1159), device)

    @classmethod
    def from_random(cls, seed=0, device="cuda", block_out_channels=(128, 256, 512, 512)):
        """Random-init FLUX-VAE-shaped weights (bench path)."""
        g = torch.Generator().manual_seed(seed)
        sd = {}
        boc = tuple(block_out_channels)

        def conv(n, co, ci, k=3):
            sd[n + ".weight"] = torch.randn(co, ci, k, k, generator=g) / (ci * k * k) ** 0.5
            sd[n + ".bias"] = torch.zeros(co)

        def gn(n, c):
            sd[n + ".weight"], sd[n + ".bias"] = torch.ones(c), torch.zeros(c)

        def res(p, ci, co):
            gn(p + "norm1", ci); conv(p + "conv1", co, ci); gn(p + "norm2", co); conv(p + "conv2", co, co)
            if ci != co:
                conv(p + "conv_shortcut", co, ci, 1)

        def mid(p, c):
            res(p + "resnets.0.", c, c); gn(p + "attentions.0.group_norm", c)
            for n in ("to_q", "to_k", "to_v", "to_out.0"):
                sd[p + f"attentions.0.{n}.weight"] = torch.randn(c, c, generator=g) / c ** 0.5
                sd[p + f"attentions.0.{n}.bias"] = torch.zeros(c)
            res(p + "resnets.1.", c, c)

        conv("encoder.conv_in", boc[0], 3)
        ci = boc[0]
        for i, c in enumerate(boc):
            for j in range(2):
                res(f"encoder.down_blocks.{i}.resnets.{j}.", ci, c); ci = c
            if i < len(boc) - 1:
                conv(f"encoder.down_blocks.{i}.downsamplers.0.conv", c, c)
        mid("encoder.mid_block.", boc[-1]); gn("encoder.conv_norm_out", boc[-1]); conv("encoder.conv_out", 32, boc[-1])
        rb = boc[::-1]
        conv("decoder.conv_in", rb[0], 16); mid("decoder.mid_block.", rb[0])
        ci = rb[0]
        for i, c in enumerate(rb):
            for j in range(3):
                res(f"decoder.up_blocks.{i}.resnets.{j}.", ci, c); ci = c
            if i < len(rb) - 1:
                conv(f"decoder.up_blocks.{i}.upsamplers.0.conv", c, c)
        gn("decoder.conv_norm_out", rb[-1]); conv("decoder.conv_out", 3, rb[-1])
        return cls(sd, boc, device=device)

    # ------------------------------------------------------------------ building blocks
    def _conv3(self, name, x, N, H, W, C, up=1, stride=1, pad=1, res=None):
        """x [N*H*W, C] NHWC bf16 -> (y [N*Ho*Wo, Cout_pad], Ho, Wo).  `res`: fused residual add in the GEMM epilogue."""
        w, b = self.W[name + ".w"], self.W[name + ".b"]
        Hs, Ws = H * up, W * up
        if stride == 1:
            Ho, Wo = Hs, Ws
        else:                                                   # Downsample2D: pad (0,1,0,1), stride 2
            Ho, Wo = (Hs + 1 - 3) // 2 + 1, (Ws + 1 - 3) // 2 + 1
        Kp = w.shape[1]
        if (self.implicit_conv and up == 2 and stride == 1 and pad == 1 and C % 64 == 0 and Kp == 9 * C and
                (Ws % 128 == 0 or (128 % Ws == 0 and Hs % (128 // Ws) == 0))):
            # Upsample2D: materialise the nearest-neighbour upsampling (4x the input, not the 36x of an im2col buffer)
            xu = torch.empty(N * Hs * Ws, C, device=self.device, dtype=torch.bfloat16)
            _lib.check(self.lib.utx_upsample2x_nhwc(_p(x), N, H, W, C, _p(xu), _stream()), "utx_upsample2x_nhwc")
            x, H, W, up = xu, Hs, Ws, 1
        if (self.implicit_conv and up == 1 and stride == 1 and pad == 1 and C % 64 == 0 and Kp == 9 * C and
                (W % 128 == 0 or (128 % W == 0 and H % (128 // W) == 0))):
            # implicit GEMM: TMA boxes of the activation shifted by the tap are the A tiles, no im2col buffer
            Co = w.shape[0]
            y = res if res is not None else torch.empty(N * H * W, Co, device=self.device, dtype=torch.bfloat16)
            _lib.check(self.lib.utx_conv3x3_nhwc(_p(x), N, H, W, C, _p(w), _p(b), Co, _p(y), y.stride(0),
                                                 _p(self._one(Co)) if res is not None else None, _p(res), 0 if res is None else res.stride(0),
                                                 _stream()), "utx_conv3x3_nhwc")
            return y, H, W
        col = torch.empty(N * Ho * Wo, Kp, device=self.device, dtype=torch.bfloat16)
        _lib.check(self.lib.utx_im2col3x3(_p(x), N, H, W, C, up, stride, pad, Ho, Wo, Kp, _p(col), _stream()), "utx_im2col3x3")
        if res is None:
            y = ops.gemm(col, w, b)
        else:
            y = res                                             # callers pass a buffer they own: updated in place
            ops.gemm(col, w, b, epi=ops.EPI_GATE_RES, gate=self._one(w.shape[0]), res=y, out=y)
        return y, Ho, Wo

    def _lin(self, name, x, res=None):
        w, b = self.W[name + ".w"], self.W[name + ".b"]
        if res is None:
            return ops.gemm(x, w, b)
        ops.gemm(x, w, b, epi=ops.EPI_GATE_RES, gate=self._one(w.shape[0]), res=res, out=res)
        return res

    def _one(self, n):
        if n not in self._ones:
            self._ones[n] = torch.ones(n, device=self.device, dtype=torch.float32)
        return self._ones[n]

    def _gn(self, name, x, N, HW, C, silu=True):
        y = torch.empty_like(x)
        need = (self.lib.utx_groupnorm_workspace_bytes(N, HW, C, self.groups) + 7) // 8     # in doubles
        if self._stats.numel() < need:
            self._stats = torch.zeros(need, device=self.device, dtype=torch.float64)
        _lib.check(self.lib.utx_groupnorm_nhwc(_p(x), _p(y), N, HW, C, self.groups, _p(self.W[name + ".w"]), _p(self.W[name + ".b"]),
                                               int(silu), _p(self._stats), _stream()), "utx_groupnorm_nhwc")
        return y

    def _resnet(self, p, x, N, H, W, Cin):
        Cout = self.W[p + "conv1.w"].shape[0]
        h = self._gn(p + "norm1", x, N, H * W, Cin)
        h, _, _ = self._conv3(p + "conv1", h, N, H, W, Cin)
        h = self._gn(p + "norm2", h, N, H * W, Cout)
        if (p + "conv_shortcut.w") in self.W:                   # 1x1 conv == Linear over channels
            sc = ops.gemm(x, self.W[p + "conv_shortcut.w"], self.W[p + "conv_shortcut.b"])
        else:
            sc = x.clone()
        y, _, _ = self._conv3(p + "conv2", h, N, H, W, Cout, res=sc)
        return y, Cout

    def _attn(self, p, x, N, HW, C):
        outs = []
        for n in range(N):
            xs = x[n * HW:(n + 1) * HW]
            h = self._gn(p + "group_norm", xs, 1, HW, C, silu=False)
            q, k, v = (self._lin(p + t, h) for t in ("to_q", "to_k", "to_v"))
            S = torch.empty(HW, HW, device=self.device, dtype=torch.float32)
            _lib.check(self.lib.utx_gemm_bf16_f32out(_p(q), C, _p(k), C, None, _p(S), HW, HW, HW, C, float(C) ** -0.5, _stream()),
                       "utx_gemm_bf16_f32out")
            P = torch.empty(HW, HW, device=self.device, dtype=torch.bfloat16)
            _lib.check(self.lib.utx_softmax_rows(_p(S), HW, _p(P), HW, HW, HW, _stream()), "utx_softmax_rows")
            vt = torch.empty(C, HW, device=self.device, dtype=torch.bfloat16)
            _lib.check(self.lib.utx_transpose_bf16(_p(v), C, _p(vt), HW, HW, C, _stream()), "utx_transpose_bf16")
            o = ops.gemm(P, vt, None)
            outs.append(self._lin(p + "to_out.0", o, res=xs.clone()))
        return torch.cat(outs, 0) if N > 1 else outs[0]

    def _mid(self, p, x, N, H, W, C):
        x, _ = self._resnet(p + "resnets.0.", x, N, H, W, C)
        x = self._attn(p + "attentions.0.", x, N, H * W, C)
        x, _ = self._resnet(p + "resnets.1.", x, N, H, W, C)
        return x

    # ------------------------------------------------------------------ public
    @torch.no_grad()
    def decode(self, z: torch.Tensor) -> torch.Tensor:
        """z [N,16,h,w] (already / scaling + shift, :689) -> image [N,3,8h,8w] bf16."""
        N, Cz, H, W = z.shape
        assert H * W % 8 == 0, "latent area must be a multiple of 8"
        x = z.to(self.device, torch.bfloat16).permute(0, 2, 3, 1).reshape(N * H * W, Cz).contiguous()
        x, _, _ = self._conv3("decoder.conv_in", x, N, H, W, Cz)
        C = self.boc[-1]
        x = self._mid("decoder.mid_block.", x, N, H, W, C)
        nb = len(self.boc)
        for i in range(nb):
            for j in range(self.layers + 1):
                x, C = self._resnet(f"decoder.up_blocks.{i}.resnets.{j}.", x, N, H, W, C)
            if i < nb - 1:
                x, H, W = self._conv3(f"decoder.up_blocks.{i}.upsamplers.0.conv", x, N, H, W, C, up=2)
        x = self._gn("decoder.conv_norm_out", x, N, H * W, C)
        y, _, _ = self._conv3("decoder.conv_out", x, N, H, W, C)
        return y[:, :self.in_channels].reshape(N, H, W, self.in_channels).permute(0, 3, 1, 2).contiguous()

    @torch.no_grad()
    def encode_moments(self, img: torch.Tensor):
        """img [N,3,H,W] in [-1,1] -> (mean, logvar) [N,16,H/8,W/8] fp32."""
        N, Ci, H, W = img.shape
        x = img.to(self.device, torch.bfloat16).permute(0, 2, 3, 1).reshape(N * H * W, Ci).contiguous()
        x, _, _ = self._conv3("encoder.conv_in", x, N, H, W, Ci)
        C = self.boc[0]
        nb = len(self.boc)
        for i in range(nb):
            for j in range(self.layers):
                x, C = self._resnet(f"encoder.down_blocks.{i}.resnets.{j}.", x, N, H, W, C)
            if i < nb - 1:
                x, H, W = self._conv3(f"encoder.down_blocks.{i}.downsamplers.0.conv", x, N, H, W, C, stride=2, pad=0)
        x = self._mid("encoder.mid_block.", x, N, H, W, C)
        x = self._gn("encoder.conv_norm_out", x, N, H * W, C)
        m, _, _ = self._conv3("encoder.conv_out", x, N, H, W, C)
        m = m[:, :2 * self.latent_channels].float().reshape(N, H, W, 2 * self.latent_channels).permute(0, 3, 1, 2)
        mean, logvar = m.chunk(2, dim=1)
        return mean.contiguous(), logvar.clamp(-30.0, 20.0).contiguous()

    @torch.no_grad()
    def encode_sample(self, img: torch.Tensor, generator: Optional[torch.Generator] = None) -> torch.Tensor:
        """latent_dist.sample(generator) (:234): mean + exp(0.5 logvar) * randn, noise drawn like diffusers' randn_tensor."""
        mean, logvar = self.encode_moments(img)
        gdev = generator.device if generator is not None else self.device
        noise = torch.randn(mean.shape, generator=generator, device=gdev, dtype=torch.bfloat16).to(self.device).float()
        return (mean + torch.exp(0.5 * logvar) * noise).to(torch.bfloat16)
