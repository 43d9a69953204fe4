"""FLUX AutoencoderKL on B200: weight packing + two calls into libunitex_b200.so (`utx_vae_encode` / `utx_vae_decode`, the C++
host loop of csrc/vae_engine.cu over implicit-GEMM tcgen05 convolutions, GroupNorm+SiLU and the chunked single-head mid-block
attention).  Replaces `self.vae.encode / self.vae.decode` of the reference sampler (flux_piplines/texturing/pipeline.py:226-238,
:688-692; diffusers AutoencoderKL [ext], FLUX config: latent 16, blocks (128,256,512,512), 2 layers/block, GN32, scaling
0.3611, shift 0.1159, no quant convs).

Conv weights are re-arranged once to [Cout, ky, kx, Cin] (K padded to a multiple of 64, Cout to a multiple of 8); torch is
used for allocation only.
"""
from __future__ import annotations

import ctypes as C
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
        self._ws = None
        self._bind()

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

    # ------------------------------------------------------------------ the C engine (csrc/vae_engine.cu)
    def _bind(self):
        """Hands the packed weights to `utx_vae_*` (include/unitex_b200.h): pointer tables only, the tensors stay owned here."""
        W, L = self.W, self.lib
        nb = len(self.boc)
        cfg = _lib.VaeConfigC(self.in_channels, self.latent_channels, nb, (C.c_int * 8)(*self.boc, *([0] * (8 - nb))), self.layers,
                              self.groups)
        self._handle = _lib.vp()
        _lib.check(L.utx_vae_create(C.byref(cfg), C.byref(self._handle)), "utx_vae_create")
        p = lambda k: W[k].data_ptr() if k in W else None

        def resnet(prefix, cin, cout):
            r = _lib.VaeResnetC()
            for f, k in (("gn1_w", "norm1.w"), ("gn1_b", "norm1.b"), ("conv1_w", "conv1.w"), ("conv1_b", "conv1.b"),
                         ("gn2_w", "norm2.w"), ("gn2_b", "norm2.b"), ("conv2_w", "conv2.w"), ("conv2_b", "conv2.b"),
                         ("short_w", "conv_shortcut.w"), ("short_b", "conv_shortcut.b")):
                setattr(r, f, p(prefix + k))
            r.cin, r.cout = cin, cout
            assert (cin != cout) == ((prefix + "conv_shortcut.w") in W), prefix
            return r

        def mid(prefix, c):
            m = _lib.VaeMidC()
            m.res0, m.res1 = resnet(prefix + "resnets.0.", c, c), resnet(prefix + "resnets.1.", c, c)
            a = prefix + "attentions.0."
            for f, k in (("gn_w", "group_norm.w"), ("gn_b", "group_norm.b"), ("wq", "to_q.w"), ("bq", "to_q.b"), ("wk", "to_k.w"),
                         ("bk", "to_k.b"), ("wv", "to_v.w"), ("bv", "to_v.b"), ("wo", "to_out.0.w"), ("bo", "to_out.0.b")):
                setattr(m.attn, f, p(a + k))
            return m

        w = _lib.VaeWeightsC()
        enc, cin = [], self.boc[0]
        for i, c in enumerate(self.boc):
            for j in range(self.layers):
                enc.append(resnet(f"encoder.down_blocks.{i}.resnets.{j}.", cin, c))
                cin = c
        rb = self.boc[::-1]
        dec, cin = [], rb[0]
        for i, c in enumerate(rb):
            for j in range(self.layers + 1):
                dec.append(resnet(f"decoder.up_blocks.{i}.resnets.{j}.", cin, c))
                cin = c
        enc_arr, dec_arr = (_lib.VaeResnetC * len(enc))(*enc), (_lib.VaeResnetC * len(dec))(*dec)
        n1 = max(nb - 1, 1)
        tabs = [(_lib.vp * n1)(*[p(f"encoder.down_blocks.{i}.downsamplers.0.conv.{s}") for i in range(nb - 1)]) for s in ("w", "b")]
        tabs += [(_lib.vp * n1)(*[p(f"decoder.up_blocks.{i}.upsamplers.0.conv.{s}") for i in range(nb - 1)]) for s in ("w", "b")]
        w.enc_conv_in_w, w.enc_conv_in_b = p("encoder.conv_in.w"), p("encoder.conv_in.b")
        w.enc_res, w.enc_down_w, w.enc_down_b = enc_arr, tabs[0], tabs[1]
        w.enc_mid = mid("encoder.mid_block.", self.boc[-1])
        w.enc_norm_w, w.enc_norm_b = p("encoder.conv_norm_out.w"), p("encoder.conv_norm_out.b")
        w.enc_conv_out_w, w.enc_conv_out_b = p("encoder.conv_out.w"), p("encoder.conv_out.b")
        w.dec_conv_in_w, w.dec_conv_in_b = p("decoder.conv_in.w"), p("decoder.conv_in.b")
        w.dec_mid = mid("decoder.mid_block.", rb[0])
        w.dec_res, w.dec_up_w, w.dec_up_b = dec_arr, tabs[2], tabs[3]
        w.dec_norm_w, w.dec_norm_b = p("decoder.conv_norm_out.w"), p("decoder.conv_norm_out.b")
        w.dec_conv_out_w, w.dec_conv_out_b = p("decoder.conv_out.w"), p("decoder.conv_out.b")
        self._keep = (enc_arr, dec_arr, tabs, w)
        _lib.check(L.utx_vae_set_weights(self._handle, C.byref(w)), "utx_vae_set_weights")

    def __del__(self):
        try:
            if getattr(self, "_handle", None):
                self.lib.utx_vae_destroy(self._handle)
        except Exception:
            pass

    def _workspace(self, N, H, W, decode):
        n = self.lib.utx_vae_workspace_bytes(self._handle, N, H, W, int(decode))
        if n == 0:
            raise _lib.UtxError(f"utx_vae_workspace_bytes: {self.lib.utx_last_error().decode()}")
        if self._ws is None or self._ws.numel() < n:
            self._ws = None
            self._ws = torch.empty(n, device=self.device, dtype=torch.uint8)
        return self._ws

    # ------------------------------------------------------------------ public
    @torch.no_grad()
    def decode(self, z: torch.Tensor) -> torch.Tensor:
        """z [N,16,h,w] (already / scaling + shift, :689) -> image [N,3,8h,8w] bf16: ONE call into the C ABI."""
        N, Cz, H, W = z.shape
        z = z.to(self.device, torch.bfloat16).contiguous()
        f = 2 ** (len(self.boc) - 1)
        img = torch.empty(N, self.in_channels, H * f, W * f, device=self.device, dtype=torch.bfloat16)
        ws = self._workspace(N, H, W, True)
        _lib.check(self.lib.utx_vae_decode(self._handle, _p(z), N, H, W, _p(img), _p(ws), ws.numel(), _stream()), "utx_vae_decode")
        return img

    @torch.no_grad()
    def encode_moments(self, img: torch.Tensor):
        """img [N,3,H,W] in [-1,1] -> (mean, logvar) [N,16,H/8,W/8] fp32 (logvar clamped to [-30, 20]): ONE call into the C ABI."""
        N, Ci, H, W = img.shape
        img = img.to(self.device, torch.bfloat16).contiguous()
        f = 2 ** (len(self.boc) - 1)
        mom = torch.empty(N, 2 * self.latent_channels, H // f, W // f, device=self.device, dtype=torch.float32)
        ws = self._workspace(N, H, W, False)
        _lib.check(self.lib.utx_vae_encode(self._handle, _p(img), N, H, W, _p(mom), _p(ws), ws.numel(), _stream()), "utx_vae_encode")
        mean, logvar = mom.chunk(2, dim=1)
        return mean.contiguous(), logvar.contiguous()

    @torch.no_grad()
    def encode_sample(self, img: torch.Tensor, generator: Optional[torch.Generator] = None) -> torch.Tensor:
        """latent_dist.sample(generator) (:234): mean + exp(0.5 logvar) * randn, noise drawn like diffusers' randn_tensor."""
        mean, logvar = self.encode_moments(img)
        gdev = generator.device if generator is not None else self.device
        noise = torch.randn(mean.shape, generator=generator, device=gdev, dtype=torch.bfloat16).to(self.device).float()
        return (mean + torch.exp(0.5 * logvar) * noise).to(torch.bfloat16)
