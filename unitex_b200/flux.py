"""Host side of the B200 FLUX.1-dev MM-DiT engine: weight packing, LoRA merge, workspace, and the calls into
`utx_flux_*` (include/unitex_b200.h).  Mirrors the surface the reference uses on diffusers'
`FluxTransformer2DModel` (flux_piplines/texturing/pipeline.py:646-656) -- `forward(hidden_states, timestep, guidance,
pooled_projections, encoder_hidden_states, txt_ids, img_ids)` -- plus `denoise()` for the whole Euler loop (:634-681).

All arithmetic happens in libunitex_b200.so; torch is used for device memory and streams only.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Dict, Optional, Tuple

import numpy as np
import torch

from . import _lib, ops


@dataclass(frozen=True)
class FluxConfig:
    """FLUX.1-dev transformer/config.json; defaults = the real model (19 double + 38 single, 24 x 128, ctx 4096)."""
    in_channels: int = 64
    num_layers: int = 19
    num_single_layers: int = 38
    attention_head_dim: int = 128
    num_attention_heads: int = 24
    joint_attention_dim: int = 4096
    pooled_projection_dim: int = 768
    guidance_embeds: bool = True
    mlp_ratio: int = 4

    @property
    def inner_dim(self):
        return self.attention_head_dim * self.num_attention_heads

    @property
    def n_mod_rows(self):
        return (12 * self.num_layers + 3 * self.num_single_layers + 2) * self.inner_dim


_DBL_PARTS = {  # packed name -> state-dict sub-names stacked on rows
    "qkv_img": ("attn.to_q", "attn.to_k", "attn.to_v"),
    "qkv_txt": ("attn.add_q_proj", "attn.add_k_proj", "attn.add_v_proj"),
    "out_img": ("attn.to_out.0",), "out_txt": ("attn.to_add_out",),
    "ff1_img": ("ff.net.0.proj",), "ff2_img": ("ff.net.2",),
    "ff1_txt": ("ff_context.net.0.proj",), "ff2_txt": ("ff_context.net.2",),
}
_DBL_RMS = {"rms_q_img": "attn.norm_q", "rms_k_img": "attn.norm_k", "rms_q_txt": "attn.norm_added_q",
            "rms_k_txt": "attn.norm_added_k"}
_SGL_PARTS = {"qkvmlp": ("attn.to_q", "attn.to_k", "attn.to_v", "proj_mlp"), "out": ("proj_out",)}
_SGL_RMS = {"rms_q": "attn.norm_q", "rms_k": "attn.norm_k"}
_TOP = {"x_embed": "x_embedder", "ctx_embed": "context_embedder",
        "t1": "time_text_embed.timestep_embedder.linear_1", "t2": "time_text_embed.timestep_embedder.linear_2",
        "g1": "time_text_embed.guidance_embedder.linear_1", "g2": "time_text_embed.guidance_embedder.linear_2",
        "p1": "time_text_embed.text_embedder.linear_1", "p2": "time_text_embed.text_embedder.linear_2",
        "proj_out": "proj_out"}


class FluxTransformer:
    """Packed bf16 weights on one GPU + a `utx_flux` handle."""

    def __init__(self, cfg: FluxConfig, device="cuda"):
        self.cfg = cfg
        self.device = torch.device(device)
        self.lib = _lib.load()
        self.T: Dict[str, torch.Tensor] = {}          # packed tensors
        self.where: Dict[str, Tuple[str, int, int]] = {}   # state-dict linear name -> (packed key, row0, row1)
        self._handle = _lib.vp()
        c = _lib.FluxConfigC(cfg.in_channels, cfg.num_layers, cfg.num_single_layers, cfg.num_attention_heads,
                             cfg.attention_head_dim, cfg.joint_attention_dim, cfg.pooled_projection_dim,
                             int(cfg.guidance_embeds), cfg.mlp_ratio)
        _lib.check(self.lib.utx_flux_create(C.byref(c), C.byref(self._handle)), "utx_flux_create")
        self._ws = None
        self._keep = None
        self._layout()

    def __del__(self):
        try:
            if self._handle:
                self.lib.utx_flux_destroy(self._handle)
        except Exception:
            pass

    # ------------------------------------------------------------------ packed layout
    def _layout(self):
        cfg, D = self.cfg, self.cfg.inner_dim
        M = D * cfg.mlp_ratio
        dims = {"x_embedder": (D, cfg.in_channels), "context_embedder": (D, cfg.joint_attention_dim),
                "time_text_embed.timestep_embedder.linear_1": (D, 256), "time_text_embed.timestep_embedder.linear_2": (D, D),
                "time_text_embed.guidance_embedder.linear_1": (D, 256), "time_text_embed.guidance_embedder.linear_2": (D, D),
                "time_text_embed.text_embedder.linear_1": (D, cfg.pooled_projection_dim),
                "time_text_embed.text_embedder.linear_2": (D, D), "proj_out": (cfg.in_channels, D)}
        sub = {"attn.to_q": (D, D), "attn.to_k": (D, D), "attn.to_v": (D, D), "attn.add_q_proj": (D, D),
               "attn.add_k_proj": (D, D), "attn.add_v_proj": (D, D), "attn.to_out.0": (D, D), "attn.to_add_out": (D, D),
               "ff.net.0.proj": (M, D), "ff.net.2": (D, M), "ff_context.net.0.proj": (M, D), "ff_context.net.2": (D, M),
               "proj_mlp": (M, D), "proj_out": (D, D + M)}
        self.shapes: Dict[str, Tuple[int, int]] = {}
        self.rms_names: Dict[str, str] = {}
        for k, n in _TOP.items():
            if not cfg.guidance_embeds and k in ("g1", "g2"):
                continue
            self.shapes[k] = dims[n]
            self.where[n] = (k, 0, dims[n][0])
        for i in range(cfg.num_layers):
            for k, parts in _DBL_PARTS.items():
                r = 0
                for pn in parts:
                    self.where[f"transformer_blocks.{i}.{pn}"] = (f"d{i}.{k}", r, r + sub[pn][0])
                    r += sub[pn][0]
                self.shapes[f"d{i}.{k}"] = (r, sub[parts[0]][1])
            for k, n in _DBL_RMS.items():
                self.rms_names[f"d{i}.{k}"] = f"transformer_blocks.{i}.{n}.weight"
        for i in range(cfg.num_single_layers):
            for k, parts in _SGL_PARTS.items():
                r = 0
                for pn in parts:
                    self.where[f"single_transformer_blocks.{i}.{pn}"] = (f"s{i}.{k}", r, r + sub[pn][0])
                    r += sub[pn][0]
                self.shapes[f"s{i}.{k}"] = (r, sub[parts[0]][1])
            for k, n in _SGL_RMS.items():
                self.rms_names[f"s{i}.{k}"] = f"single_transformer_blocks.{i}.{n}.weight"
        r = 0
        for i in range(cfg.num_layers):
            for n in ("norm1.linear", "norm1_context.linear"):
                self.where[f"transformer_blocks.{i}.{n}"] = ("mod", r, r + 6 * D)
                r += 6 * D
        for i in range(cfg.num_single_layers):
            self.where[f"single_transformer_blocks.{i}.norm.linear"] = ("mod", r, r + 3 * D)
            r += 3 * D
        self.where["norm_out.linear"] = ("mod", r, r + 2 * D)
        r += 2 * D
        assert r == cfg.n_mod_rows
        self.shapes["mod"] = (r, D)

    def _alloc(self):
        for k, (o, i) in self.shapes.items():
            self.T["w_" + k] = torch.empty(o, i, device=self.device, dtype=torch.bfloat16)
            self.T["b_" + k] = torch.empty(o, device=self.device, dtype=torch.bfloat16)
        for k in self.rms_names:
            self.T[k] = torch.ones(128, device=self.device, dtype=torch.bfloat16)

    # ------------------------------------------------------------------ weights
    def random_init_(self, seed: int = 0, std: float = 0.02):
        """Random-init FLUX-shaped weights generated directly in the packed layout on the device (bench path:
        BASELINE.json asks for random-init weights of the real architecture)."""
        self._alloc()
        g = torch.Generator(device=self.device).manual_seed(seed)
        for k in self.shapes:
            self.T["w_" + k].normal_(0.0, std, generator=g)
            self.T["b_" + k].normal_(0.0, std, generator=g)
        self._commit()
        return self

    def load_state_dict(self, sd: Dict[str, torch.Tensor]):
        """diffusers-named state dict (any dtype/device) -> packed bf16 device tensors."""
        self._alloc()
        for name, (k, r0, r1) in self.where.items():
            self.T["w_" + k][r0:r1].copy_(sd[name + ".weight"].to(torch.bfloat16))
            self.T["b_" + k][r0:r1].copy_(sd[name + ".bias"].to(torch.bfloat16))
        for k, n in self.rms_names.items():
            self.T[k].copy_(sd[n].to(torch.bfloat16))
        self._commit()
        return self

    def clone(self) -> "FluxTransformer":
        """Second resident weight set (texture-merged / delight-merged: 2 x 23.7 GB fits 180 GB HBM)."""
        o = FluxTransformer(self.cfg, self.device)
        o.T = {k: v.clone() for k, v in self.T.items()}
        o._commit()
        return o

    def merge_lora_(self, lora: Dict[str, torch.Tensor], scale: float, replace_modules: bool = True):
        """W += scale * (alpha / r) * B @ A for every `<linear>.lora_A.weight / lora_B.weight` pair (peft naming without the
        `transformer.` prefix).  `scale` = adapter_weight * lora_alpha / r as resolved by the caller (pipeline.py:245,263); a
        per-module `<linear>.alpha` entry (kohya-style files) overrides the file-level ratio for that module:
        scale_module = scale * alpha / r.
        Plain `<module>.weight / .bias` entries are peft `modules_to_save` replacements (trainer.py:297-304: x_embedder): they
        REPLACE the base tensor and, unlike LoRA deltas, do not blend -- peft routes the forward through the active adapter's
        copy.  `replace_modules=False` skips them (adapters merged with weight 0, or all but the last active adapter:
        `PBRFluxPipeline.set_adapters` lets the LAST adapter with a non-zero weight win, which is peft's behaviour for the
        reference's two calls, where exactly one of texture / delight is active).
        Any other key suffix, or a module this model does not have, raises: silently dropping or mis-broadcasting a tensor
        would corrupt the weights."""
        names = sorted(k[: -len(".lora_A.weight")] for k in lora if k.endswith(".lora_A.weight"))
        handled = set()
        for n in names:
            if n not in self.where:
                raise KeyError(f"merge_lora_: LoRA target '{n}' is not a Linear of this transformer")
            if n + ".lora_B.weight" not in lora:
                raise KeyError(f"merge_lora_: '{n}.lora_A.weight' without its lora_B")
            k, r0, r1 = self.where[n]
            A, B = lora[n + ".lora_A.weight"], lora[n + ".lora_B.weight"]
            s = float(scale)
            if n + ".alpha" in lora:
                s = s * float(lora[n + ".alpha"]) / A.shape[0]
                handled.add(n + ".alpha")
            ops.lora_merge_(self.T["w_" + k][r0:r1], A, B, s)
            handled.update((n + ".lora_A.weight", n + ".lora_B.weight"))
        for key, v in lora.items():
            if key in handled:
                continue
            n, _, kind = key.rpartition(".")
            if ".lora_" in key or kind not in ("weight", "bias") or n not in self.where:
                raise KeyError(f"merge_lora_: unsupported adapter entry '{key}'")
            if not replace_modules:
                continue
            k, r0, r1 = self.where[n]
            dst = self.T[("w_" if kind == "weight" else "b_") + k][r0:r1]
            if tuple(v.shape) != tuple(dst.shape):
                raise ValueError(f"merge_lora_: '{key}' has shape {tuple(v.shape)}, the module expects {tuple(dst.shape)}")
            dst.copy_(v.to(torch.bfloat16))
        return self

    def _commit(self):
        cfg, T = self.cfg, self.T
        p = lambda k: T[k].data_ptr()
        dbl = (_lib.DoubleBlockC * cfg.num_layers)()
        for i in range(cfg.num_layers):
            for f in _lib._DBL:
                setattr(dbl[i], f, p(f"d{i}.{f}" if f.startswith("rms_") else f"{f[:2]}d{i}.{f[2:]}"))
        sgl = (_lib.SingleBlockC * cfg.num_single_layers)()
        for i in range(cfg.num_single_layers):
            for f in _lib._SGL:
                setattr(sgl[i], f, p(f"s{i}.{f}" if f.startswith("rms_") else f"{f[:2]}s{i}.{f[2:]}"))
        w = _lib.FluxWeightsC()
        for k in _TOP:
            if k == "proj_out":
                continue
            for pre in ("w_", "b_"):
                setattr(w, f"{pre}{k}", p(pre + k) if (pre + k) in T else None)
        w.w_mod, w.b_mod = p("w_mod"), p("b_mod")
        w.w_proj_out, w.b_proj_out = p("w_proj_out"), p("b_proj_out")
        w.double_blocks, w.single_blocks = dbl, sgl
        self._keep = (dbl, sgl, w)
        _lib.check(self.lib.utx_flux_set_weights(self._handle, C.byref(w)), "utx_flux_set_weights")

    # ------------------------------------------------------------------ per-call
    def prepare(self, ids: torch.Tensor, encoder_hidden_states: Optional[torch.Tensor] = None,
                pooled_projections: Optional[torch.Tensor] = None, s_txt: int = 512):
        """ids = cat(txt_ids, img_ids) [S,3]; enc [s_txt, joint_dim] (None -> zeros, pipeline.py:538-543)."""
        cfg = self.cfg
        S = ids.shape[0]
        self.s_txt, self.s_img = s_txt, S - s_txt
        if encoder_hidden_states is None:
            encoder_hidden_states = torch.zeros(s_txt, cfg.joint_attention_dim, device=self.device, dtype=torch.bfloat16)
        if pooled_projections is None:
            pooled_projections = torch.zeros(cfg.pooled_projection_dim, device=self.device, dtype=torch.float32)
        enc = encoder_hidden_states.reshape(s_txt, cfg.joint_attention_dim).to(self.device, torch.bfloat16).contiguous()
        pooled = pooled_projections.reshape(-1).to(self.device, torch.float32).contiguous()
        ids = ids.to(self.device, torch.float32).contiguous()
        if getattr(self, "_sp_direct", False):
            self._bind_peer_region()
        nbytes = self.lib.utx_flux_workspace_bytes(self._handle, self.s_txt, self.s_img)
        if nbytes == 0:
            raise _lib.UtxError("utx_flux_workspace_bytes: the sequence length must be a multiple of the sequence-parallel world size")
        if self._ws is None or self._ws.numel() < nbytes:
            self._ws = torch.empty(nbytes, device=self.device, dtype=torch.uint8)
        _lib.check(self.lib.utx_flux_prepare(self._handle, self._ws.data_ptr(), nbytes, ids.data_ptr(), enc.data_ptr(),
                                             pooled.data_ptr(), self.s_txt, self.s_img, ops._stream()), "utx_flux_prepare")
        return self

    def forward(self, latents: torch.Tensor, timestep: float, guidance: float, out: Optional[torch.Tensor] = None):
        """latents [s_img, 64] bf16 -> v [s_img, 64] bf16; timestep = t/1000, guidance raw (:648-649)."""
        assert latents.is_cuda and latents.dtype == torch.bfloat16 and latents.is_contiguous()
        assert latents.shape == (self.s_img, self.cfg.in_channels)
        if out is None:
            out = torch.empty_like(latents)
        _lib.check(self.lib.utx_flux_forward(self._handle, latents.data_ptr(), float(timestep), float(guidance),
                                             out.data_ptr(), ops._stream()), "utx_flux_forward")
        return out

    def denoise_(self, latents: torch.Tensor, s_noise: int, sigmas, guidance: float = 3.5):
        """In-place Euler loop over `latents` [s_img, 64] (noise rows first, clean condition rows after)."""
        assert latents.is_cuda and latents.dtype == torch.bfloat16 and latents.is_contiguous()
        sig = np.ascontiguousarray(np.asarray(sigmas, dtype=np.float32))
        _lib.check(self.lib.utx_flux_denoise(self._handle, latents.data_ptr(), int(s_noise),
                                             sig.ctypes.data_as(_lib.fp), len(sig) - 1, float(guidance), ops._stream()),
                   "utx_flux_denoise")
        return latents

    def set_sequence_parallel(self, comm=None, direct: bool = False):
        """Sequence-parallel ("Ulysses") mode over the ranks of `comm` (a `parallel.TileComm`; None = off): ONE grid's tokens are
        split over the ranks, see include/unitex_b200.h: utx_flux_set_sequence_parallel.  Every rank then calls prepare /
        forward / denoise_ with the SAME arguments and ends with the same result.  Call prepare() again afterwards.
        direct=True: the exchanges around each attention are fused into the producing kernels' epilogues over NVLink peer memory
        (utx_flux_set_sp_peers) instead of NCCL all-to-alls; the peer region is allocated at the next prepare()."""
        self._sp_comm = comm                       # keep the communicator alive as long as the engine uses it
        self._sp_direct = bool(direct) and comm is not None
        if getattr(self, "_sp_region", None) is not None:
            _lib.check(self.lib.utx_flux_set_sp_peers(self._handle, None, 0), "utx_flux_set_sp_peers")
            self._sp_region.close()
            self._sp_region = None
        _lib.check(self.lib.utx_flux_set_sequence_parallel(self._handle, comm._handle if comm is not None else None),
                   "utx_flux_set_sequence_parallel")
        self._ws = None
        return self

    def _bind_peer_region(self):
        """direct mode: (re)allocate the peer exchange region for the current sequence and hand the mappings to the engine."""
        from .parallel import PeerRegion
        need = self.lib.utx_flux_sp_region_bytes(self._handle, self.s_txt, self.s_img)
        reg = getattr(self, "_sp_region", None)
        if reg is not None and reg.nbytes >= need:
            return
        if reg is not None:
            _lib.check(self.lib.utx_flux_set_sp_peers(self._handle, None, 0), "utx_flux_set_sp_peers")
            reg.close()
        self._sp_region = PeerRegion(need, self.device)
        arr = (_lib.vp * len(self._sp_region.ptrs))(*self._sp_region.ptrs)
        _lib.check(self.lib.utx_flux_set_sp_peers(self._handle, arr, need), "utx_flux_set_sp_peers")

    def graph_replays(self) -> int:
        """Denoise steps that ran as one CUDA-graph launch (include/unitex_b200.h: utx_flux_graph_replays)."""
        return int(self.lib.utx_flux_graph_replays(self._handle))

    def profile(self, enable: bool = True):
        _lib.check(self.lib.utx_flux_profile(self._handle, int(enable)), "utx_flux_profile")

    def profile_read(self, reset: bool = True):
        """-> ({'gemm','attn','elem','other'} launches, same keys -> ms) accumulated since the last reset."""
        n = (C.c_long * 4)()
        ms = (C.c_float * 4)()
        _lib.check(self.lib.utx_flux_profile_read(self._handle, n, ms, int(reset)), "utx_flux_profile_read")
        keys = ("gemm", "attn", "elem", "other")
        return dict(zip(keys, list(n))), dict(zip(keys, list(ms)))

    def weight_bytes(self) -> int:
        return sum(t.numel() * t.element_size() for t in self.T.values())
