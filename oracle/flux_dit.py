"""Oracle: eager PyTorch restatement of the FLUX.1-dev MM-DiT forward.

TEST INFRASTRUCTURE (see oracle/__init__.py).  PARITY UNPINNED: the arithmetic
lives in `diffusers.FluxTransformer2DModel` [ext, not in /root/reference, not
installed]; this file restates that published architecture and anchors on the
reference call sites:

  * call signature + tensor shapes      flux_piplines/texturing/pipeline.py:646-656
  * joint attention math                flux_piplines/texturing/attention_processor.py:24-110
  * zero text embeddings                flux_piplines/texturing/pipeline.py:538-543
  * LoRA target module names            flux_piplines/texturing/trainer.py:283-305

Parameter names follow the diffusers state-dict so a real FLUX.1-dev checkpoint
would load unchanged.  The op sequence is the unfused eager one (Linear ->
LayerNorm -> modulate -> SDPA -> GELU-tanh ...), evaluated in the dtype of the
parameters: fp32 = error yardstick, bf16 = the reference's own rounding points
(fp32 islands where diffusers upcasts: sinusoid, RoPE, RMSNorm mean, LN stats).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, Tuple

import torch
import torch.nn.functional as F


@dataclass(frozen=True)
class FluxConfig:
    """FLUX.1-dev `transformer/config.json` [ext]; defaults are the real model."""
    in_channels: int = 64
    num_layers: int = 19
    num_single_layers: int = 38
    attention_head_dim: int = 128
    num_attention_heads: int = 24
    joint_attention_dim: int = 4096
    pooled_projection_dim: int = 768
    guidance_embeds: bool = True
    axes_dims_rope: Tuple[int, int, int] = (16, 56, 56)
    mlp_ratio: int = 4
    rope_theta: float = 10000.0

    @property
    def inner_dim(self) -> int:
        return self.attention_head_dim * self.num_attention_heads

    @property
    def mlp_dim(self) -> int:
        return self.inner_dim * self.mlp_ratio

    @staticmethod
    def tiny(num_layers: int = 2, num_single_layers: int = 2, heads: int = 2) -> "FluxConfig":
        """Same topology, head_dim kept at 128 (RoPE axes need it), 2 heads."""
        return FluxConfig(num_layers=num_layers, num_single_layers=num_single_layers,
                          num_attention_heads=heads, joint_attention_dim=256,
                          pooled_projection_dim=64)


def _linear_names(cfg: FluxConfig):
    """(name, out_features, in_features) of every Linear in state-dict order."""
    D, M = cfg.inner_dim, cfg.mlp_dim
    out = [("x_embedder", D, cfg.in_channels),
           ("context_embedder", D, cfg.joint_attention_dim),
           ("time_text_embed.timestep_embedder.linear_1", D, 256),
           ("time_text_embed.timestep_embedder.linear_2", D, D)]
    if cfg.guidance_embeds:
        out += [("time_text_embed.guidance_embedder.linear_1", D, 256),
                ("time_text_embed.guidance_embedder.linear_2", D, D)]
    out += [("time_text_embed.text_embedder.linear_1", D, cfg.pooled_projection_dim),
            ("time_text_embed.text_embedder.linear_2", D, D)]
    for i in range(cfg.num_layers):
        p = f"transformer_blocks.{i}."
        out += [(p + "norm1.linear", 6 * D, D), (p + "norm1_context.linear", 6 * D, D)]
        for n in ("to_q", "to_k", "to_v", "add_q_proj", "add_k_proj", "add_v_proj",
                  "to_out.0", "to_add_out"):
            out.append((p + "attn." + n, D, D))
        out += [(p + "ff.net.0.proj", M, D), (p + "ff.net.2", D, M),
                (p + "ff_context.net.0.proj", M, D), (p + "ff_context.net.2", D, M)]
    for i in range(cfg.num_single_layers):
        p = f"single_transformer_blocks.{i}."
        out += [(p + "norm.linear", 3 * D, D)]
        for n in ("to_q", "to_k", "to_v"):
            out.append((p + "attn." + n, D, D))
        out += [(p + "proj_mlp", M, D), (p + "proj_out", D, D + M)]
    out += [("norm_out.linear", 2 * D, D), ("proj_out", cfg.in_channels, D)]
    return out


def _rms_names(cfg: FluxConfig):
    out = []
    for i in range(cfg.num_layers):
        p = f"transformer_blocks.{i}.attn."
        out += [p + "norm_q.weight", p + "norm_k.weight",
                p + "norm_added_q.weight", p + "norm_added_k.weight"]
    for i in range(cfg.num_single_layers):
        p = f"single_transformer_blocks.{i}.attn."
        out += [p + "norm_q.weight", p + "norm_k.weight"]
    return out


def init_params(cfg: FluxConfig, seed: int = 0, dtype=torch.float32, device="cpu",
                std: float = 0.02, norm_weight_std: float = 0.0) -> Dict[str, torch.Tensor]:
    """Random-init weights: every Linear weight and bias ~ N(0, std^2), RMSNorm
    weights = 1 (+ N(0, norm_weight_std^2) so tests exercise the multiply).
    Drawn tensor by tensor from one seeded generator on `device` (SURVEY 8d cfg 1).
    """
    g = torch.Generator(device=device).manual_seed(seed)
    P: Dict[str, torch.Tensor] = {}
    for name, o, i in _linear_names(cfg):
        P[name + ".weight"] = (torch.randn(o, i, generator=g, device=device, dtype=torch.float32) * std).to(dtype)
        P[name + ".bias"] = (torch.randn(o, generator=g, device=device, dtype=torch.float32) * std).to(dtype)
    for name in _rms_names(cfg):
        w = torch.ones(cfg.attention_head_dim, device=device, dtype=torch.float32)
        if norm_weight_std:
            w = w + torch.randn(cfg.attention_head_dim, generator=g, device=device) * norm_weight_std
        P[name] = w.to(dtype)
    return P


# ---------------------------------------------------------------------------
# building blocks
# ---------------------------------------------------------------------------
def sinusoid_256(t: torch.Tensor) -> torch.Tensor:
    """diffusers `get_timestep_embedding(t, 256, flip_sin_to_cos=True, shift=0)` [ext]:
    freq_i = exp(-ln(1e4) i/128); out = cat[cos(t f), sin(t f)]  (fp32)."""
    half = 128
    exponent = -math.log(10000.0) * torch.arange(half, dtype=torch.float32, device=t.device) / half
    emb = t[:, None].float() * torch.exp(exponent)[None, :]
    return torch.cat([torch.cos(emb), torch.sin(emb)], dim=-1)


def rope_table(ids: torch.Tensor, cfg: FluxConfig) -> Tuple[torch.Tensor, torch.Tensor]:
    """`FluxPosEmbed` [ext]: per axis, freqs = pos (x) theta^(-2i/d) in float64, cos/sin
    repeat-interleaved x2, axes concatenated -> ([S,128], [S,128]) fp32."""
    pos = ids.float()
    cos_out, sin_out = [], []
    for a, d in enumerate(cfg.axes_dims_rope):
        freqs = 1.0 / (cfg.rope_theta ** (torch.arange(0, d, 2, dtype=torch.float64, device=ids.device) / d))
        ang = torch.outer(pos[:, a].to(torch.float64), freqs)
        cos_out.append(ang.cos().repeat_interleave(2, dim=1).float())
        sin_out.append(ang.sin().repeat_interleave(2, dim=1).float())
    return torch.cat(cos_out, dim=-1), torch.cat(sin_out, dim=-1)


def apply_rope(x: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor) -> torch.Tensor:
    """`apply_rotary_emb` [ext] as used at attention_processor.py:85-87.  x: [B,H,S,hd]."""
    xr, xi = x.reshape(*x.shape[:-1], -1, 2).unbind(-1)
    x_rot = torch.stack([-xi, xr], dim=-1).flatten(3)
    return (x.float() * cos[None, None] + x_rot.float() * sin[None, None]).to(x.dtype)


def rms_norm(x: torch.Tensor, weight: torch.Tensor, eps: float = 1e-6) -> torch.Tensor:
    """diffusers `RMSNorm` [ext]: fp32 mean of squares, cast to weight dtype, x weight."""
    var = x.to(torch.float32).pow(2).mean(-1, keepdim=True)
    x = x * torch.rsqrt(var + eps)
    if weight.dtype in (torch.float16, torch.bfloat16):
        x = x.to(weight.dtype)
    return x * weight


def layer_norm(x: torch.Tensor) -> torch.Tensor:
    """nn.LayerNorm(D, eps=1e-6, elementwise_affine=False)."""
    return F.layer_norm(x, (x.shape[-1],), eps=1e-6)


def _lin(P, name, x):
    return F.linear(x, P[name + ".weight"], P[name + ".bias"])


def time_text_embed(P, cfg: FluxConfig, timestep, guidance, pooled):
    """`CombinedTimestepGuidanceTextProjEmbeddings` [ext].  timestep/guidance are
    already x1000 and in the model dtype."""
    dt = pooled.dtype
    t = _lin(P, "time_text_embed.timestep_embedder.linear_2",
             F.silu(_lin(P, "time_text_embed.timestep_embedder.linear_1", sinusoid_256(timestep).to(dt))))
    if cfg.guidance_embeds:
        g = _lin(P, "time_text_embed.guidance_embedder.linear_2",
                 F.silu(_lin(P, "time_text_embed.guidance_embedder.linear_1", sinusoid_256(guidance).to(dt))))
        t = t + g
    p = _lin(P, "time_text_embed.text_embedder.linear_2",
             F.silu(_lin(P, "time_text_embed.text_embedder.linear_1", pooled)))
    return t + p


def _heads(x, H):
    B, S, _ = x.shape
    return x.view(B, S, H, -1).transpose(1, 2)


def joint_attention(P, prefix, cfg, x, ctx, cos, sin):
    """attention_processor.py:31-110 (NativeFluxAttnProcessor2_0) restated."""
    H = cfg.num_attention_heads
    q = rms_norm(_heads(_lin(P, prefix + "to_q", x), H), P[prefix + "norm_q.weight"])
    k = rms_norm(_heads(_lin(P, prefix + "to_k", x), H), P[prefix + "norm_k.weight"])
    v = _heads(_lin(P, prefix + "to_v", x), H)
    if ctx is not None:
        cq = rms_norm(_heads(_lin(P, prefix + "add_q_proj", ctx), H), P[prefix + "norm_added_q.weight"])
        ck = rms_norm(_heads(_lin(P, prefix + "add_k_proj", ctx), H), P[prefix + "norm_added_k.weight"])
        cv = _heads(_lin(P, prefix + "add_v_proj", ctx), H)
        q, k, v = torch.cat([cq, q], 2), torch.cat([ck, k], 2), torch.cat([cv, v], 2)   # [txt, img] :81-83
    q, k = apply_rope(q, cos, sin), apply_rope(k, cos, sin)                              # :85-87
    o = F.scaled_dot_product_attention(q, k, v, dropout_p=0.0, is_causal=False)         # :89-91
    o = o.transpose(1, 2).reshape(x.shape[0], -1, cfg.inner_dim).to(q.dtype)
    if ctx is not None:
        n = ctx.shape[1]
        return _lin(P, prefix + "to_out.0", o[:, n:]), _lin(P, prefix + "to_add_out", o[:, :n])
    return o


def double_block(P, i, cfg, x, ctx, temb, cos, sin):
    """`FluxTransformerBlock.forward` [ext]."""
    p = f"transformer_blocks.{i}."
    se = F.silu(temb)
    sh_a, sc_a, g_a, sh_m, sc_m, g_m = _lin(P, p + "norm1.linear", se).chunk(6, dim=1)
    csh_a, csc_a, cg_a, csh_m, csc_m, cg_m = _lin(P, p + "norm1_context.linear", se).chunk(6, dim=1)
    nx = layer_norm(x) * (1 + sc_a[:, None]) + sh_a[:, None]
    nc = layer_norm(ctx) * (1 + csc_a[:, None]) + csh_a[:, None]
    ao, co = joint_attention(P, p + "attn.", cfg, nx, nc, cos, sin)
    x = x + g_a.unsqueeze(1) * ao
    nx = layer_norm(x) * (1 + sc_m[:, None]) + sh_m[:, None]
    ff = _lin(P, p + "ff.net.2", F.gelu(_lin(P, p + "ff.net.0.proj", nx), approximate="tanh"))
    x = x + g_m.unsqueeze(1) * ff
    ctx = ctx + cg_a.unsqueeze(1) * co
    nc = layer_norm(ctx) * (1 + csc_m[:, None]) + csh_m[:, None]
    cff = _lin(P, p + "ff_context.net.2", F.gelu(_lin(P, p + "ff_context.net.0.proj", nc), approximate="tanh"))
    ctx = ctx + cg_m.unsqueeze(1) * cff
    if ctx.dtype == torch.float16:
        ctx = ctx.clip(-65504, 65504)
    return ctx, x


def single_block(P, i, cfg, x, temb, cos, sin):
    """`FluxSingleTransformerBlock.forward` [ext]."""
    p = f"single_transformer_blocks.{i}."
    sh, sc, gate = _lin(P, p + "norm.linear", F.silu(temb)).chunk(3, dim=1)
    nx = layer_norm(x) * (1 + sc[:, None]) + sh[:, None]
    mlp = F.gelu(_lin(P, p + "proj_mlp", nx), approximate="tanh")
    ao = joint_attention(P, p + "attn.", cfg, nx, None, cos, sin)
    out = gate.unsqueeze(1) * _lin(P, p + "proj_out", torch.cat([ao, mlp], dim=2))
    x = x + out
    if x.dtype == torch.float16:
        x = x.clip(-65504, 65504)
    return x


@torch.no_grad()
def flux_forward(P: Dict[str, torch.Tensor], cfg: FluxConfig, hidden_states, timestep, guidance,
                 pooled_projections, encoder_hidden_states, txt_ids, img_ids,
                 return_intermediates: bool = False, scalar_dtype=torch.bfloat16):
    """`FluxTransformer2DModel.forward` [ext] as called at
    flux_piplines/texturing/pipeline.py:646-656.  hidden_states [B,S_img,64],
    timestep [B] (= t/1000), guidance [B] fp32, pooled [B,768], enc [B,S_txt,4096],
    txt_ids [S_txt,3], img_ids [S_img,3]  ->  [B,S_img,64].

    `scalar_dtype`: the reference always runs the transformer in bf16 (pipeline.py:102), so
    `timestep.to(dtype) * 1000` / `guidance.to(dtype) * 1000` round to bf16 (3.5 -> 3504).  That
    changes the *value* fed to the sinusoid, not just noise, so the fp32 yardstick keeps it."""
    dt = P["x_embedder.weight"].dtype
    hidden_states = hidden_states.to(dt)
    x = _lin(P, "x_embedder", hidden_states)
    timestep = (timestep.to(scalar_dtype) * 1000).to(dt)
    guidance = (guidance.to(scalar_dtype) * 1000).to(dt) if guidance is not None else None
    temb = time_text_embed(P, cfg, timestep, guidance, pooled_projections.to(dt))
    ctx = _lin(P, "context_embedder", encoder_hidden_states.to(dt))
    ids = torch.cat([txt_ids, img_ids], dim=0)
    cos, sin = rope_table(ids, cfg)
    inter = {"temb": temb, "x0": x, "ctx0": ctx}
    for i in range(cfg.num_layers):
        ctx, x = double_block(P, i, cfg, x, ctx, temb, cos, sin)
        if return_intermediates:
            inter[f"double{i}.x"], inter[f"double{i}.ctx"] = x, ctx
    n_txt = ctx.shape[1]
    x = torch.cat([ctx, x], dim=1)
    for i in range(cfg.num_single_layers):
        x = single_block(P, i, cfg, x, temb, cos, sin)
        if return_intermediates:
            inter[f"single{i}.x"] = x
    x = x[:, n_txt:]
    scale, shift = _lin(P, "norm_out.linear", F.silu(temb).to(x.dtype)).chunk(2, dim=1)   # AdaLayerNormContinuous: (scale, shift)
    x = layer_norm(x) * (1 + scale)[:, None, :] + shift[:, None, :]
    out = _lin(P, "proj_out", x)
    return (out, inter) if return_intermediates else out


def dit_flops(cfg: FluxConfig, S: int, S_txt: int = 512) -> float:
    """Algorithmic FLOPs of one forward: linears + 4 S^2 D per block + small."""
    D, M = cfg.inner_dim, cfg.mlp_dim
    per_tok_double = 2 * (4 * D * D + 2 * D * M)
    per_tok_single = 2 * (3 * D * D + D * M + (D + M) * D)
    lin = cfg.num_layers * per_tok_double * S + cfg.num_single_layers * per_tok_single * S
    attn = (cfg.num_layers + cfg.num_single_layers) * 4 * D * S * S
    S_img = S - S_txt
    small = 2 * (S_img * cfg.in_channels * D * 2 + S_txt * cfg.joint_attention_dim * D)
    small += 2 * D * D * (cfg.num_layers * 12 + cfg.num_single_layers * 3 + 2 + 6)
    return float(lin + attn + small)
