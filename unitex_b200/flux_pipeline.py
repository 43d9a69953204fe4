"""`PBRFluxPipeline`: the reference's condition-token FLUX sampler, host side in Python, all arithmetic in
libunitex_b200.so.  Drop-in for `flux_piplines/{texturing,delight}/pipeline.py::PBRFluxPipeline` of the reference
(the two files are byte-identical there); same constructor helpers (`from_pretrained`, `load_lora_weights`,
`set_adapters`, `to`) and the same `__call__` keyword surface (reference pipeline.py:404-431), no diffusers/peft.

Differences that are deliberate and documented in DESIGN.md:
  * LoRA adapters are MERGED into the bf16 weights (one resident weight set per adapter-weight vector; 2 x 23.7 GB fits
    a B200) instead of being evaluated as two extra GEMMs per Linear per adapter.
  * text encoders are never built: the reference loads them as None and feeds zeros (pipeline.py:104-105,538-543).
"""
from __future__ import annotations

import json
import math
import os
from dataclasses import dataclass
from typing import Any, Callable, Dict, List, Optional, Union

import numpy as np
import torch

from .flux import FluxConfig, FluxTransformer


def calculate_shift(image_seq_len, base_seq_len: int = 256, max_seq_len: int = 4096, base_shift: float = 0.5,
                    max_shift: float = 1.16):
    """reference flux_piplines/texturing/pipeline.py:59-69 (same defaults, incl. the 1.16 that callers override)."""
    m = (max_shift - base_shift) / (max_seq_len - base_seq_len)
    b = base_shift - m * base_seq_len
    return image_seq_len * m + b


@dataclass
class SchedulerConfig:
    """FLUX.1-dev scheduler/scheduler_config.json [ext]."""
    num_train_timesteps: int = 1000
    base_image_seq_len: int = 256
    max_image_seq_len: int = 4096
    base_shift: float = 0.5
    max_shift: float = 1.15
    order: int = 1


class FlowMatchEulerSchedule:
    """The part of diffusers' FlowMatchEulerDiscreteScheduler [ext] the reference uses (retrieve_timesteps with custom
    sigmas + mu, :603-610): dynamic time shift, fp32 sigmas with the terminal 0, timesteps = 1000 sigma."""

    def __init__(self, config: Optional[SchedulerConfig] = None):
        self.config = config or SchedulerConfig()
        self.order = self.config.order
        self.sigmas = None
        self.timesteps = None

    def set_timesteps(self, sigmas, mu: float):
        s = np.array(sigmas).astype(np.float32)          # diffusers [ext] casts the grid to fp32 first, the shift runs in fp32
        s = (math.exp(mu) / (math.exp(mu) + (1 / s - 1) ** 1.0)).astype(np.float32)
        self.timesteps = s * np.float32(self.config.num_train_timesteps)
        self.sigmas = np.concatenate([s, np.zeros(1, np.float32)])
        return self.timesteps


@dataclass
class PBRFluxPipelineOutput:
    images: Union[List[Any], np.ndarray, torch.Tensor]


def _randn(shape, generator, device, dtype):
    """diffusers randn_tensor [ext]: a CPU generator draws on the CPU in `dtype`, then moves (draw order matters:
    noise, dual, control share one generator, reference pipeline.py:152)."""
    gdev = generator.device if generator is not None else torch.device(device)
    return torch.randn(shape, generator=generator, device=gdev, dtype=dtype).to(device)


class PBRFluxPipeline:
    vae_scale_factor = 8
    default_sample_size = 128

    def __init__(self, transformer: FluxTransformer, vae=None, scheduler: Optional[FlowMatchEulerSchedule] = None):
        self.transformer = transformer          # base weights (no adapter)
        self.vae = vae                          # unitex_b200.vae.AutoencoderKLB200 or None
        self.scheduler = scheduler or FlowMatchEulerSchedule()
        self.text_encoder = None
        self.text_encoder_2 = None
        self.redux_pipeline = None
        self._num_inference_steps = 28
        self._adapters: Dict[str, Dict[str, torch.Tensor]] = {}
        self._adapter_scale: Dict[str, float] = {}
        self._merged: Dict[tuple, FluxTransformer] = {}
        self._active = ()
        self._guidance_scale = 3.5
        self._joint_attention_kwargs = None
        self._interrupt = False
        self._num_timesteps = 0
        self.device = transformer.device

    # ------------------------------------------------------------------ construction (reference pipeline.py:96-126)
    @classmethod
    def from_pretrained(cls, ckpt_id: str, transformer=None, text_encoder=None, text_encoder_2=None,
                        torch_dtype=torch.bfloat16, device="cuda", **_):
        """Reads `<ckpt>/transformer/*.safetensors` (+ `<ckpt>/vae/`) in the diffusers layout.  `transformer` may be an
        already built FluxTransformer.  Text encoders are ignored exactly like the reference (None)."""
        if transformer is None:
            from safetensors.torch import load_file
            tdir = os.path.join(ckpt_id, "transformer")
            cfgj = json.load(open(os.path.join(tdir, "config.json")))
            cfg = FluxConfig(in_channels=cfgj["in_channels"], num_layers=cfgj["num_layers"],
                             num_single_layers=cfgj["num_single_layers"], attention_head_dim=cfgj["attention_head_dim"],
                             num_attention_heads=cfgj["num_attention_heads"], joint_attention_dim=cfgj["joint_attention_dim"],
                             pooled_projection_dim=cfgj["pooled_projection_dim"], guidance_embeds=cfgj.get("guidance_embeds", True))
            sd = {}
            for f in sorted(os.listdir(tdir)):
                if f.endswith(".safetensors"):
                    sd.update(load_file(os.path.join(tdir, f)))
            transformer = FluxTransformer(cfg, device).load_state_dict(sd)
        vae = None
        vdir = os.path.join(ckpt_id, "vae")
        if os.path.isdir(vdir):
            from .vae import AutoencoderKLB200
            vae = AutoencoderKLB200.from_pretrained(vdir, device=device)
        return cls(transformer, vae)

    @classmethod
    def from_random(cls, cfg: Optional[FluxConfig] = None, seed: int = 0, device="cuda", with_vae: bool = False):
        tr = FluxTransformer(cfg or FluxConfig(), device).random_init_(seed)
        vae = None
        if with_vae:
            from .vae import AutoencoderKLB200
            vae = AutoencoderKLB200.from_random(seed=seed + 1, device=device)
        return cls(tr, vae)

    def to(self, device=None, **_):
        return self

    def load_lora_weights(self, path_or_dict, adapter_name: str = "default", lora_alpha: Optional[float] = None):
        """peft/diffusers LoRA file: keys `transformer.<module>.lora_A.weight` / `.lora_B.weight` (+ modules_to_save
        full tensors).  scale = alpha / r; without alpha metadata peft's diffusers loader uses alpha = r (scale 1)."""
        if isinstance(path_or_dict, (str, os.PathLike)):
            from safetensors.torch import load_file
            sd = load_file(str(path_or_dict))
        else:
            sd = dict(path_or_dict)
        clean = {}
        for k, v in sd.items():
            k = k[len("transformer."):] if k.startswith("transformer.") else k
            k = k.replace(".lora_A.default.", ".lora_A.").replace(".lora_B.default.", ".lora_B.")
            k = k.replace(".modules_to_save.default.", ".").replace(".base_layer.", ".")
            clean[k] = v
        ranks = [v.shape[0] for k, v in clean.items() if k.endswith(".lora_A.weight")]
        r = ranks[0] if ranks else 1
        self._adapters[adapter_name] = clean
        self._adapter_scale[adapter_name] = (lora_alpha / r) if lora_alpha is not None else 1.0
        self._merged.clear()

    def set_adapters(self, adapter_names, adapter_weights=None):
        """reference pipeline.py:245,263.  Weight 0 adapters contribute exactly 0 and are skipped."""
        if isinstance(adapter_names, str):
            adapter_names = [adapter_names]
        if adapter_weights is None:
            adapter_weights = [1.0] * len(adapter_names)
        self._active = tuple((n, float(w)) for n, w in zip(adapter_names, adapter_weights) if float(w) != 0.0)

    def set_sequence_parallel(self, comm=None, direct: bool = True):
        """ONE asset over all ranks of `comm` (`unitex_b200.parallel.tile_comm(device)`; None = off): every engine of this
        pipeline -- the base weights and each merged adapter set -- splits the token sequence over the ranks
        (`FluxTransformer.set_sequence_parallel`).  Every rank then makes the SAME `__call__` with the same arguments and generator
        seed and gets the same images; the VAE passes are replicated (16 ms against 28 steps)."""
        self._sp = (comm, bool(direct))
        for eng in [self.transformer, *self._merged.values()]:
            eng.set_sequence_parallel(comm, direct=direct)
        return self

    def _engine(self) -> FluxTransformer:
        if not self._active:
            return self.transformer
        if self._active not in self._merged:
            eng = self.transformer.clone()
            for i, (name, w) in enumerate(self._active):
                # LoRA deltas blend by weight; `modules_to_save` replacements (x_embedder) do not -- the last active adapter's
                # copy is the one in effect (see FluxTransformer.merge_lora_)
                eng.merge_lora_(self._adapters[name], w * self._adapter_scale[name], replace_modules=(i == len(self._active) - 1))
            if getattr(self, "_sp", (None, False))[0] is not None:
                eng.set_sequence_parallel(self._sp[0], direct=self._sp[1])
            self._merged[self._active] = eng
        return self._merged[self._active]

    # ------------------------------------------------------------------ latent bookkeeping (:240-275)
    @staticmethod
    def _pack_latents(latents, batch_size, num_channels_latents, height, width, pixel_shuffle=True):
        if pixel_shuffle:
            latents = latents.view(batch_size, num_channels_latents, height // 2, 2, width // 2, 2)
            latents = latents.permute(0, 2, 4, 1, 3, 5)
            return latents.reshape(batch_size, (height // 2) * (width // 2), num_channels_latents * 4)
        return latents.permute(0, 2, 3, 1).reshape(batch_size, height * width, num_channels_latents)

    @staticmethod
    def _unpack_latents(latents, height, width, vae_scale_factor):
        batch_size, num_patches, channels = latents.shape
        height = 2 * (int(height) // (vae_scale_factor * 2))
        width = 2 * (int(width) // (vae_scale_factor * 2))
        latents = latents.view(batch_size, height // 2, width // 2, channels // 4, 2, 2).permute(0, 3, 1, 4, 2, 5)
        return latents.reshape(batch_size, channels // 4, height, width)

    @staticmethod
    def _prepare_latent_image_ids(batch_size, height, width, device, dtype, offset_x=0, offset_y=0, offset_z=0):
        ids = torch.zeros(height, width, 3)
        ids[..., 1] = ids[..., 1] + torch.arange(offset_y, offset_y + height)[:, None]
        ids[..., 2] = ids[..., 2] + torch.arange(offset_x, offset_x + width)[None, :]
        if offset_z != 0:
            ids[..., 0] = ids[..., 0] + offset_z
        return ids.reshape(height * width, 3).to(device=device, dtype=dtype)

    def _encode_vae_image(self, image: torch.Tensor, generator):
        """:226-238: latent_dist.sample(generator), then (z - shift) * scale."""
        if self.vae is None:
            raise RuntimeError("PBRFluxPipeline: a VAE is required to encode control/dual images "
                               "(pass control_latents=/dual_latents= to skip VAE encoding)")
        z = self.vae.encode_sample(image, generator)
        return (z - self.vae.shift_factor) * self.vae.scaling_factor

    @staticmethod
    def _preprocess_image(img) -> torch.Tensor:
        """VaeImageProcessor.preprocess [ext] at native size: PIL/ndarray -> [1,3,H,W] float in [-1,1]."""
        if isinstance(img, torch.Tensor):
            return img
        a = np.asarray(img.convert("RGB") if hasattr(img, "convert") else img).astype(np.float32) / 255.0
        return torch.from_numpy(a).permute(2, 0, 1)[None] * 2.0 - 1.0

    def prepare_latents_and_image_ids(self, batch_size, num_channels_latents, height, width, dtype, device, generator,
                                      dual_image=None, redux_image=None, control_image=None, dual_latents=None,
                                      control_latents=None):
        """:277-402.  Generator draw order: noise, dual, control.  `*_latents` (unpacked [B,16,h,w], already
        shift/scaled) bypass the VAE for callers that hold latents."""
        HL = 2 * (int(height) // (self.vae_scale_factor * 2))
        WL = 2 * (int(width) // (self.vae_scale_factor * 2))
        noise = _randn((batch_size, num_channels_latents, HL, WL), generator, device, dtype)
        noise_latents = self._pack_latents(noise, batch_size, num_channels_latents, HL, WL)
        noise_ids = self._prepare_latent_image_ids(batch_size, HL // 2, WL // 2, device, dtype)

        def cond(image, latents, off_x, off_y):
            if image is None and latents is None:
                return None, None
            if latents is None:
                latents = self._encode_vae_image(self._preprocess_image(image).to(device), generator)
            latents = latents.to(device=device, dtype=dtype)
            if latents.shape[0] == 1 and batch_size > 1:
                latents = latents.repeat(batch_size, 1, 1, 1)
            _, C, h, w = latents.shape
            packed = self._pack_latents(latents, batch_size, C, h, w)
            ids = self._prepare_latent_image_ids(batch_size, h // 2, w // 2, device, dtype, offset_x=off_x, offset_y=off_y)
            return packed, ids

        dual_l, dual_ids = cond(dual_image, dual_latents, WL // 2, HL // 2)
        ctrl_l, ctrl_ids = cond(control_image, control_latents, 0, HL // 2)
        return noise_latents, noise_ids, dual_l, dual_ids, None, None, ctrl_l, ctrl_ids

    # ------------------------------------------------------------------ properties the reference touches
    @property
    def guidance_scale(self):
        return self._guidance_scale

    @property
    def joint_attention_kwargs(self):
        return self._joint_attention_kwargs

    @property
    def num_timesteps(self):
        return self._num_timesteps

    @property
    def interrupt(self):
        return self._interrupt

    # ------------------------------------------------------------------ __call__ (:404-700)
    @torch.no_grad()
    def __call__(self, prompt=None, prompt_2=None, dual_image=None, redux_image=None, control_image=None,
                 height: Optional[int] = None, width: Optional[int] = None, n_rows: Optional[int] = None,
                 n_cols: Optional[int] = None, num_inference_steps: int = 28, timesteps: List[int] = None,
                 guidance_scale: float = 3.5, num_images_per_prompt: Optional[int] = 1, generator=None,
                 latents=None, prompt_embeds=None, pooled_prompt_embeds=None, output_type: Optional[str] = "pil",
                 return_dict: bool = True, joint_attention_kwargs: Optional[Dict[str, Any]] = None,
                 callback_on_step_end: Optional[Callable] = None,
                 callback_on_step_end_tensor_inputs: List[str] = ["latents"], max_sequence_length: int = 512,
                 control_latents=None, dual_latents=None):
        height = height or self.default_sample_size * self.vae_scale_factor
        width = width or self.default_sample_size * self.vae_scale_factor
        if height % (self.vae_scale_factor * 2) or width % (self.vae_scale_factor * 2):
            raise ValueError(f"`height` and `width` have to be divisible by 16 but are {height} and {width}.")
        if max_sequence_length is not None and max_sequence_length > 512:
            raise ValueError(f"`max_sequence_length` cannot be greater than 512 but is {max_sequence_length}")
        if prompt is None and prompt_embeds is None:
            raise ValueError("Provide either `prompt` or `prompt_embeds`.")
        self._guidance_scale, self._joint_attention_kwargs, self._interrupt = guidance_scale, joint_attention_kwargs, False
        batch_size = 1 if isinstance(prompt, str) else (len(prompt) if prompt is not None else prompt_embeds.shape[0])
        if batch_size * (num_images_per_prompt or 1) != 1:
            raise NotImplementedError("batch_size 1 per GPU, like the reference (pipeline.py:87,125); shard batches over ranks")
        device, dtype = self.device, torch.bfloat16
        eng = self._engine()
        cfg = eng.cfg
        # :538-543 -- text encoders are None: zero embeddings
        pooled = torch.zeros(cfg.pooled_projection_dim, device=device) if pooled_prompt_embeds is None or self.text_encoder is None \
            else pooled_prompt_embeds.reshape(-1).float()
        enc = torch.zeros(max_sequence_length, cfg.joint_attention_dim, device=device, dtype=dtype) \
            if prompt_embeds is None or self.text_encoder_2 is None else prompt_embeds[0].to(dtype)
        s_txt = enc.shape[0]
        (noise_latents, noise_ids, dual_l, dual_ids, _, _, ctrl_l, ctrl_ids) = self.prepare_latents_and_image_ids(
            1, cfg.in_channels // 4, height, width, dtype, device, generator, dual_image=dual_image,
            redux_image=redux_image, control_image=control_image, dual_latents=dual_latents, control_latents=control_latents)
        if latents is not None:
            noise_latents = latents.to(device=device, dtype=dtype)
        conds = [(l, i) for l, i in ((ctrl_l, ctrl_ids), (dual_l, dual_ids)) if l is not None]   # :580-582 cat[control, dual]
        cond_latents = torch.cat([c[0] for c in conds], dim=1) if conds else None
        ids = torch.cat([noise_ids] + [c[1] for c in conds], dim=0)
        # :594-610
        sigmas = np.linspace(1.0, 1 / num_inference_steps, num_inference_steps)
        s_noise = noise_latents.shape[1]
        sc = self.scheduler.config
        mu = calculate_shift(s_noise, sc.base_image_seq_len, sc.max_image_seq_len, sc.base_shift, sc.max_shift)
        ts = self.scheduler.set_timesteps(sigmas, mu)
        self._num_timesteps = len(ts)
        sig = self.scheduler.sigmas
        lat = noise_latents if cond_latents is None else torch.cat([noise_latents, cond_latents], dim=1)
        lat = lat[0].contiguous()
        txt_ids = torch.zeros(s_txt, 3, device=device, dtype=torch.float32)
        eng.prepare(torch.cat([txt_ids, ids.to(device).float()], dim=0), enc, pooled, s_txt=s_txt)
        if callback_on_step_end is None:
            eng.denoise_(lat, s_noise, sig, guidance_scale)            # whole loop inside the C ABI
        else:
            for i in range(len(ts)):
                if self._interrupt:
                    continue
                eng.denoise_(lat, s_noise, sig[i:i + 2], guidance_scale)
                kw = {"latents": lat[None]}
                outs = callback_on_step_end(self, i, torch.tensor(ts[i]), {k: kw.get(k) for k in callback_on_step_end_tensor_inputs})
                if outs and outs.get("latents") is not None:
                    lat = outs["latents"][0].contiguous()
        out = lat[None, :s_noise]
        if output_type == "latent":
            image = out
        else:
            if self.vae is None:
                raise RuntimeError("PBRFluxPipeline: output_type != 'latent' needs a VAE")
            z = self._unpack_latents(out, height, width, self.vae_scale_factor)
            z = (z / self.vae.scaling_factor) + self.vae.shift_factor
            image = self.vae.decode(z)
            image = self._postprocess(image, output_type)
        if not return_dict:
            return (image,)
        return PBRFluxPipelineOutput(images=image)

    @staticmethod
    def _postprocess(image: torch.Tensor, output_type: str):
        """VaeImageProcessor.postprocess [ext]: (x/2+0.5).clamp(0,1); 'pil' -> round(255 x) uint8 PIL images."""
        image = (image.float() / 2 + 0.5).clamp(0, 1)
        if output_type == "pt":
            return image
        arr = image.permute(0, 2, 3, 1).cpu().numpy()
        if output_type == "np":
            return arr
        from PIL import Image
        return [Image.fromarray((a * 255).round().astype("uint8")) for a in arr]
