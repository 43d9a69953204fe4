"""Runs the reference's OWN Python (under /root/reference, in this container only) with its absent third-party packages
stubbed, so that the in-tree parts of the hot path can be executed and their outputs committed as golden fixtures.

TEST INFRASTRUCTURE.  Never imported by the product, by `-m gpu` tests, by smoke() or by bench.py: /root/reference does not
exist on the GPU box.  Used by `make_reference_golden.py` only.

What is real and what is substituted:
* REAL (imported from /root/reference, unmodified): `texturetools.render.nvdiffrast.renderer_inverse.NVDiffRendererInverse`
  (`mv_to_pcd`, `uv_to_pcd`, `bake_mv_to_uv_reproject_blur`, `bake_mv_to_uv_kdtree`, `get_boundary_mask`, `infer`),
  `texturetools.mesh.structure_v2.PBRMesh`, `texturetools.pcd.structure.PointCloud`, `texturetools.pcd.knn.knn`,
  `texturetools.texture.stitching.mip.pull_push`, `texturetools.image.lens_blur.lens_blur_torch`,
  `texturetools.camera.{conversion,generator}`, `texturetools.raytracing.RayTracing` (the dispatcher),
  `flux_piplines.{texturing,delight}.pipeline.PBRFluxPipeline` (`__call__`, `prepare_latents_and_image_ids`, pack / unpack / ids,
  `calculate_shift`, `retrieve_timesteps`) and `flux_piplines.texturing.attention_processor`.
* SUBSTITUTED [ext] (packages absent from the reference tree and from this image; each stand-in is the oracle's restatement,
  so these pieces stay "parity unpinned"): `nvdiffrast.torch.rasterize / interpolate` -> oracle/bake_ref.c; the Slang LBVH
  tracer behind `APRMISRayTracing` (slangtorch cannot compile here) -> oracle/bake_ref.c restatement of the in-tree .slang
  sources; `torch_kdtree.build_kd_tree(...).query` -> exact brute force; `diffusers` FluxPipeline base class /
  FluxTransformer2DModel / FlowMatchEulerDiscreteScheduler / AutoencoderKL / VaeImageProcessor / randn_tensor ->
  oracle/flux_dit.py, flux_sampler.py, vae.py.
* Everything else that is merely imported at module level (trimesh, open3d, pymeshlab, cupy, imageio, ...) -> inert stubs.
"""
from __future__ import annotations

import importlib.abc
import importlib.machinery
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

ABSENT = ("nvdiffrast", "trimesh", "pymeshlab", "open3d", "gpytoolbox", "cupy", "imageio", "slangtorch", "torch_kdtree", "diffusers", "pyexr", "rembg", "fpsample", "xatlas", "peft", "lpips", "open_clip", "kornia", "pygltflib", "fast_simplification", "pyfqmr", "mcubes", "skimage", "timeout_decorator", "pymeshfix", "igl", "loguru", "omegaconf", "lightning", "pytorch_lightning", "safetensors_stub_never")


class _StubMeta(type):
    def __getattr__(cls, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _make_stub(f"{cls.__name__}.{name}")


def _make_stub(name):
    return _StubMeta(name.split(".")[-1], (), {"__init__": lambda self, *a, **k: None, "__call__": lambda self, *a, **k: self,
                                               "__getattr__": lambda self, n: _make_stub(n)() if not n.startswith("__") else (_ for _ in ()).throw(AttributeError(n))})


class _StubModule(types.ModuleType):
    """Inert module: any attribute is a fresh dummy class (usable as a base class, in isinstance and in annotations)."""
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        v = _make_stub(name)
        setattr(self, name, v)
        return v


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path=None, target=None):
        if fullname.split(".")[0] in ABSENT:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        m = _StubModule(spec.name)
        m.__path__ = []
        return m

    def exec_module(self, module):
        pass


_installed = False


def install():
    """Put the stubs and the reference on the import path (idempotent).  Real packages always win: only names in ABSENT that
    fail to import are stubbed."""
    global _installed
    if _installed:
        return
    if not os.path.isdir(REF):
        raise RuntimeError(f"{REF} is absent: the reference harness only runs in the build container")
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    from transformers import (CLIPImageProcessor, CLIPTextModel, CLIPTokenizer, CLIPVisionModelWithProjection,   # noqa: F401
                              T5EncoderModel, T5TokenizerFast)     # real package: resolve its lazy imports before any stub exists
    sys.meta_path.append(_StubFinder())            # appended: consulted only after the real finders fail
    sys.path.insert(0, os.path.join(REF, "TextureTools"))
    sys.path.append(REF)                           # `flux_piplines` of the reference is imported under an alias, see flux()
    _install_nvdiffrast()
    _install_torch_kdtree()
    _install_diffusers()
    _installed = True


# --------------------------------------------------------------------------------------------------------------------------
# nvdiffrast.torch [ext] -> oracle rasteriser / interpolator
# --------------------------------------------------------------------------------------------------------------------------
def _install_nvdiffrast():
    from oracle import bake as ob
    import importlib
    dr = importlib.import_module("nvdiffrast.torch")        # the stub

    class _Ctx:
        def __init__(self, device=None):
            self.device = device

    def rasterize(ctx, pos, tri, resolution, ranges=None, grad_db=True):
        pos_np = pos.detach().cpu().numpy().astype(np.float32)
        if pos_np.ndim == 2:
            pos_np = pos_np[None]
        out = ob.rasterize(np.ascontiguousarray(pos_np), np.ascontiguousarray(tri.detach().cpu().numpy().astype(np.int32)),
                           int(resolution[0]), int(resolution[1]))
        return torch.from_numpy(out), None

    def interpolate(attr, rast, tri, rast_db=None, diff_attrs=None):
        a = np.ascontiguousarray(attr.detach().cpu().numpy().astype(np.float32))
        out = ob.interpolate(a, np.ascontiguousarray(rast.detach().cpu().numpy()),
                             np.ascontiguousarray(tri.detach().cpu().numpy().astype(np.int32)))
        return torch.from_numpy(out), None

    dr.RasterizeCudaContext = _Ctx
    dr.RasterizeGLContext = _Ctx
    dr.rasterize = rasterize
    dr.interpolate = interpolate


# --------------------------------------------------------------------------------------------------------------------------
# torch_kdtree [ext] -> exact brute force (ascending distance, lowest index on ties)
# --------------------------------------------------------------------------------------------------------------------------
def _install_torch_kdtree():
    from oracle import bake as ob
    import torch_kdtree

    class _Tree:
        def __init__(self, src):
            self.src = src.detach().float().cpu()

        def query(self, dst, nr_nns_searches=1):
            dist, idx = ob.nearest_k(self.src, dst.detach().float().cpu(), int(nr_nns_searches))
            return dist, idx.to(torch.int32)

    torch_kdtree.build_kd_tree = lambda src, *a, **k: _Tree(src)


# --------------------------------------------------------------------------------------------------------------------------
# the Slang tracer behind APRMISRayTracing -> oracle restatement (slangtorch is absent; the .slang sources are restated in C)
# --------------------------------------------------------------------------------------------------------------------------
def _patch_aprmis():
    from oracle import bake as ob
    import slangtorch
    slangtorch.loadModule = lambda *a, **k: _make_stub("SlangModule")()
    from texturetools.raytracing import rt_aprmis

    def _build(self, vertices, faces):
        self.vertices = vertices.float().contiguous().cpu()
        self.faces = faces.int().contiguous().cpu()
        self._v = np.ascontiguousarray(self.vertices.numpy(), np.float32)
        self._f = np.ascontiguousarray(self.faces.numpy(), np.int32)
        self.LBVHNode_info, self.LBVHNode_aabb, _ = ob.lbvh_build(self._v, self._f)

    def intersects_closest(self, rays_o, rays_d):
        rays_o, rays_d = torch.broadcast_tensors(rays_o, rays_d)
        batch_shape = rays_o.shape[:-1]
        ro = np.ascontiguousarray(rays_o.reshape(-1, 3).cpu().numpy(), np.float32)
        rd = np.ascontiguousarray(rays_d.reshape(-1, 3).cpu().numpy(), np.float32)
        hit, tid, pos, uv = ob.intersect(self._v, self._f, self.LBVHNode_info, self.LBVHNode_aabb, ro, rd)
        return (torch.from_numpy(np.asarray(hit)).bool().reshape(*batch_shape), None,
                torch.from_numpy(np.asarray(tid).astype(np.int64)).reshape(*batch_shape),
                torch.from_numpy(np.asarray(pos)).reshape(*batch_shape, 3), torch.from_numpy(np.asarray(uv)).reshape(*batch_shape, 2))

    rt_aprmis.APRMISRayTracing.__init__ = _build
    rt_aprmis.APRMISRayTracing.update_raw = _build
    rt_aprmis.APRMISRayTracing.intersects_closest = intersects_closest


def texturetools():
    """-> the reference's `texturetools.render.nvdiffrast.renderer_inverse` module, ready to run on CPU."""
    install()
    from texturetools.geometry import utils as gu          # the only CUDA assumption on the path: default device='cuda'
    gu.to_tensor_f.__defaults__ = ("cpu",)
    gu.to_tensor_i.__defaults__ = ("cpu",)
    _patch_aprmis()
    from texturetools.render.nvdiffrast import renderer_inverse as ri
    ri.link_rgb_to_mesh = lambda src_path, rgb_path, dst_path=None: None        # trimesh I/O: not part of the arithmetic
    ri.link_pbr_to_mesh = lambda *a, **k: None
    return ri


# --------------------------------------------------------------------------------------------------------------------------
# diffusers [ext] -> oracle restatements behind the names the reference imports
# --------------------------------------------------------------------------------------------------------------------------
class _Cfg(dict):
    __getattr__ = dict.__getitem__


def _install_diffusers():
    import contextlib
    from importlib import import_module as im          # `from pkg import sub` on a stub package would return a dummy class
    du, tu = im("diffusers.utils"), im("diffusers.utils.torch_utils")
    pf, sched = im("diffusers.pipelines.flux.pipeline_flux"), im("diffusers.schedulers.scheduling_flow_match_euler_discrete")
    ip = im("diffusers.image_processor")
    du.USE_PEFT_BACKEND = True
    du.is_torch_xla_available = lambda: False
    du.replace_example_docstring = lambda doc: (lambda fn: fn)
    du.scale_lora_layers = lambda *a, **k: None
    du.unscale_lora_layers = lambda *a, **k: None

    class _Log:
        def __getattr__(self, n):
            return lambda *a, **k: None
    du.logging = types.SimpleNamespace(get_logger=lambda name: _Log())

    class BaseOutput(dict):
        def __init__(self, **kw):
            super().__init__(**kw)
            self.__dict__.update(kw)
    du.BaseOutput = BaseOutput

    tu.is_compiled_module = lambda m: False
    tu.is_torch_version = lambda op, v: True

    def randn_tensor(shape, generator=None, device=None, dtype=None, layout=None):
        # diffusers.utils.torch_utils.randn_tensor [ext]: draws on the generator's device, then moves
        return torch.randn(shape, generator=generator, dtype=dtype).to(device)
    tu.randn_tensor = randn_tensor

    class FluxPipeline:
        """The slice of diffusers.FluxPipeline [ext] that PBRFluxPipeline.__call__ touches."""
        _callback_tensor_inputs = ["latents", "prompt_embeds"]

        def __init__(self, scheduler, vae, text_encoder, tokenizer, text_encoder_2, tokenizer_2, transformer, **kw):
            self.scheduler, self.vae, self.text_encoder, self.tokenizer = scheduler, vae, text_encoder, tokenizer     # register_modules [ext]
            self.text_encoder_2, self.tokenizer_2, self.transformer = text_encoder_2, tokenizer_2, transformer

        def check_inputs(self, *a, **k):
            return None

        @property
        def _execution_device(self):
            return torch.device("cpu")

        @property
        def joint_attention_kwargs(self):
            return self._joint_attention_kwargs

        @property
        def interrupt(self):
            return self._interrupt

        @property
        def guidance_scale(self):
            return self._guidance_scale

        def encode_prompt(self, prompt=None, prompt_2=None, prompt_embeds=None, pooled_prompt_embeds=None, device=None,
                          num_images_per_prompt=1, max_sequence_length=512, lora_scale=None):
            # [ext] with given embeddings: repeat per image, text ids = zeros [S_txt, 3]
            text_ids = torch.zeros(prompt_embeds.shape[1], 3).to(device=device, dtype=prompt_embeds.dtype)
            return prompt_embeds, pooled_prompt_embeds, text_ids

        @contextlib.contextmanager
        def progress_bar(self, total=None):
            yield types.SimpleNamespace(update=lambda *a: None)

        def maybe_free_model_hooks(self):
            return None
    pf.FluxPipeline = FluxPipeline
    pf.EXAMPLE_DOC_STRING = ""

    class FlowMatchEulerDiscreteScheduler:
        """[ext] FLUX.1-dev scheduler_config; set_timesteps(sigmas=, mu=) with dynamic shifting; step = fp32 Euler."""
        order = 1

        def __init__(self):
            self.config = _Cfg(base_image_seq_len=256, max_image_seq_len=4096, base_shift=0.5, max_shift=1.15,
                               num_train_timesteps=1000, shift=3.0, use_dynamic_shifting=True)
            self._step_index = None

        def set_timesteps(self, num_inference_steps=None, device=None, sigmas=None, mu=None, timesteps=None):
            import math
            s = np.array(sigmas).astype(np.float32) if sigmas is not None else None
            s = torch.from_numpy(s).to(torch.float32)
            s = math.exp(mu) / (math.exp(mu) + (1 / s - 1) ** 1.0)          # time_shift(mu, 1.0, sigmas) in fp32
            self.timesteps = (s * 1000.0).to(device)
            self.sigmas = torch.cat([s, torch.zeros(1)]).to(device)
            self._step_index = None

        def step(self, model_output, timestep, sample, return_dict=False):
            if self._step_index is None:
                self._step_index = int((self.timesteps == timestep).nonzero()[0].item())
            x = sample.to(torch.float32)
            s, sn = self.sigmas[self._step_index], self.sigmas[self._step_index + 1]
            prev = x + (sn - s) * model_output
            self._step_index += 1
            return (prev.to(model_output.dtype),)
    sched.FlowMatchEulerDiscreteScheduler = FlowMatchEulerDiscreteScheduler

    class VaeImageProcessor:
        """[ext] preprocess: PIL -> [1,3,H,W] in [-1,1] (no resize when sizes match); postprocess('pil'): (x/2+0.5).clamp -> uint8."""
        def __init__(self, vae_scale_factor=8):
            self.vae_scale_factor = vae_scale_factor

        def preprocess(self, image, height=None, width=None):
            from PIL import Image
            a = np.asarray(image.convert("RGB").resize((width, height), Image.LANCZOS) if image.size != (width, height) else image.convert("RGB"))
            return torch.from_numpy(a.astype(np.float32) / 255.0).permute(2, 0, 1)[None] * 2.0 - 1.0

        def postprocess(self, image, output_type="pil"):
            from PIL import Image
            x = (image.float() / 2 + 0.5).clamp(0, 1)
            if output_type == "pt":
                return x
            a = (x.permute(0, 2, 3, 1).numpy() * 255).round().astype("uint8")
            return a if output_type == "np" else [Image.fromarray(i) for i in a]
    ip.VaeImageProcessor = VaeImageProcessor
    ip.PipelineImageInput = object


def flux(task="texturing"):
    """-> the reference's `flux_piplines.<task>.pipeline` module (imported under the alias `_ref_flux_piplines`, because this
    repo ships drop-in modules under the same `flux_piplines` name)."""
    install()
    import importlib.util
    alias = "_ref_flux_piplines"
    if alias not in sys.modules:
        pkg = types.ModuleType(alias)
        pkg.__path__ = [os.path.join(REF, "flux_piplines")]
        sys.modules[alias] = pkg
    return importlib.import_module(f"{alias}.{task}.pipeline")


def flux_attention(task="texturing"):
    install()
    import importlib
    flux(task)
    return importlib.import_module(f"_ref_flux_piplines.{task}.attention_processor")


# --------------------------------------------------------------------------------------------------------------------------
# export_condition / top-level pipeline glue: needs 'cuda' -> 'cpu' (this container has no GPU) and a mesh without trimesh
# --------------------------------------------------------------------------------------------------------------------------
_cuda_remapped = False


def remap_cuda_to_cpu():
    """The reference hard-codes device='cuda' in places (`.to(device='cuda')`, `image_to_tensor(..., device='cuda')`); in
    this CPU-only container those calls are redirected to 'cpu'.  Arithmetic is untouched."""
    global _cuda_remapped
    if _cuda_remapped:
        return
    _to, _as_tensor = torch.Tensor.to, torch.as_tensor

    def fix(x):
        return "cpu" if (isinstance(x, str) and x.startswith("cuda")) or (isinstance(x, torch.device) and x.type == "cuda") else x

    def to(self, *a, **k):
        return _to(self, *[fix(x) for x in a], **{n: fix(v) for n, v in k.items()})

    def as_tensor(*a, **k):
        return _as_tensor(*a, **{n: fix(v) for n, v in k.items()})
    _load = torch.load

    def load(*a, **k):
        return _load(*a, **{n: fix(v) for n, v in k.items()})
    torch.Tensor.to = to
    torch.as_tensor = as_tensor
    torch.load = load
    torch.cuda.empty_cache = lambda: None
    torch.Tensor.cuda = lambda self, *a, **k: self
    _cuda_remapped = True


def video_exporter(vertices: np.ndarray, faces: np.ndarray, vertex_normals: np.ndarray):
    """-> the reference's `VideoExporter` (video/export_nvdiffrast_video.py) whose mesh loader hands back the given arrays as
    the reference's own `Mesh` (mesh/structure.py:306-).  trimesh [ext] would supply `vertex_normals`; they are an input here."""
    install()
    remap_cuda_to_cpu()
    texturetools()
    from texturetools.mesh import structure as ms
    from texturetools.video import export_nvdiffrast_video as ev

    def from_trimesh(_):
        m = ms.Mesh(v_pos=torch.from_numpy(vertices).float(), t_pos_idx=torch.from_numpy(faces).long())
        m._v_nrm = torch.from_numpy(vertex_normals).float()
        return types.SimpleNamespace(mesh=m)
    ev.load_whole_mesh = lambda path: None
    ev.Texture = types.SimpleNamespace(from_trimesh=from_trimesh)
    return ev.VideoExporter()


def top_level_pipeline():
    """-> the reference's top-level `pipeline.py` module (aliased `_ref_pipeline`)."""
    install()
    remap_cuda_to_cpu()
    texturetools()
    import importlib.util
    if "_ref_pipeline" in sys.modules:
        return sys.modules["_ref_pipeline"]
    sys.path.insert(0, REF)        # `from TextureTools.texturetools...` / `from TSD_SR...` resolve against the reference root
    spec = importlib.util.spec_from_file_location("_ref_pipeline", os.path.join(REF, "pipeline.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules["_ref_pipeline"] = mod
    spec.loader.exec_module(mod)
    return mod
