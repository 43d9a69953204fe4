"""GPU, 2 ranks (NCCL): ONE grid's token sequence split over the GPUs (SURVEY 8e "if one grid must span GPUs": Ulysses).
Every Linear / LayerNorm runs on S / P rows per rank, the joint attention (attention_processor.py:81-91) on all rows for
H / P heads per rank, with an all-to-all before and after; the QKV GEMM's epilogue writes straight into the send layout.
The result must be BIT-IDENTICAL to the single-GPU engine: rows and heads are independent in every kernel.
Skipped with fewer than 2 GPUs (run under `gpurun --gpus 2`)."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


def _problem(s_txt):
    from oracle import flux_dit as fd
    from oracle import flux_sampler as fs
    ocfg = fd.FluxConfig.tiny(2, 2, heads=4)
    P = {k: v.to(torch.bfloat16).float() for k, v in fd.init_params(ocfg, 0, norm_weight_std=0.1).items()}
    img_ids = fs.build_ids(32, 32, (32, 32), (16, 16))                    # 256 + 256 + 64 image tokens
    g = torch.Generator().manual_seed(63)
    s_noise = 256
    lat = torch.randn(img_ids.shape[0], 64, generator=g).to(torch.bfloat16)
    ids = torch.cat([torch.zeros(s_txt, 3), img_ids])
    enc = (torch.randn(s_txt, ocfg.joint_attention_dim, generator=g) * 0.5).to(torch.bfloat16)
    return ocfg, P, ids, enc, lat, s_noise, fs.flow_match_sigmas(3, s_noise)


def _engine(ocfg, P):
    from unitex_b200.flux import FluxConfig, FluxTransformer
    cfg = FluxConfig(num_layers=ocfg.num_layers, num_single_layers=ocfg.num_single_layers, num_attention_heads=ocfg.num_attention_heads,
                     joint_attention_dim=ocfg.joint_attention_dim, pooled_projection_dim=ocfg.pooled_projection_dim)
    return FluxTransformer(cfg).load_state_dict(P)


def _worker(rank, world, port, s_txt, ref_path, direct):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from unitex_b200 import parallel as par
        ocfg, P, ids, enc, lat, s_noise, sig = _problem(s_txt)
        eng = _engine(ocfg, P).set_sequence_parallel(par.tile_comm(torch.device("cuda", rank)), direct=direct)
        eng.prepare(ids, enc, None, s_txt=s_txt)
        x = lat.cuda().contiguous()
        v = eng.forward(x, 0.62, 3.5)
        eng.denoise_(x, s_noise, sig, 3.5)
        torch.cuda.synchronize()
        ref = torch.load(ref_path)
        assert torch.equal(v.cpu(), ref["v"]), f"rank {rank}: forward differs from the single-GPU engine"
        assert torch.equal(x.cpu(), ref["x"]), f"rank {rank}: denoised latents differ from the single-GPU engine"
        eng.set_sequence_parallel(None)                      # releases the peer region (collective)
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (run under gpurun --gpus 2)")
@pytest.mark.parametrize("direct", [False, True])  # NCCL all-to-alls / exchanges fused into the epilogues over NVLink peer memory
@pytest.mark.parametrize("s_txt", [128, 64])      # 704 / 640 tokens: rank 0 holds the text rows + part of the image rows
def test_two_rank_sequence_parallel_is_bit_identical(lib, tmp_path, s_txt, direct):
    import torch.multiprocessing as mp
    ocfg, P, ids, enc, lat, s_noise, sig = _problem(s_txt)
    eng = _engine(ocfg, P)
    eng.prepare(ids, enc, None, s_txt=s_txt)
    x = lat.cuda().contiguous()
    v = eng.forward(x, 0.62, 3.5)
    eng.denoise_(x, s_noise, sig, 3.5)
    torch.cuda.synchronize()
    ref_path = str(tmp_path / "ref.pt")
    torch.save({"v": v.cpu(), "x": x.cpu()}, ref_path)
    del eng
    mp.spawn(_worker, args=(2, 29500 + os.getpid() % 2000, s_txt, ref_path, direct), nprocs=2, join=True)


def _pipe_call(sp_comm=None):
    """The public sampler call (PIL in -> VAE encode -> denoise -> VAE decode -> PIL out) on a tiny FLUX + tiny VAE."""
    import numpy as np
    from PIL import Image
    from oracle import vae as ov
    from flux_piplines.texturing.pipeline import PBRFluxPipeline
    from unitex_b200.vae import AutoencoderKLB200
    ocfg, P, *_ = _problem(128)
    vcfg = ov.VaeConfig.tiny()
    VP = {k: v.to(torch.bfloat16).float() for k, v in ov.init_params(vcfg, 5).items()}
    vae = AutoencoderKLB200(VP, vcfg.block_out_channels, vcfg.layers_per_block, vcfg.latent_channels, vcfg.in_channels,
                            vcfg.norm_num_groups, vcfg.scaling_factor, vcfg.shift_factor)
    pipe = PBRFluxPipeline(_engine(ocfg, P), vae)
    g = torch.Generator().manual_seed(1)
    pipe.load_lora_weights({"transformer.transformer_blocks.0.attn.to_q.lora_A.weight": torch.randn(8, 512, generator=g) * 0.05,
                            "transformer.transformer_blocks.0.attn.to_q.lora_B.weight": torch.randn(512, 8, generator=g) * 0.05}, adapter_name="texture")
    pipe.set_adapters(["texture"], [1.0])
    if sp_comm is not None:
        pipe.set_sequence_parallel(sp_comm, direct=True)
    rng = np.random.default_rng(0)
    ctrl = Image.fromarray(rng.integers(0, 255, (256, 256, 3), dtype=np.uint8))
    dual = Image.fromarray(rng.integers(0, 255, (128, 128, 3), dtype=np.uint8))
    img = pipe(prompt="[MVFLUX]", control_image=ctrl, dual_image=dual, height=256, width=256, n_rows=1, n_cols=1,
               num_inference_steps=3, guidance_scale=3.5, max_sequence_length=128, generator=torch.Generator().manual_seed(63)).images[0]
    torch.cuda.synchronize()
    if sp_comm is not None:
        pipe.set_sequence_parallel(None)
    return np.asarray(img)


def _pipe_worker(rank, world, port, ref_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import numpy as np
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from unitex_b200 import parallel as par
        img = _pipe_call(par.tile_comm(torch.device("cuda", rank)))
        assert np.array_equal(img, np.load(ref_path)), f"rank {rank}: the image differs from the single-GPU call"
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (run under gpurun --gpus 2)")
def test_two_rank_pipeline_call_equals_one_gpu(lib, tmp_path):
    """`PBRFluxPipeline.set_sequence_parallel`: the whole public call with a merged adapter, one asset over two GPUs (fused
    peer-memory exchanges), returns the single-GPU image byte for byte on every rank."""
    import numpy as np
    import torch.multiprocessing as mp
    ref_path = str(tmp_path / "img.npy")
    np.save(ref_path, _pipe_call(None))
    mp.spawn(_pipe_worker, args=(2, 29500 + (os.getpid() + 7) % 2000, ref_path), nprocs=2, join=True)


def test_sequence_parallel_off_is_the_default(lib):
    """set_sequence_parallel(None) is the plain engine (and the mode needs a fresh prepare)."""
    ocfg, P, ids, enc, lat, s_noise, sig = _problem(128)
    eng = _engine(ocfg, P)
    eng.prepare(ids, enc, None, s_txt=128)
    x = lat.cuda().contiguous()
    a = eng.forward(x, 0.62, 3.5).clone()
    eng.set_sequence_parallel(None)
    with pytest.raises(Exception):
        eng.forward(x, 0.62, 3.5)                                      # not prepared for the (re)set mode
    eng.prepare(ids, enc, None, s_txt=128)
    assert torch.equal(eng.forward(x, 0.62, 3.5), a)
