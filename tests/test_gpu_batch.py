"""GPU: the multi-asset data plane (reference run.py:5-10 runs assets one after another on one GPU).
`CustomRGBTextureFullPipeline.run_batch` shards independent assets over the ranks, all-gathers the ACTUAL VAE-decoded uint8
view tiles once and bakes per rank from the gathered tiles.  The 2-rank NCCL run must reproduce the 1-GPU tiles and atlases
byte for byte (asset g always draws from seed base + g).  Needs >= 2 GPUs for the NCCL half; with one GPU only the
single-process half runs."""
import os
import sys

import numpy as np
import pytest
import torch

from tests.bake_meshes import two_spheres

pytestmark = pytest.mark.gpu


def _assets(root, n):
    from PIL import Image
    from unitex_b200.export import save_obj
    v, f, uv, fuv = two_spheres(16, 32)
    out = []
    for g in range(n):
        d = os.path.join(root, f"asset{g}")
        os.makedirs(d, exist_ok=True)
        mesh_path, img_path = os.path.join(d, "mesh.obj"), os.path.join(d, "image.png")
        save_obj(mesh_path, v * (1.0 + 0.1 * g), f, (uv + 1) / 2, fuv)
        Image.fromarray(np.random.default_rng(g).integers(0, 255, (128, 128, 3), dtype=np.uint8)).save(img_path)
        out.append((os.path.join(d, "out"), img_path, mesh_path))
    return out


def _run(assets):
    import warnings
    from pipeline import CustomRGBTextureFullPipeline
    from unitex_b200.flux import FluxConfig
    cfg = FluxConfig(num_layers=1, num_single_layers=1, num_attention_heads=2, joint_attention_dim=256, pooled_projection_dim=64)
    pipe = CustomRGBTextureFullPipeline(pretrain_models=cfg, super_resolutions=False, seed=63)
    pipe.pipeline._num_inference_steps = 2
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        res = pipe.run_batch(assets)
    torch.cuda.synchronize()
    return res, [t.cpu().numpy() for t in pipe.last_tiles]


def _worker(rank, world, port, root_single, root_multi, n):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        res, tiles = _run(_assets(root_multi, n))
        ref = np.load(os.path.join(root_single, "tiles.npz"))
        for g in range(n):                                        # every rank holds every asset's tile, equal to the 1-GPU run's
            assert np.array_equal(tiles[g], ref[f"t{g}"]), f"rank {rank}: tile {g} differs from the single-GPU run"
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_run_batch_single_process(lib, tmp_path):
    from PIL import Image
    assets = _assets(str(tmp_path), 2)
    res, tiles = _run(assets)
    assert len(res) == 2 and len(tiles) == 2 and tiles[0].shape == (1024, 1536, 3) and tiles[0].dtype == np.uint8
    assert not np.array_equal(tiles[0], tiles[1])                 # different seeds, different assets
    for (png, glb), (save_dir, _, _), t in zip(res, assets, tiles):
        assert os.path.exists(png) and os.path.exists(glb)
        assert np.array_equal(np.asarray(Image.open(os.path.join(save_dir, "mv_rgb.png"))), t)   # the gathered tile is what was baked


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (run under gpurun --gpus 2)")
def test_run_batch_two_ranks_equals_one_gpu(lib, tmp_path):
    import torch.multiprocessing as mp
    n = 3                                                         # uneven shards: rank 0 owns assets 0 and 2, rank 1 asset 1
    single, multi = str(tmp_path / "single"), str(tmp_path / "multi")
    res, tiles = _run(_assets(single, n))
    np.savez(os.path.join(single, "tiles.npz"), **{f"t{g}": t for g, t in enumerate(tiles)})
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, single, multi, n), nprocs=2, join=True)
    for g in range(n):                                            # and the baked atlases agree byte for byte
        a = open(os.path.join(single, f"asset{g}", "out", "cache", "wo_LTM", "completed_uv.png"), "rb").read()
        b = open(os.path.join(multi, f"asset{g}", "out", "cache", "wo_LTM", "completed_uv.png"), "rb").read()
        assert a == b, f"asset {g}: baked atlas differs between 1 and 2 GPUs"
