"""CPU, world_size 2, gloo: the N > 1 host logic of the path (grid sharding, the one all-gather, max-over-ranks timing)."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from unitex_b200 import parallel as par
    mine = par.shard_grids(5, rank, world)
    seeds = [par.grid_seed(63, g) for g in mine]
    tile = torch.full((4, 8), float(rank), dtype=torch.bfloat16)
    tiles = par.all_gather_tiles(tile)
    t = par.max_over_ranks([10.0 + rank, 3.0 - rank], "cpu")
    # uneven shards (3 + 2 grids): one padded gather per batch, every rank ends with all five tiles in grid order
    local = [torch.full((2, 3), float(g), dtype=torch.uint8) for g in mine]
    allt = par.all_gather_grid_tiles(local, 5, (2, 3), torch.uint8, "cpu")
    assert [int(x[0, 0]) for x in allt] == [0, 1, 2, 3, 4] and all(x.shape == (2, 3) for x in allt)
    q.put((rank, mine, seeds, [float(x[0, 0]) for x in tiles], t))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_and_gather():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, g0, s0, t0, m0), (r1, g1, s1, t1, m1) = res
    assert g0 == [0, 2, 4] and g1 == [1, 3] and s0 == [63, 65, 67] and s1 == [64, 66]
    assert t0 == [0.0, 1.0] and t1 == [0.0, 1.0]          # every rank holds every tile, in rank order
    assert m0 == [11.0, 3.0] and m1 == [11.0, 3.0]


def test_single_process_is_identity():
    from unitex_b200 import parallel as par
    t = torch.ones(2, 2)
    assert par.all_gather_tiles(t)[0] is t and par.shard_grids(3, 0, 1) == [0, 1, 2]
    assert par.max_over_ranks([1.5], "cpu") == [1.5]
    tiles = par.all_gather_grid_tiles([torch.full((2,), 7, dtype=torch.uint8)] * 3, 3, (2,), torch.uint8, "cpu")
    assert len(tiles) == 3 and all(int(x[0]) == 7 for x in tiles)
