"""Multi-GPU plumbing for the path: ranks shard INDEPENDENT grids/assets -- all views of one grid share one
attention sequence (flux_piplines/texturing/pipeline.py:630-656), so views are never split -- and the only exchange is one
all-gather of the finished tiles before UV projection: `utx_allgather_tiles` (one ncclAllGather issued by libunitex_b200.so) on
GPUs, torch.distributed/gloo in the CPU tests.  torch.distributed is the process-group plumbing (rendezvous, barriers, the
max-over-ranks reduction of timings)."""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence

import torch
import torch.distributed as dist


class TileComm:
    """The native data plane of the gather: `utx_comm_*` / `utx_allgather_tiles` of libunitex_b200.so (one ncclAllGather on the
    caller's stream).  torch.distributed is only the bootstrap here -- it carries NCCL's 128-byte unique id from rank 0 to the
    others (any out-of-band channel would do) -- and stays the data plane of the CPU (gloo) tests."""

    def __init__(self, device):
        from . import _lib
        self._lib_mod, self.lib = _lib, _lib.load()
        self.device = torch.device(device)
        self.world, self.rank = dist.get_world_size(), dist.get_rank()
        idbuf = torch.zeros(128, dtype=torch.uint8)
        if self.rank == 0:
            raw = (C.c_ubyte * 128)()
            _lib.check(self.lib.utx_comm_unique_id(raw), "utx_comm_unique_id")
            idbuf = torch.tensor(list(raw), dtype=torch.uint8)
        box = [idbuf]
        dist.broadcast_object_list(box, src=0)
        raw = (C.c_ubyte * 128)(*box[0].tolist())
        self._handle = _lib.vp()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.utx_comm_init(C.byref(self._handle), raw, self.world, self.rank), "utx_comm_init")

    def all_gather(self, tile: torch.Tensor) -> torch.Tensor:
        """tile: contiguous CUDA tensor, same shape on every rank -> [world, *tile.shape]."""
        assert tile.is_cuda and tile.is_contiguous()
        out = torch.empty((self.world, *tile.shape), dtype=tile.dtype, device=tile.device)
        nbytes = tile.numel() * tile.element_size()
        with torch.cuda.device(tile.device):
            self._lib_mod.check(self.lib.utx_allgather_tiles(self._handle, tile.data_ptr(), out.data_ptr(), nbytes,
                                                             torch.cuda.current_stream().cuda_stream), "utx_allgather_tiles")
        return out

    def close(self):
        if getattr(self, "_handle", None):
            self.lib.utx_comm_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class PeerRegion:
    """A device buffer of `nbytes` on every rank, each mapped into every other rank's address space (`utx_peer_*`: cudaMalloc +
    CUDA IPC over NVLink P2P).  torch.distributed only carries the 64-byte handles.  `ptrs[r]` = rank r's buffer as seen from
    this process (`ptrs[rank]` is the local allocation).  Zero-filled; one region per engine handle."""

    def __init__(self, nbytes: int, device):
        from . import _lib
        self._lib_mod, self.lib = _lib, _lib.load()
        self.device = torch.device(device)
        self.world, self.rank = dist.get_world_size(), dist.get_rank()
        self.nbytes = int(nbytes)
        self.ptrs: List[int] = [0] * self.world
        with torch.cuda.device(self.device):
            mine = _lib.vp()
            _lib.check(self.lib.utx_peer_alloc(C.byref(mine), self.nbytes), "utx_peer_alloc")
            raw = (C.c_ubyte * 64)()
            _lib.check(self.lib.utx_peer_export(mine, raw), "utx_peer_export")
            handles = [None] * self.world
            dist.all_gather_object(handles, bytes(raw))
            for r, hb in enumerate(handles):
                if r == self.rank:
                    self.ptrs[r] = mine.value
                else:
                    p = _lib.vp()
                    _lib.check(self.lib.utx_peer_import((C.c_ubyte * 64)(*hb), C.byref(p)), "utx_peer_import")
                    self.ptrs[r] = p.value
            torch.cuda.synchronize()
        dist.barrier()                               # every rank has mapped every region before anyone writes

    def close(self):
        if not self.ptrs:
            return
        torch.cuda.synchronize()
        if dist.is_initialized():
            dist.barrier()                           # nobody still writes into a region that is about to go away
        for r, p in enumerate(self.ptrs):
            if p and r != self.rank:
                self.lib.utx_peer_close(p)
        if self.ptrs[self.rank]:
            self.lib.utx_peer_free(self.ptrs[self.rank])
        self.ptrs = []

    def __del__(self):
        try:
            if self.ptrs and self.ptrs[self.rank]:
                for r, p in enumerate(self.ptrs):
                    if p and r != self.rank:
                        self.lib.utx_peer_close(p)
                self.lib.utx_peer_free(self.ptrs[self.rank])
                self.ptrs = []
        except Exception:
            pass


_TILE_COMM: Optional[TileComm] = None


def tile_comm(device) -> Optional[TileComm]:
    """The process's native communicator (created on first use; None without torch.distributed, on CPU, or at world size 1)."""
    global _TILE_COMM
    dev = torch.device(device)
    if not dist.is_initialized() or dist.get_world_size() == 1 or dev.type != "cuda":
        return None
    if _TILE_COMM is None:
        _TILE_COMM = TileComm(dev)
    return _TILE_COMM


def shard_grids(n_grids: int, rank: int, world: int) -> List[int]:
    """Grid indices owned by `rank`: round-robin, so seeds 63, 64, ... (run.py:5) land on ranks 0, 1, ... ."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world {world}")
    return list(range(rank, n_grids, world))


def grid_seed(base_seed: int, grid_index: int) -> int:
    return base_seed + grid_index


def all_gather_tiles(tile: torch.Tensor) -> List[torch.Tensor]:
    """One collective per batch of assets: every rank receives every rank's finished tile (same shape/dtype)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return [tile]
    out = [torch.empty_like(tile) for _ in range(dist.get_world_size())]
    dist.all_gather(out, tile.contiguous())
    return out


def all_gather_grid_tiles(local_tiles: Sequence[torch.Tensor], n_grids: int, tile_shape, dtype, device) -> List[torch.Tensor]:
    """The batch form of the path's one collective (reference pipeline.py:231-291 runs the assets one after another on one
    GPU; here rank r owns grids r, r + world, ... -- `shard_grids`).  Shards are UNEVEN when world does not divide n_grids
    (5 grids on 2 ranks: 3 and 2), and a collective that is entered once per finished grid would hang the rank with fewer
    grids.  So: ONE all-gather per batch over a stack padded to ceil(n_grids / world) slots per rank; the padding slots
    are dropped on the way out by the same round-robin rule.  Returns the n_grids tiles in grid order on every rank."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    mine = shard_grids(n_grids, rank, world)
    if len(local_tiles) != len(mine):
        raise ValueError(f"rank {rank} owns {len(mine)} grids but holds {len(local_tiles)} tiles")
    slots = (n_grids + world - 1) // world
    stack = torch.zeros((slots, *tile_shape), dtype=dtype, device=device)
    for j, t in enumerate(local_tiles):
        stack[j].copy_(t)
    if world == 1:
        return [stack[j] for j in range(n_grids)]
    comm = tile_comm(device)
    if comm is not None:                                                              # GPUs: the C ABI's ncclAllGather
        out = comm.all_gather(stack).reshape(world * slots, *tile_shape)
    else:                                                                             # CPU tests: gloo
        out = torch.empty((world * slots, *tile_shape), dtype=dtype, device=device)  # rank-major concatenation
        dist.all_gather_into_tensor(out, stack)
    return [out[(g % world) * slots + g // world] for g in range(n_grids)]


def max_over_ranks(values: Sequence[float], device) -> List[float]:
    """Timing convention of bench.py: a multi-GPU number is the max over ranks."""
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.tolist()
