"""Host-side logic of the multi-GPU path on CPU: world_size-2 gloo processes (no kernels are launched here; the
kernels' own parity is covered by the -m gpu tests).  Decode: bands tile the frame and need no collective.
Training: identical LOD schedule on every rank, distinct crops, one sum all-reduce of the flat gradient buffer."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from neural_image_compression_v2_b200 import parallel as P


def test_shard_rows_tile_the_frame():
    for size in (1, 7, 8, 64, 100, 512, 4096, 1080):
        for world in (1, 2, 3, 4, 8):
            bands = [P.shard_rows(size, r, world) for r in range(world)]
            assert bands[0][0] == 0
            for (a0, an), (b0, bn) in zip(bands, bands[1:]):
                assert a0 + an == b0                      # contiguous, no overlap
            assert bands[-1][0] + bands[-1][1] == size    # complete
            assert all(r0 % P.TILE_ROWS == 0 for r0, _ in bands if r0 < size)
            full = [n for _, n in bands if n]
            assert max(full) - min(full) <= P.TILE_ROWS or size % P.TILE_ROWS   # balanced to one tile row


def test_shard_range_balanced_and_ragged():
    assert [P.shard_range(10, r, 4) for r in range(4)] == [(0, 3), (3, 6), (6, 8), (8, 10)]
    assert [P.shard_range(2, r, 4) for r in range(4)] == [(0, 1), (1, 2), (2, 2), (2, 2)]
    assert P.shard_range(0, 0, 1) == (0, 0)
    assert P.shard_range(10 ** 9, 7, 8) == (875000000, 10 ** 9)
    with pytest.raises(ValueError):
        P.shard_range(4, 4, 4)


def test_single_process_plan_defaults():
    plan = P.DataParallelPlan(max_mip_level=9, seed=3)
    assert (plan.rank, plan.world) == (0, 1)
    lods = [plan.next_lod() for _ in range(400)]
    assert min(lods) == 0 and max(lods) <= 9
    assert np.mean(np.array(lods) == 0) > 0.6            # P(lod = 0) = 3/4 for the non-uniform draws
    flat = torch.arange(4.0)
    assert plan.all_reduce_flat(flat) is flat             # no process group: no-op


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _frame(size):
    r = torch.arange(size, dtype=torch.float32)
    return (r[:, None] * 1000 + r[None, :])[..., None].repeat(1, 1, 3)


def _worker(rank, world, port, size, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        res = {}
        # ---- decode: this rank's band, then (optionally) the assembled frame
        row0, rows = P.shard_rows(size, rank, world)
        band = _frame(size)[row0:row0 + rows]
        full = P.gather_bands(band, size)
        res["frame_ok"] = bool(torch.equal(full, _frame(size)))
        # ---- training plan: same LODs, different crops
        plan = P.DataParallelPlan(max_mip_level=5, seed=11)
        lods = [plan.next_lod() for _ in range(64)]
        crops = plan.crop_origins(512, 256, 8, 2)
        gathered = [None] * world
        dist.all_gather_object(gathered, (lods, crops.tolist(), plan.noise_seed))
        res["lods_equal"] = all(g[0] == gathered[0][0] for g in gathered)
        res["crops_differ"] = gathered[0][1] != gathered[1][1]
        res["noise_seeds_differ"] = gathered[0][2] != gathered[1][2]
        res["crops_in_range"] = bool((crops >= 0).all() and (crops <= 256).all())
        # ---- the exchange step: gradients pre-scaled by the GLOBAL sample count sum to the single-process mean
        g = torch.Generator().manual_seed(5)
        x = torch.rand(world * 16, 6, generator=g, dtype=torch.float64)     # all samples of the step
        mine = x[rank * 16:(rank + 1) * 16]
        flat = mine.sum(0) / plan.global_samples(16)
        plan.all_reduce_flat(flat)
        res["allreduce_ok"] = bool(torch.allclose(flat, x.mean(0), rtol=1e-12))
        out[rank] = res
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("size", [64, 100])
def test_world2_gloo(size):
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), size, out), nprocs=world, join=True)
    assert len(out) == world
    for rank in range(world):
        for key, ok in out[rank].items():
            assert ok, f"rank {rank}: {key}"
