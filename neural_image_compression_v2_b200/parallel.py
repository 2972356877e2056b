"""Multi-GPU plumbing of the path: one process per GPU, `torch.distributed` for the rendezvous.

* Decode shards with NO collective: texels are independent, grids (+ decoder) are replicated on every rank, and a
  rank decodes a contiguous band of rows of the frame (`shard_rows`), a range of frames of a sequence or a range
  of random-access queries (`shard_range`).  `gather_bands` exists only for callers that want the whole frame on
  every rank; it is not part of the timed decode path.
* Training is data parallel with ONE exchange step per training step: `DataParallelPlan` makes every rank draw
  the SAME LOD (shapes and the active pyramid level must agree, image_compression.py:221-226 + :29-34) and
  DIFFERENT crop origins / noise streams, and `all_reduce_flat` sums the flat `[dG0 | dG1 | dMLP | loss]` buffer
  that `FusedTrainer` fills (gradients are pre-scaled by the GLOBAL sample count inside the kernel).

Nothing here computes on tensors' values except through torch.distributed; the host logic is exercised on CPU
with the gloo backend at world_size 2 (tests/test_parallel_gloo.py).
"""
import math
import random

import torch
import torch.distributed as dist

TILE_ROWS = 8          # the tensor-core decode tiles 8 x 16 texels; bands aligned to it stay on the fast path


def world_info(group=None):
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def shard_range(n, rank, world):
    """[start, stop) of `n` independent units (frames, queries) for `rank`: contiguous, sizes differ by <= 1."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError(f"rank {rank} of world {world}")
    base, extra = divmod(max(n, 0), world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def shard_rows(size, rank, world, align=TILE_ROWS):
    """(row0, rows) of a frame with `size` rows (first image axis) for `rank`: contiguous bands whose boundaries
    are multiples of `align` (so every band stays on the aligned tensor-core path); the bands tile [0, size)."""
    if align < 1:
        raise ValueError("align must be >= 1")
    units = (size + align - 1) // align
    u0, u1 = shard_range(units, rank, world)
    r0, r1 = min(u0 * align, size), min(u1 * align, size)
    return r0, r1 - r0


def decode_band(fp, decoder, mip_level=0, size=None, rank=None, world=None, group=None, **kw):
    """This rank's band of a 2-D frame: rows [row0, row0+rows) x all columns.  Returns (row0, band) with band
    `[rows, size, Cout]`; an empty band (more ranks than tile rows) is returned as a 0-row tensor."""
    from . import image_compression as ic
    from . import var2
    r, w = world_info(group)
    rank = r if rank is None else rank
    world = w if world is None else world
    if size is None:
        size = var2.IMAGE_SIZE // pow(2, mip_level)
    row0, rows = shard_rows(size, rank, world)
    if rows == 0:
        cout = decoder.parameters_list()[4].shape[0]
        return row0, torch.empty((0, size, cout), dtype=kw.get("out_dtype", torch.float32), device=fp[0].device)
    return row0, ic.decode(fp, decoder, mip_level, size=(rows, size), origin=(row0, 0), **kw)


def decode_slab(fp, decoder, volume, mip_level=0, rank=None, world=None, group=None, **kw):
    """This rank's slab of a 3-D volume of `volume` = (S0, S1, S2) texels: frames [f0, f0 + n) along the FIRST axis
    (BASELINE config 5: frame-sharded video decode).  A rank only touches the grid nodes under its slab, but the grids
    are small enough to replicate.  Returns (f0, slab) with slab `[n, S1, S2, Cout]`."""
    from . import image_compression as ic
    r, w = world_info(group)
    rank = r if rank is None else rank
    world = w if world is None else world
    f0, f1 = shard_range(volume[0], rank, world)
    if f1 == f0:
        cout = decoder.parameters_list()[4].shape[0]
        return f0, torch.empty((0, volume[1], volume[2], cout), dtype=kw.get("out_dtype", torch.float32), device=fp[0].device)
    return f0, ic.decode(fp, decoder, mip_level, size=(f1 - f0, volume[1], volume[2]), origin=(f0, 0, 0), **kw)


def gather_bands(band, size, group=None):
    """Assembles the full frame on every rank from the per-rank bands (all_gather of padded bands).  Convenience
    for callers that need it; decode itself never communicates."""
    rank, world = world_info(group)
    if world == 1:
        return band
    rows_max = max(shard_rows(size, r, world)[1] for r in range(world))
    pad = torch.zeros((rows_max,) + tuple(band.shape[1:]), dtype=band.dtype, device=band.device)
    pad[:band.shape[0]] = band
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    return torch.cat([parts[r][:shard_rows(size, r, world)[1]] for r in range(world)], dim=0)


class DataParallelPlan:
    """Host-side schedule of data-parallel training.

    Every rank constructs it with the same `seed`; `next_lod()` then returns the same LOD sequence everywhere
    (reference: accumulator of UNIFORM_DISTRIBUTION_RATE, image_compression.py:221-226; uniform draw or
    P(lod = k) ~ 4^-k, :29-34), while `crop_origins()` draws from a rank-specific stream so ranks see different
    crops.  `noise_seed` is the Philox key this rank passes to nic_train_step."""

    def __init__(self, max_mip_level, uniform_rate=0.05, seed=0, rank=None, world=None, group=None):
        r, w = world_info(group)
        self.rank = r if rank is None else rank
        self.world = w if world is None else world
        self.group = group
        self.max_mip_level = max_mip_level
        self.uniform_rate = uniform_rate
        self._lod_rng = random.Random(seed)                                # shared stream: identical on all ranks
        self._crop_rng = torch.Generator().manual_seed(seed * 1000003 + 7919 * self.rank + 1)   # per-rank stream
        self.noise_seed = seed + 7919 * self.rank
        self._acc = 0.0

    def next_lod(self):
        self._acc += self.uniform_rate
        if self._acc >= 1.0:
            self._acc -= 1.0
            return self._lod_rng.randint(0, self.max_mip_level)
        u = self._lod_rng.random()
        lod = int(math.floor(-math.log2(u) / 2)) if u > 0 else self.max_mip_level
        return min(lod, self.max_mip_level)

    def crop_origins(self, data_size, crop_size, num_crops, dim):
        """`num_crops` integer origins in [0, data_size - crop_size] per axis (image_compression.py:40-41)."""
        return torch.randint(0, data_size - crop_size + 1, (num_crops, dim), generator=self._crop_rng)

    def global_samples(self, local_samples):
        """Denominator of the MSE mean under DP: every rank contributes the same number of samples per step."""
        return local_samples * self.world

    def all_reduce_flat(self, flat):
        """THE exchange step: one sum all-reduce of the flat gradient buffer."""
        if self.world > 1:
            dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)
        return flat
